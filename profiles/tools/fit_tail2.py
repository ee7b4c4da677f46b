"""Phase timeline (engine._timeline: events on whatever stream is current, streams not serialised) of mid vs tail steps of Model.fit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
vocabs = bench.CRITEO_VOCABS
model, KL = bench.build_deepfm_model(vocabs)
model.compile(optimizer=KL.Adam(learning_rate=1e-3), loss=KL.binary_crossentropy)
B, steps = bench.BATCH, 24
g = np.random.RandomState(0)
x = {f"C{i + 1}": g.randint(0, v, (B * steps, 1)).astype(np.int32) for i, v in enumerate(vocabs)}
x.update({f"I{j + 1}": g.rand(B * steps, 1).astype(np.float32) for j in range(bench.N_DENSE)})
y = (g.rand(B * steps) < 0.25).astype(np.float32)
model.fit({k: v[: 10 * B] for k, v in x.items()}, y[: 10 * B], batch_size=B, epochs=1)
torch.cuda.synchronize()
eng = model._fused.engine
for rep in range(2):
    eng._timeline = []
    model.fit(x, y, batch_size=B, epochs=1)
    torch.cuda.synchronize()
    tl, eng._timeline = eng._timeline, None
    # split into steps at every "pack_dense"
    per, cur = [], None
    for name, e in tl:
        if name == "pack_dense":
            cur = []
            per.append(cur)
        cur.append((name, e))
    names = [n for n, _ in per[10]]
    print(f"rep {rep}: {len(per)} steps; ms since the step's first mark")
    print("phase".ljust(28) + " ".join(f"s{i:02d}".rjust(7) for i in (10, steps - 5, steps - 4, steps - 3, steps - 2, steps - 1)))
    for j, n in enumerate(names):
        row = []
        for i in (10, steps - 5, steps - 4, steps - 3, steps - 2, steps - 1):
            row.append(per[i][0][1].elapsed_time(per[i][j][1]))
        print(n.ljust(28) + " ".join(f"{v:7.3f}" for v in row))
