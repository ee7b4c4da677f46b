"""Where the fixed cost of one Model.fit() call goes (host profile of the bench's e2e call)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
vocabs = bench.CRITEO_VOCABS
model, KL = bench.build_deepfm_model(vocabs)
model.compile(optimizer=KL.Adam(learning_rate=1e-3), loss=KL.binary_crossentropy)
B, steps = bench.BATCH, 20
g = np.random.RandomState(0)
x = {f"C{i + 1}": g.randint(0, v, (B * steps, 1)).astype(np.int32) for i, v in enumerate(vocabs)}
x.update({f"I{j + 1}": g.rand(B * steps, 1).astype(np.float32) for j in range(bench.N_DENSE)})
y = (g.rand(B * steps) < 0.25).astype(np.float32)
model.fit({k: v[: 10 * B] for k, v in x.items()}, y[: 10 * B], batch_size=B, epochs=1)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    model.fit(x, y, batch_size=B, epochs=1)
    torch.cuda.synchronize()
    print("fit", rep, (time.perf_counter() - t0) * 1e3 / steps, "ms/step")
pr = cProfile.Profile()
pr.enable()
model.fit(x, y, batch_size=B, epochs=1)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
