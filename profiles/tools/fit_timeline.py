"""GPU-side timeline of Model.fit(dict): when does every step start / end relative to the call, where are the bubbles."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
vocabs = bench.CRITEO_VOCABS
model, KL = bench.build_deepfm_model(vocabs)
model.compile(optimizer=KL.Adam(learning_rate=1e-3), loss=KL.binary_crossentropy)
B, steps = bench.BATCH, int(sys.argv[1]) if len(sys.argv) > 1 else 20
g = np.random.RandomState(0)
x = {f"C{i + 1}": g.randint(0, v, (B * steps, 1)).astype(np.int32) for i, v in enumerate(vocabs)}
x.update({f"I{j + 1}": g.rand(B * steps, 1).astype(np.float32) for j in range(bench.N_DENSE)})
y = (g.rand(B * steps) < 0.25).astype(np.float32)
model.fit({k: v[: 10 * B] for k, v in x.items()}, y[: 10 * B], batch_size=B, epochs=1)
torch.cuda.synchronize()
eng = model._fused.engine
orig = eng.train_step_on_device
ev, cpu = [], []


def wrapped(*a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cpu.append(time.perf_counter())
    e0.record()
    orig(*a)
    e1.record()
    ev.append((e0, e1))


eng.train_step_on_device = wrapped
for rep in range(2):
    ev.clear(); cpu.clear()
    z = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    z.record()
    model.fit(x, y, batch_size=B, epochs=1)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"rep {rep}: fit wall {1e3 * (t1 - t0):.2f} ms; first step starts on GPU at {z.elapsed_time(ev[0][0]):.2f} ms, last ends at {z.elapsed_time(ev[-1][1]):.2f} ms")
    print(" step: cpu_enqueue_ms gpu_start_ms gpu_dur_ms gap_before_ms")
    prev = None
    for i, (a, b) in enumerate(ev):
        s = z.elapsed_time(a)
        gap = s - prev if prev is not None else s
        prev = z.elapsed_time(b)
        print(f"  {i:2d}: {1e3 * (cpu[i] - t0):7.2f} {s:7.2f} {a.elapsed_time(b):6.3f} {gap:6.3f}")
