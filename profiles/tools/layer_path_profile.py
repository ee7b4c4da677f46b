"""DIN / retrieval step through the Keras-like Model: wall per step, GPU busy time per step (sum of kernel durations from the torch profiler),
and the host profile of train_on_batch."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench_models as BM
wl = sys.argv[1] if len(sys.argv) > 1 else "din"
dev = torch.device("cuda", 0)
if wl == "din":
    B = BM.DIN_B
    model = BM.build_din()
    x, y = BM.din_data(B * 4)
else:
    B = BM.RET_B
    model = BM.build_retrieval()
    x, y = BM.retrieval_data(B * 4, BM.RET_ITEMS)
pool = [({k: torch.from_numpy(v[i * B : (i + 1) * B]).to(dev) for k, v in x.items()}, torch.from_numpy(y[i * B : (i + 1) * B]).to(dev)) for i in range(4)]
for s in range(5):
    model.train_on_batch(*pool[s % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(20):
    model.train_on_batch(*pool[s % 4])
torch.cuda.synchronize()
print(f"{wl}: {1e3 * (time.perf_counter() - t0) / 20:.3f} ms per step (wall)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for s in range(4):
        model.train_on_batch(*pool[s % 4])
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total for e in ev) / 4 / 1e3
print(f"GPU busy: {tot:.3f} ms per step over {len(ev) / 4:.0f} kernels/copies")
agg = {}
for e in ev:
    d = e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
    a = agg.setdefault(e.name[:70], [0, 0.0])
    a[0] += 1
    a[1] += d
for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"  {d / 4:9.1f} us/step  x{c / 4:5.1f}  {n}")
pr = cProfile.Profile()
pr.enable()
for s in range(8):
    model.train_on_batch(*pool[s % 4])
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
pstats.Stats(pr).sort_stats("cumulative").print_stats(60)
