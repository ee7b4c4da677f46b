import sys, os, torch
sys.path.insert(0, '.')
from handyrec_b200 import kernels as K
import bench
dev = torch.device('cuda:0')
B, D = 65536, 16
vocabs = bench.CRITEO_VOCABS
tables = []
for f, v in enumerate(vocabs):
    t = torch.empty(v, D, device=dev); K.init_uniform(t, 7 + f); tables.append(t)
plan = K.LookupPlan(tables, [(f, 1, "none", f, 16 + f * D) for f in range(len(vocabs))])
g = torch.Generator(device=dev).manual_seed(1)
pool = [torch.stack([torch.randint(0, v, (B,), device=dev, generator=g) for v in vocabs], 1).to(torch.int32).contiguous() for _ in range(4)]
X0 = torch.zeros(B, 432, device=dev)
w = torch.randn(D, device=dev); w0 = torch.zeros(1, device=dev)
for _ in range(5):
    plan.forward(pool[0], out=X0, fm=(w, w0), want_fm_sum=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(40):
    plan.forward(pool[i % 4], out=X0, fm=(w, w0), want_fm_sum=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 40
print(os.environ.get("HRB_LOOKUP_KERNEL", "tile"), ms * 1e3, "us", 3432 * B / ms / 1e6, "GB/s", 3432 * B / ms / 1e6 / 6551.7)
