import sys, torch
sys.path.insert(0, '.')
from handyrec_b200 import kernels as k
dev = torch.device('cuda:0')
M, Kd, N = 65536, 432, 429
x = torch.randn(M, Kd, device=dev); wt = torch.randn(N, Kd, device=dev) / 20; b = torch.zeros(N, device=dev)
out = torch.zeros(M, 432, device=dev); out_t = torch.zeros(N, M, device=dev)
for mode in ("fwd_t", "fwd"):
    for _ in range(3):
        k.dense_fwd_t(x, wt, b, "relu", out=out[:, :N], out_t=out_t if mode == "fwd_t" else None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        k.dense_fwd_t(x, wt, b, "relu", out=out[:, :N], out_t=out_t if mode == "fwd_t" else None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(mode, ms, "ms", 2 * M * Kd * N / ms / 1e9, "TFLOP/s")
