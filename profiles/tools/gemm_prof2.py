import sys, torch
sys.path.insert(0, '.')
from handyrec_b200 import kernels as k
dev = torch.device('cuda:0')
M = 65536
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (Kd, N) in [(432, 429), (432, 256), (256, 128)]:
    x = torch.randn(M, Kd, device=dev); wt = torch.randn(N, Kd, device=dev) / 20; b = torch.zeros(N, device=dev)
    ld = (N + 3) // 4 * 4
    out = torch.zeros(M, ld, device=dev); out_t = torch.zeros(ld, M, device=dev)
    ms = timeit(lambda: k.dense_fwd_t(x, wt, b, "relu", out=out[:, :N], out_t=out_t[:N]))
    ms2 = timeit(lambda: k.dense_fwd_t(x, wt, b, "relu", out=out[:, :N]))
    print(f"fwd   K={Kd} N={N}: {ms*1e3:.0f} us ({2*M*Kd*N/ms/1e9:.0f} TF)  no-Ct {ms2*1e3:.0f} us")
for (Kd, N) in [(432, 429), (432, 256), (256, 128)]:   # dx[M,Kd] = dz[M,N] @ w[Kd,N]^T
    dz = torch.randn(M, (N + 3) // 4 * 4, device=dev)[:, :N]; w = (torch.randn(Kd, (N + 3) // 4 * 4, device=dev) / 20)[:, :N]
    ap = torch.relu(torch.randn(M, Kd, device=dev))
    out = torch.zeros(M, Kd, device=dev); out_t = torch.zeros(Kd, M, device=dev)
    ms = timeit(lambda: k.dense_bwd_x_t(dz, w, a_prev=ap, act_prev="relu", out=out, out_t=out_t))
    ms2 = timeit(lambda: k.dense_bwd_x_t(dz, w, out=out))
    print(f"bwd_x K={Kd} N={N}: {ms*1e3:.0f} us ({2*M*Kd*N/ms/1e9:.0f} TF)  plain {ms2*1e3:.0f} us")
for (Kd, N) in [(432, 429), (432, 256), (256, 128)]:
    xt = torch.randn(Kd, M, device=dev); dzt = torch.randn(N, M, device=dev); dz = torch.randn(M, N, device=dev)
    ms = timeit(lambda: k.dense_bwd_w_t(xt, dzt, dz))
    ms2 = timeit(lambda: k.dense_bwd_w_t(xt, dzt))
    print(f"bwd_w K={Kd} N={N}: {ms*1e3:.0f} us ({2*M*Kd*N/ms/1e9:.0f} TF)  no-bias {ms2*1e3:.0f} us")
