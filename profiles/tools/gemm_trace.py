import sys, torch
sys.path.insert(0, '.')
from handyrec_b200 import kernels as k
dev = torch.device('cuda:0')
M, Kd, N = 65536, 432, 429
x = torch.randn(M, Kd, device=dev); wt = torch.randn(N, Kd, device=dev) / 20; b = torch.zeros(N, device=dev)
ld = (N + 3) // 4 * 4
out = torch.zeros(M, ld, device=dev)
for _ in range(2):
    k.dense_fwd_t(x, wt, b, "relu", out=out[:, :N])
torch.cuda.synchronize()
