import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import oracle
from handyrec_b200 import kernels as k
dev = torch.device("cuda:0")
B, D = 65536, 16
vocabs = [4, 11, 300, 5000, 250000, 2000000]
g = torch.Generator().manual_seed(7)
tables = [torch.randn(v, D, generator=g) * 0.05 for v in vocabs]
ids = torch.stack([torch.full((B,), min(3, v - 1), dtype=torch.int32) for v in vocabs], 1)
ids[::7, 5] = 123456
dout = torch.randn(B, len(vocabs) * D, generator=g) / 256
for zero_tab in (False, True):
    dt = [(torch.zeros_like(t) if zero_tab else t.clone()).to(dev) for t in tables]
    plan = k.LookupPlan(dt, [(f, 1, "none", f, f * D) for f in range(len(vocabs))])
    plan.backward_update(ids.contiguous().to(dev), dout.to(dev), opt="sgd", lr=0.5)
    torch.cuda.synchronize()
    gr64 = dout[:, :D].double().sum(0)
    gr32 = oracle.embedding_grad_dense(4, ids[:, 0], dout[:, :D])[3]
    base = torch.zeros(D) if zero_tab else tables[0][3]
    got = (base - dt[0].cpu()[3]) / 0.5
    print("zero_tab", zero_tab)
    print(" got-f64 ", ((got.double() - gr64) / gr64.abs().max()).tolist())
    print(" orc-f64 ", ((gr32.double() - gr64) / gr64.abs().max()).tolist())
