"""Debug helper: the embedding backward with one hot id per table, for growing sets of tables."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import oracle
from handyrec_b200 import kernels as k
dev = torch.device("cuda:0")
B, D = 65536, 16
ALL = [4, 11, 300, 5000, 250000, 2000000]
def run(vocabs, dense_idx=None, tag=""):
    g = torch.Generator().manual_seed(7)
    tables = [torch.randn(v, D, generator=g) * 0.05 for v in vocabs]
    ids = torch.stack([torch.full((B,), min(3, v - 1), dtype=torch.int32) for v in vocabs], 1)
    if vocabs[-1] == 2000000:
        ids[::7, len(vocabs) - 1] = 123456
    dout = torch.randn(B, len(vocabs) * D, generator=g) / 256
    dt = [t.clone().to(dev) for t in tables]
    plan = k.LookupPlan(dt, [(f, 1, "none", f, f * D) for f in range(len(vocabs))])
    dg = None
    if dense_idx is not None:
        dg = torch.zeros_like(dt[dense_idx])
        plan.set_dense_grads([dg if f == dense_idx else None for f in range(len(vocabs))])
    plan.backward_update(ids.contiguous().to(dev), dout.to(dev), opt="sgd", lr=0.5)
    torch.cuda.synchronize()
    out = []
    for f, v in enumerate(vocabs):
        gr = oracle.embedding_grad_dense(v, ids[:, f], dout[:, f * D:(f + 1) * D])
        got = dg.cpu() if f == dense_idx else (tables[f] - dt[f].cpu()) / 0.5
        err = float((got - gr).abs().max() / gr.abs().max())
        out.append(f"{v}:{err:.1e}")
    print(tag, vocabs, "dense", dense_idx, " ".join(out), flush=True)
for n in range(1, 7):
    run(ALL[:n])
run(ALL, 3)
run(ALL, 3)
run(ALL[::-1])
run([4, 4, 4, 4])
run([11, 4])
