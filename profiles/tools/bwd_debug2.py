import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from handyrec_b200 import kernels as k
dev = torch.device("cuda:0")
B, D = 65536, 16
def run(vocabs, every=7):
    tables = [torch.zeros(v, D) for v in vocabs]
    ids = torch.stack([torch.full((B,), min(3, v - 1), dtype=torch.int32) for v in vocabs], 1)
    if every:
        ids[::every, len(vocabs) - 1] = 123456
    dout = torch.zeros(B, len(vocabs) * D)
    b = torch.arange(B)
    for f in range(len(vocabs)):
        dout[:, f * D + 0] = 1
        dout[:, f * D + 1] = (b % 256).float()
        dout[:, f * D + 2] = (b // 256).float()
    dt = [t.clone().to(dev) for t in tables]
    plan = k.LookupPlan(dt, [(f, 1, "none", f, f * D) for f in range(len(vocabs))])
    plan.backward_update(ids.contiguous().to(dev), dout.to(dev), opt="sgd", lr=1.0)
    torch.cuda.synchronize()
    got = -dt[0].cpu()[min(3, vocabs[0] - 1), :3].double()
    want = torch.tensor([B, float((b % 256).sum()), float((b // 256).sum())]).double()
    print(vocabs, every, "got", got.tolist(), "deficit", (want - got).tolist(), flush=True)
run([4, 2000000])
run([4, 2000000], every=0)
run([4, 2000000], every=3)
run([4, 11, 300, 5000, 250000, 2000000])
run([4, 300000])
