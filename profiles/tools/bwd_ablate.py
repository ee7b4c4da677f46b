"""Embedding backward time of the bench configuration under HRB_BWD_DEBUG ablation bits (needs a HRB_DEVTOOLS=1 build; one process per value)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from handyrec_b200 import _lib
from handyrec_b200._lib import call
from handyrec_b200.engine import DeepFMEngine
from handyrec_b200 import kernels as K

dev = torch.device("cuda", 0)
vocabs, B = bench.CRITEO_VOCABS, bench.BATCH
tabs = []
for f, v in enumerate(vocabs):
    t = torch.empty(v, bench.EMB_DIM, device=dev)
    K.init_uniform(t, seed=7 + f)
    tabs.append(t)
eng = DeepFMEngine(tabs, [(f, 1, "none") for f in range(len(vocabs))], bench.N_DENSE, bench.DNN_HIDDEN, "relu", batch_size=B, optimizer="adam",
                   dense_table_max_rows=131072)
eng.autotune_embedding_bwd = False
call("hrb_plan_set_bwd_algo", eng.plan._h, _lib.BWD_UNITS)
g = torch.Generator(device=dev).manual_seed(1)
pool = []
for _ in range(4):
    ids = torch.stack([torch.randint(0, v, (B,), device=dev, generator=g) for v in vocabs], 1).to(torch.int32).contiguous()
    pool.append((ids, torch.rand(B, bench.N_DENSE, device=dev, generator=g), (torch.rand(B, device=dev, generator=g) < 0.25).float()))
for s in range(6):
    eng.train_step_on_device(*pool[s % 4])
ph = {}
for s in range(12):
    for k, v in eng.profile_step(*pool[s % 4]).items():
        ph[k] = ph.get(k, 0.0) + v / 12
print(f"HRB_BWD_DEBUG={os.environ.get('HRB_BWD_DEBUG', '0')}: embedding_bwd_update {ph['embedding_bwd_update']:.4f} ms", flush=True)
