"""Embedding backward (units path) by class of table: which columns of the bench configuration cost what."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from handyrec_b200 import _lib
from handyrec_b200._lib import call
from handyrec_b200.engine import DeepFMEngine
from handyrec_b200 import kernels as K

dev = torch.device("cuda", 0)
B = bench.BATCH
ALL = bench.CRITEO_VOCABS
classes = {
    "all26": ALL,
    "big9 (>131072 rows, row update)": [v for v in ALL if v > 131072],
    "mid7 (5k..100k rows, dense grads)": [v for v in ALL if 2100 < v <= 131072],
    "tiny10 (<=2001 rows, dense grads)": [v for v in ALL if v <= 2100],
    "small17": [v for v in ALL if v <= 131072],
}
for name, vocabs in classes.items():
    tabs = []
    for f, v in enumerate(vocabs):
        t = torch.empty(v, bench.EMB_DIM, device=dev)
        K.init_uniform(t, seed=7 + f)
        tabs.append(t)
    eng = DeepFMEngine(tabs, [(f, 1, "none") for f in range(len(vocabs))], bench.N_DENSE, bench.DNN_HIDDEN, "relu", batch_size=B, optimizer="adam",
                       dense_table_max_rows=131072)
    eng.autotune_embedding_bwd = False
    g = torch.Generator(device=dev).manual_seed(1)
    pool = []
    for _ in range(4):
        ids = torch.stack([torch.randint(0, v, (B,), device=dev, generator=g) for v in vocabs], 1).to(torch.int32).contiguous()
        pool.append((ids, torch.rand(B, bench.N_DENSE, device=dev, generator=g), (torch.rand(B, device=dev, generator=g) < 0.25).float()))
    out = []
    for algo, tag in ((_lib.BWD_UNITS, "units"), (_lib.BWD_SORT, "sort")):
        call("hrb_plan_set_bwd_algo", eng.plan._h, algo)
        for s in range(4):
            eng.train_step_on_device(*pool[s % 4])
        ph = {}
        for s in range(8):
            for k, v in eng.profile_step(*pool[s % 4]).items():
                ph[k] = ph.get(k, 0.0) + v / 8
        out.append(f"{tag}={ph['embedding_bwd_update']:.4f}")
    print(f"{name}: {len(vocabs)} columns, {len(vocabs) * B} pairs: " + " ".join(out) + " ms", flush=True)
    del eng, tabs, pool
    torch.cuda.empty_cache()
