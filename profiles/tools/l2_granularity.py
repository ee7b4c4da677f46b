"""Does the context's L2 fetch granularity (CU_LIMIT_MAX_L2_FETCH_GRANULARITY, 32..128 B) change the cost of the random 64-byte row
gathers?  ncu shows about twice the algorithmic DRAM read bytes on lookup_tile_kernel and bwd_unit_kernel (r02_ncu_full_summary.md)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from handyrec_b200 import _lib
from handyrec_b200._lib import call
from handyrec_b200.engine import DeepFMEngine
from handyrec_b200 import kernels as K

cu = ctypes.CDLL("libcuda.so.1")
LIM = 0x05  # CU_LIMIT_MAX_L2_FETCH_GRANULARITY
dev = torch.device("cuda", 0)
torch.zeros(1, device=dev)
vocabs = bench.CRITEO_VOCABS
B = bench.BATCH
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


def measure(tag):
    cur = ctypes.c_size_t(0)
    cu.cuCtxGetLimit(ctypes.byref(cur), LIM)
    for s in range(6):
        eng.train_step_on_device(*pool[s % 4])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for s in range(20):
        eng.plan.forward(pool[s % 4][0], out=eng.X0, fm=(eng.fm_w, eng.fm_w0), want_fm_sum=False)
    b.record()
    torch.cuda.synchronize()
    lk = a.elapsed_time(b) / 20
    a.record()
    for s in range(20):
        eng.train_step_on_device(*pool[s % 4])
    b.record()
    torch.cuda.synchronize()
    step = a.elapsed_time(b) / 20
    ph = {}
    for s in range(8):
        for k, v in eng.profile_step(*pool[s % 4]).items():
            ph[k] = ph.get(k, 0.0) + v / 8
    print(f"{tag}: limit={cur.value} lookup_alone={lk:.4f} ms step={step:.4f} ms lookup_in_step={ph.get('lookup_fm_fwd', 0):.4f} "
          f"emb_bwd={sum(v for k, v in ph.items() if 'embedding' in k):.4f}", flush=True)


measure("default")
for gran in (32, 64, 128):
    rc = cu.cuCtxSetLimit(LIM, ctypes.c_size_t(gran))
    measure(f"set {gran} (rc={rc})")
