"""Host packing throughput on this box (hrb_host_pack_*), per thread count."""
import sys, os, time, ctypes, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from handyrec_b200._lib import call
from handyrec_b200.lowering import _ColumnSet
n = 65536 * 8
g = np.random.RandomState(0)
ids = [g.randint(0, 1000, (n, 1)).astype(np.int32) for _ in range(26)]
dn = [g.rand(n, 1).astype(np.float32) for _ in range(13)]
ci, cd = _ColumnSet(ids), _ColumnSet(dn)
dst = torch.empty(65536, 26, dtype=torch.int32).pin_memory()
dd = torch.empty(65536, 13).pin_memory()
print("cores", os.cpu_count())
for nt in (1, 2, 4, 8, 16, 0):
    for rep in range(2):
        t0 = time.perf_counter()
        for k in range(8):
            call("hrb_host_pack_i32", ci.ptrs, ci.dtype, ci.width, ci.ld, ci.n, k * 65536, 65536, ctypes.c_void_p(dst.data_ptr()), 26, nt)
            call("hrb_host_pack_f32", cd.ptrs, cd.dtype, cd.width, cd.ld, cd.n, k * 65536, 65536, ctypes.c_void_p(dd.data_ptr()), 13, nt)
    print(nt, round((time.perf_counter() - t0) / 8 * 1e3, 3), "ms per batch")
