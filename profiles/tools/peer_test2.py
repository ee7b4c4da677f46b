import os, sys, ctypes, torch, torch.distributed as dist
sys.path.insert(0, '.')
import torch.distributed._symmetric_memory as symm_mem
from handyrec_b200 import kernels as K
from handyrec_b200._lib import call
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
V, D = 1000, 16
t = symm_mem.empty((V, D), dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
t.copy_(torch.full((V, D), float(rank + 1), device=dev) + torch.arange(V, device=dev).float().unsqueeze(1) * 1e-3)
torch.cuda.synchronize(); dist.barrier()
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "mine", hex(t.data_ptr()), flush=True)
other = hdl.get_buffer(1 - rank, (V, D), torch.float32)
print(rank, "peer view device", other.device, float(other[5, 0]), flush=True)
ids = torch.arange(0, 64, device=dev, dtype=torch.int32)
out, _ = K.embedding_fwd(other, ids, False, check_ids=False)
torch.cuda.synchronize()
print(rank, "ldg kernel ok", float(out[5, 0]), flush=True)
plan = K.LookupPlan([t], [(0, 1, "none", 0, 0)])
ptrs = (ctypes.c_void_p * 2)(*[int(p) for p in hdl.buffer_ptrs])
full = (ctypes.c_int64 * 1)(2 * V)
call("hrb_plan_set_peers", plan._h, 2, ptrs, full)
gid = torch.arange(0, 128, device=dev, dtype=torch.int32).reshape(-1, 1).contiguous()
res = plan.forward(gid)["out"]
torch.cuda.synchronize()
print(rank, "tile kernel ok", res[:4, 0].tolist(), flush=True)
dist.barrier(); dist.destroy_process_group()
