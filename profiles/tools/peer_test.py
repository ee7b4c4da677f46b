import os, sys, ctypes, torch, torch.distributed as dist
sys.path.insert(0, '.')
from handyrec_b200 import kernels as K
from handyrec_b200._lib import call
from handyrec_b200.sharded import TorchDistComm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
comm = TorchDistComm()
V, D = 1000, 16
t = torch.full((V, D), float(rank + 1), device=dev) + torch.arange(V, device=dev).float().unsqueeze(1) * 1e-3
peers = comm.share_tables([t])
other = peers[1 - rank][0]
print(rank, "peer tensor device", other.device, "ptr", hex(other.data_ptr()), flush=True)
# (1) torch op on the peer tensor (runs in the peer device's context)
print(rank, "torch read", float(other[5, 0]), flush=True)
# (2) my LDG kernel on this device reading peer memory
ids = torch.arange(0, 64, device=dev, dtype=torch.int32)
out, _ = K.embedding_fwd(other, ids, False, check_ids=False)
torch.cuda.synchronize()
print(rank, "ldg kernel ok", float(out[5, 0]), flush=True)
# (3) tile kernel (cp.async) through a peer plan
plan = K.LookupPlan([t], [(0, 1, "none", 0, 0)])
ptrs = (ctypes.c_void_p * 2)(peers[0][0].data_ptr(), peers[1][0].data_ptr())
full = (ctypes.c_int64 * 1)(2 * V)
call("hrb_plan_set_peers", plan._h, 2, ptrs, full)
gid = torch.arange(0, 128, device=dev, dtype=torch.int32).reshape(-1, 1).contiguous()
res = plan.forward(gid)["out"]
torch.cuda.synchronize()
print(rank, "tile kernel ok", res[:4, 0].tolist(), flush=True)
dist.barrier(); dist.destroy_process_group()
