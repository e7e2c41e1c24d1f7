"""Fabric check for the exchange step at N ranks (no compute running): (a) every rank pushes one 464 MB slot to every
peer with copy-engine DMAs over CUDA IPC (the bench's pattern, 1 / 4 / 7 streams), (b) NCCL all_gather_into_tensor of the
same volume.  Prints per-GPU outbound GB/s.  Run with torchrun --nproc-per-node N."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import shard
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = mr.api.Context(64, 48, local)
SLOT = 8 * 1920 * 1080 * 28
x = shard.PeerExchange(ctx, SLOT, dev)
cudart = C.CDLL("libcudart.so.12")
streams = [torch.cuda.Stream() for _ in range(8)]
REPS = 3
def push_all(ns):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    off = rank * x.slot_bytes
    for rep in range(REPS):
        for k in range(1, world):
            p = (rank + k) % world
            st = streams[(k - 1) % ns]
            cudart.cudaMemcpyAsync(C.c_void_p(x.peer[p] + off), C.c_void_p(x.ptr + off), C.c_size_t(SLOT), C.c_int(4), C.c_void_p(st.cuda_stream))
    torch.cuda.synchronize(); dist.barrier()
    return REPS * (world - 1) * SLOT / (time.perf_counter() - t0) / 1e9
res = {}
for ns in (1, 4, 7):
    push_all(ns)
    res[f"dma_push_streams{ns}"] = push_all(ns)
send = torch.empty(SLOT // 4, dtype=torch.float32, device=dev)
recv = torch.empty(world * (SLOT // 4), dtype=torch.float32, device=dev)
def nccl_ag():
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for rep in range(REPS):
        dist.all_gather_into_tensor(recv, send)
    torch.cuda.synchronize()
    return REPS * (world - 1) * SLOT / (time.perf_counter() - t0) / 1e9
nccl_ag()
res["nccl_all_gather"] = nccl_ag()
out = [None] * world
dist.all_gather_object(out, {k: round(v) for k, v in res.items()})
if rank == 0:
    print("per-GPU outbound GB/s, world", world)
    for r, o in enumerate(out):
        print(" rank", r, o, flush=True)
dist.barrier()
x.close()
dist.destroy_process_group()
