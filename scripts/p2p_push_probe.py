"""NVLink copy-engine push bandwidth between two ranks (CUDA IPC buffers of the library): one 464 MB slot pushed 8x per
round over 1..8 streams, both ranks pushing at the same time.  Run with torchrun --nproc-per-node 2."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import shard
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = mr.api.Context(64, 48, local)
SLOT = 8 * 1920 * 1080 * 28
x = shard.PeerExchange(ctx, SLOT, torch.device("cuda", local))
peer = x.peer[(rank + 1) % world]
cudart = C.CDLL("libcudart.so.12")
streams = [torch.cuda.Stream() for _ in range(8)]
def run(ns, chunks):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    per = SLOT // chunks // 256 * 256
    for rep in range(4):
        for c in range(chunks):
            st = streams[c % ns]
            off = rank * x.slot_bytes + c * per
            cudart.cudaMemcpyAsync(C.c_void_p(peer + off), C.c_void_p(x.ptr + off), C.c_size_t(per), C.c_int(4), C.c_void_p(st.cuda_stream))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return 4 * chunks * per / dt / 1e9
for ns, chunks in [(1, 1), (2, 2), (4, 4), (8, 8), (8, 32)]:
    run(ns, chunks)
    bw = run(ns, chunks)
    if rank == 0:
        print(f"streams={ns} chunks={chunks}: {bw:.0f} GB/s per direction (both ranks pushing)", flush=True)
dist.barrier()
x.close()
dist.destroy_process_group()
