"""Copy-engine pushes to an NVSwitch multicast address vs unicast peer pushes, idle GPUs (torchrun, N >= 2):
GB/s of payload leaving each GPU and time until it has landed everywhere, for 1 / 2 / 4 / 8 chunks on as many streams."""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 464 * 1024 * 1024
buf = symm.empty(nbytes * world, dtype=torch.uint8, device=dev)
hdl = symm.rendezvous(buf, group=dist.group.WORLD)
mc = int(hdl.multicast_ptr or 0)
peers = [int(p) for p in hdl.buffer_ptrs]
src = torch.full((nbytes,), rank + 1, dtype=torch.uint8, device=dev)
from cuda.bindings import runtime as cudart  # noqa: E402
D2D = cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice
streams = [torch.cuda.Stream(device=dev) for _ in range(8)]
flag = torch.zeros(1, dtype=torch.int32, device=dev)


def push(dst_base, nchunks):
    step = (nbytes // nchunks + 255) // 256 * 256
    for c in range(nchunks):
        o = c * step
        n = min(step, nbytes - o)
        if n <= 0:
            break
        with torch.cuda.stream(streams[c % len(streams)]):
            err, = cudart.cudaMemcpyAsync(dst_base + o, src.data_ptr() + o, n, D2D, streams[c % len(streams)].cuda_stream)
            assert err == cudart.cudaError_t.cudaSuccess, err


def timed(fn, reps=5):
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        for s in streams:
            s.synchronize()
    dist.all_reduce(flag)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


if rank == 0:
    print(f"# world {world}, {nbytes / 1e6:.0f} MB per rank; multicast ptr {'yes' if mc else 'NO'}", flush=True)
for nch in (1, 2, 4, 8):
    if mc:
        t = timed(lambda: push(mc + rank * nbytes, nch))
        ok = all(int(buf[r * nbytes + 12345]) == r + 1 for r in range(world))
        if rank == 0:
            print(f"multicast  chunks {nch}: {t * 1e3:7.2f} ms  ({nbytes / t / 1e9:6.1f} GB/s payload per GPU, {nbytes * (world - 1) / t / 1e9:6.1f} GB/s delivered to peers)  content ok {ok}", flush=True)
    def uni():
        for k in range(1, world):
            p = (rank + k) % world
            with torch.cuda.stream(streams[k % nch]):
                cudart.cudaMemcpyAsync(peers[p] + rank * nbytes, src.data_ptr(), nbytes, D2D, streams[k % nch].cuda_stream)
    t = timed(uni)
    if rank == 0:
        print(f"unicast    streams {nch}: {t * 1e3:7.2f} ms  ({nbytes * (world - 1) / t / 1e9:6.1f} GB/s leaving each GPU)", flush=True)
dist.barrier()
dist.destroy_process_group()
