"""mr_allgather_points (C ABI, NCCL): the path's one exchange step.

* world = 1 communicator in-process: runs on the single-GPU tier (plumbing: run-time NCCL resolution, counts, offsets);
* world = 2 through torch.distributed.run, one rank per GPU: skipped with fewer than two GPUs.  The checker is
  shard.allgather_points (torch.distributed), itself covered on CPU with gloo in test_shard_gloo.py.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import shard
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = mr.api.Context(64, 48, local)
comm = shard.raw_nccl_comm(local)
g = torch.Generator().manual_seed(100 + rank)
for trial, count in enumerate([1000 + 37 * rank, 0 if rank == 0 else 5, 123456 * (rank + 1)]):
    rows = torch.rand((count + 3, 7), generator=g).cuda()          # capacity > count: only the first `count` rows travel
    got, counts = shard.allgather_points_cabi(ctx, comm, rows, count)
    ref, ref_counts = shard.allgather_points(rows, count)
    assert counts == ref_counts, (counts, ref_counts)
    assert got.shape == ref.shape and torch.equal(got, ref), trial
# kernel-free exchange: slots pushed into the peers' buffers over NVLink (CUDA IPC), completion = tiny all-reduce
SLOT_ROWS = 50000
x = shard.PeerExchange(ctx, SLOT_ROWS * 28 + 4, torch.device("cuda", local))
mine = x.slot(rank, (SLOT_ROWS, 7))
cnt = x.slot(rank, (1,), torch.int32, offset_bytes=SLOT_ROWS * 28)
flag = torch.zeros(1, dtype=torch.int32, device="cuda")
for trial in range(3):
    src = torch.rand((SLOT_ROWS, 7), generator=g).cuda() + trial
    with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream)):       # producer work on the library stream
        mine.copy_(src)
        cnt.fill_(1000 * trial + rank)
    x.push()
    with torch.cuda.stream(x.signal_stream()):
        dist.all_reduce(flag)
    torch.cuda.synchronize()
    ref, _ = shard.allgather_points(src, SLOT_ROWS)
    for p in range(world):
        assert torch.equal(x.slot(p, (SLOT_ROWS, 7)), ref[p * SLOT_ROWS:(p + 1) * SLOT_ROWS]), (trial, p)
        assert int(x.slot(p, (1,), torch.int32, offset_bytes=SLOT_ROWS * 28)) == 1000 * trial + p
    dist.barrier()                                                       # nobody overwrites a slot a peer is still checking
x.close()
# the same exchange through an NVSwitch multicast mapping: ONE push per rank (multimem.st CTAs, then the copy-engine
# variant) lands in every rank's buffer, the sender's own slot included
if world > 1:
    for mode in ("sm", "ce"):
        os.environ["MR_MCAST_MODE"] = mode
        try:
            m = shard.McastExchange(ctx, SLOT_ROWS * 28 + 4, torch.device("cuda", local))
        except RuntimeError as e:                                            # no NVSwitch multicast on this box: collective, every rank lands here
            print("RANK", rank, "multicast unavailable:", e)
            break
        out = m.out((SLOT_ROWS, 7))
        cnt = m.out((1,), torch.int32, offset_bytes=SLOT_ROWS * 28)
        for trial in range(3):
            src = torch.rand((SLOT_ROWS, 7), generator=g).cuda() + trial
            with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream)):
                out.copy_(src)
                cnt.fill_(7000 * trial + rank)
            m.push()
            with torch.cuda.stream(m.signal_stream()):
                dist.all_reduce(flag)
            torch.cuda.synchronize()
            ref, _ = shard.allgather_points(src, SLOT_ROWS)
            for p in range(world):
                assert torch.equal(m.slot(p, (SLOT_ROWS, 7)), ref[p * SLOT_ROWS:(p + 1) * SLOT_ROWS]), (mode, trial, p)
                assert int(m.slot(p, (1,), torch.int32, offset_bytes=SLOT_ROWS * 28)) == 7000 * trial + p
            dist.barrier()
        m.close()
        print("RANK", rank, "multicast", mode, "OK")
# an output buffer that is too small on ANY rank is refused on EVERY rank (the abort is collective: nobody is left
# waiting inside the row broadcasts), and so is a bad argument on one rank only
import ctypes as C
small = torch.empty((1, 7), device="cuda")
big = torch.empty((10 * world, 7), device="cuda")
mine_out = small if rank == world - 1 else big
rc = ctx.lib.mr_allgather_points(ctx.h, C.c_void_p(comm), C.c_void_p(rows.data_ptr()), 10, C.c_void_p(mine_out.data_ptr()), mine_out.shape[0], None, None)
assert rc == -1, rc
host = np.zeros((4, 7), np.float32)
src = host.ctypes.data_as(C.c_void_p) if rank == 0 else C.c_void_p(rows.data_ptr())       # rank 0 passes HOST rows
rc = ctx.lib.mr_allgather_points(ctx.h, C.c_void_p(comm), src, 4, C.c_void_p(big.data_ptr()), big.shape[0], None, None)
assert rc == -1, rc
if rank == 0:
    assert b"device memory" in ctx.lib.mr_last_error(ctx.h)
tot = C.c_longlong(-1)
rc = ctx.lib.mr_allgather_points(ctx.h, C.c_void_p(comm), C.c_void_p(rows.data_ptr()), 10, C.c_void_p(big.data_ptr()), big.shape[0], None, C.byref(tot))
assert rc == 0 and tot.value == 10 * world, (rc, tot.value)
ctx.synchronize()
dist.barrier()
shard.destroy_raw_nccl_comm(comm)
dist.destroy_process_group()
print("RANK", rank, "OK")
"""


def _run(world):
    script = os.path.join(ROOT, "tests", "_exchange_worker_tmp.py")
    with open(script, "w") as f:
        f.write(WORKER.format(root=ROOT))
    try:
        port = 29600 + (os.getpid() % 300)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), script]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
        for r in range(world):
            assert f"RANK {r} OK" in p.stdout
    finally:
        os.remove(script)


def test_allgather_points_cabi_world1():
    _run(1)


def test_allgather_points_cabi_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2)


def test_allgather_points_argument_errors():
    import ctypes as C
    import torch
    import mesh_reconstruction_b200 as mr
    ctx = mr.api.Context(64, 48)
    rows = torch.zeros((4, 7), device="cuda")
    assert ctx.lib.mr_allgather_points(ctx.h, None, C.c_void_p(rows.data_ptr()), 4, C.c_void_p(rows.data_ptr()), 4, None, None) == -1   # no communicator
    assert b"communicator" in ctx.lib.mr_last_error(ctx.h)
    # (host rows / too small buffers are refused collectively: covered by the worker script, which owns a communicator)
