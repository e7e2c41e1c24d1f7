"""world_size-2 gloo tests (CPU) of the N>1 host logic: contiguous sharding of main frames and
the variable-length all-gather that must reproduce the reference's row order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mesh_reconstruction_b200.shard import allgather_points, shard_main_frames


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rows_for_frame(i):
    rng = np.random.default_rng(1000 + i)
    m = int(rng.integers(0, 40))            # ragged, may be empty
    return rng.normal(size=(m, 7)).astype(np.float32)


def _worker(rank, world, port, n_main, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_main_frames(n_main, world, rank)
    mine = [_rows_for_frame(i) for i in range(lo, hi)]
    rows = np.concatenate(mine, 0) if mine else np.zeros((0, 7), np.float32)
    cap = torch.zeros((rows.shape[0] + 5, 7))          # capacity larger than the count, like the device buffer
    cap[:len(rows)] = torch.from_numpy(rows)
    out, counts = allgather_points(cap, len(rows))
    if rank == 0:
        q.put((out.numpy(), counts))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_blocks_cover_in_order():
    for n in (0, 1, 7, 299, 600):
        for world in (1, 2, 3, 8):
            blocks = [shard_main_frames(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_allgather_reproduces_serial_row_order():
    world, n_main = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_main, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, counts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    serial = np.concatenate([_rows_for_frame(i) for i in range(n_main)], 0)
    assert sum(counts) == len(serial)
    assert np.array_equal(out, serial)
