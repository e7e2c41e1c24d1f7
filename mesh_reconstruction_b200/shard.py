"""Multi-GPU host logic of the path: frame-pair sharding and the one exchange step.

The reference is single-process (recon.cpp:65-119 loops serially over main cameras); every main
frame is an independent unit, so ranks take CONTIGUOUS blocks of the sorted main-camera list
(heuristic.cpp:484 sorts it) and a variable-length all-gather of the M_r x 7 point rows, concatenated
in rank order, reproduces the reference's row order (recon.cpp:115-116 appends per main frame).
One process per GPU; NCCL over NVLink for CUDA tensors, gloo for the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_main_frames(n_main, world, rank):
    """Contiguous block [lo, hi) of main-frame indices for `rank`; blocks differ by at most one."""
    base, rem = divmod(n_main, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_points(rows, count=None, group=None):
    """All-gather of variable-length point rows.

    rows : (capacity, 7) float32 tensor on this rank (CUDA for NCCL, CPU for gloo); only the first
           `count` rows are meaningful (count defaults to rows.shape[0]).
    Returns (all_rows, counts): the rows of every rank concatenated in rank order -- i.e. the
    reference's append order -- and the per-rank counts.  Two collectives: counts (world x int64),
    then the rows padded to the largest count (so the volume is world x max_r M_r x 28 bytes, not
    world x capacity)."""
    world = dist.get_world_size(group)
    count = rows.shape[0] if count is None else int(count)
    dev = rows.device
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([count], dtype=torch.int64, device=dev), group=group)
    counts_h = counts.cpu().tolist()
    mx = max(max(counts_h), 1)
    if rows.shape[0] >= mx:
        send = rows[:mx]
    else:
        send = torch.zeros((mx, rows.shape[1]), dtype=rows.dtype, device=dev)
        send[:count] = rows[:count]
    recv = torch.empty((world * mx, rows.shape[1]), dtype=rows.dtype, device=dev)   # concatenated layout (gloo + nccl)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    recv = recv.view(world, mx, rows.shape[1])
    out = torch.cat([recv[r, :counts_h[r]] for r in range(world)], 0)
    return out, counts_h
