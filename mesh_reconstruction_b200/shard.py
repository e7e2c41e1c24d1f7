"""Multi-GPU host logic of the path: frame-pair sharding and the one exchange step.

The reference is single-process (recon.cpp:65-119 loops serially over main cameras); every main
frame is an independent unit, so ranks take CONTIGUOUS blocks of the sorted main-camera list
(heuristic.cpp:484 sorts it) and a variable-length all-gather of the M_r x 7 point rows, concatenated
in rank order, reproduces the reference's row order (recon.cpp:115-116 appends per main frame).
One process per GPU; NCCL over NVLink for CUDA tensors, gloo for the CPU tests.
"""
import os

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


# ---- the same exchange through the C ABI (mr_allgather_points) -------------------------------------------
# A C/C++ host creates its ncclComm_t itself; from Python the helpers below build one with the NCCL copy that is
# already mapped into the process (torch's), using torch.distributed only to pass the 128-byte unique id around.

def _nccl_lib():
    import ctypes as C
    path = "libnccl.so.2"
    try:
        with open("/proc/self/maps") as f:
            for line in f:
                if "libnccl.so" in line:
                    path = line.split()[-1]
                    break
    except OSError:
        pass
    return C.CDLL(path, mode=C.RTLD_GLOBAL)   # global: libmeshrecon_b200.so resolves NCCL from the process first


def raw_nccl_comm(device, group=None):
    """ncclComm_t (as an int) spanning the ranks of `group`, for mr_allgather_points.  Collective."""
    import ctypes as C

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_char * 128)]

    nccl = _nccl_lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = UniqueId()
    if rank == 0:
        rc = nccl.ncclGetUniqueId(C.byref(uid))
        if rc != 0:
            raise RuntimeError(f"ncclGetUniqueId failed ({rc})")
    box = [C.string_at(C.byref(uid), 128) if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    C.memmove(C.byref(uid), box[0], 128)
    torch.cuda.set_device(device)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    rc = nccl.ncclCommInitRank(C.byref(comm), world, uid, rank)
    if rc != 0:
        raise RuntimeError(f"ncclCommInitRank failed ({rc})")
    return comm.value


def destroy_raw_nccl_comm(comm):
    import ctypes as C
    nccl = _nccl_lib()
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(C.c_void_p(comm))


def allgather_points_cabi(ctx, comm, rows, count, out=None):
    """Variable-length all-gather of CUDA point rows through ``mr_allgather_points`` (exact counts, rank-order
    concatenation).  ctx: api.Context of this rank's GPU; comm: raw ncclComm_t.  Returns (all_rows, counts)."""
    import ctypes as C
    # the communicator's own size (not torch.distributed's: the C side writes one count per NCCL rank)
    nccl = _nccl_lib()
    nccl.ncclCommCount.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    n = C.c_int(0)
    if nccl.ncclCommCount(C.c_void_p(comm), C.byref(n)) != 0 or n.value < 1:
        raise RuntimeError("ncclCommCount failed on the given communicator")
    world = n.value
    counts = (C.c_int * world)()
    total = C.c_longlong(0)
    if out is None:
        cap = torch.tensor([int(count)], dtype=torch.int64, device=rows.device)
        if world > 1:
            dist.all_reduce(cap)                      # capacity = exact total (one tiny collective; a C host sizes for the worst case)
        out = torch.empty((max(int(cap.item()), 1), 7), dtype=torch.float32, device=rows.device)
    ctx.check(ctx.lib.mr_allgather_points(ctx.h, C.c_void_p(comm), C.c_void_p(rows.data_ptr()), int(count), C.c_void_p(out.data_ptr()),
                                          out.shape[0], counts, C.byref(total)))
    ctx.synchronize()
    return out[:total.value], [int(c) for c in counts][:world]


# ---- kernel-free exchange over NVLink peer memory (mr_xchg_*) --------------------------------------------------

class _DevMem:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerExchange:
    """One receive buffer per rank with a slot for every rank, mapped into every peer over CUDA IPC.

    ``push()`` DMAs this rank's own slot into the same slot of every peer's buffer (copy engines over NVLink, no
    SMs), ordered after the work queued on the context's stream; ``signal_stream()`` is the stream on which to order
    the completion barrier (e.g. an async 4-byte all-reduce): when it has completed on every rank, every slot of
    every buffer has landed.  Slots are ``slot_bytes`` each; ``slot(r, ...)`` views slot r of the LOCAL buffer as a
    tensor -- pass ``slot(rank, ...)`` as the output buffer of the path so the rows are produced in place."""

    def __init__(self, ctx, slot_bytes, device, group=None):
        import ctypes as C
        self.ctx, self.lib, self.group = ctx, ctx.lib, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.slot_bytes = (int(slot_bytes) + 255) // 256 * 256
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        # Collective and failure-consistent: every rank reaches every collective below whatever fails locally, and
        # either all ranks end up with a working exchange or all of them raise.
        ptr, handle, err = C.c_void_p(), C.create_string_buffer(64), None
        self.ptr, self.peer = 0, {}
        if self.lib.mr_xchg_alloc(ctx.h, self.slot_bytes * self.world, C.byref(ptr), handle) == 0:
            self.ptr = ptr.value
        else:
            err = self.lib.mr_last_error(ctx.h).decode()
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw if self.ptr else None, group=group)
        if err is None and any(h is None for h in handles):
            err = "a peer could not allocate its exchange buffer"
        if err is None:
            for p in range(self.world):
                if p == self.rank:
                    continue
                pp = C.c_void_p()
                if self.lib.mr_xchg_open(ctx.h, handles[p], C.byref(pp)) != 0:
                    err = self.lib.mr_last_error(ctx.h).decode()
                    break
                self.peer[p] = pp.value
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError("peer-memory exchange unavailable: " + (err or "failed on another rank"))

    def slot(self, r, shape, dtype=torch.float32, offset_bytes=0):
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        return torch.as_tensor(_DevMem(self.ptr + r * self.slot_bytes + offset_bytes, shape, typestr), device=self.device)

    def push(self, nbytes=None):
        import ctypes as C
        nbytes = self.slot_bytes if nbytes is None else int(nbytes)
        off = self.rank * self.slot_bytes
        for k in range(1, self.world):                       # staggered: rank r starts with peer r+1, so no peer is hit by all at once
            p = (self.rank + k) % self.world
            self.ctx.check(self.lib.mr_xchg_push(self.ctx.h, C.c_void_p(self.peer[p] + off), C.c_void_p(self.ptr + off), nbytes))

    def signal_stream(self):
        return torch.cuda.ExternalStream(self.lib.mr_xchg_stream(self.ctx.h), device=self.device)

    def close(self):
        import ctypes as C
        for p, ptr in self.peer.items():
            self.lib.mr_xchg_close(self.ctx.h, C.c_void_p(ptr))
        self.peer = {}
        if dist.is_initialized():
            dist.barrier(group=self.group)                   # nobody frees a buffer a peer still has mapped
        if self.ptr:
            self.lib.mr_xchg_free(self.ctx.h, C.c_void_p(self.ptr))
            self.ptr = 0


class McastExchange:
    """The same exchange with ONE push per rank and step: every rank's receive buffer is bound to an NVSwitch multicast
    object, and a copy-engine DMA to the multicast address lands in the same slot of EVERY rank's buffer (the switch
    replicates the writes), so a GPU sends its rows once instead of world - 1 times (at 8 GPUs the seven
    unicast pushes of 464 MB per step are what the copy engines cannot sustain next to the path's kernels).

    The symmetric allocation, the exchange of the shareable handles between the processes and the multicast binding are
    torch.distributed's symmetric memory (plumbing: cuMemCreate / cuMulticastCreate / cuMulticastBindMem under the hood);
    the DMA itself is ``mr_xchg_push`` on the library's push streams, ordered after the path's kernels, exactly like
    :class:`PeerExchange`.  The rows are produced into ``send`` (a plain device buffer: a multicast write also lands in the
    sender's own slot, and source and destination of a DMA must not alias).  Raises on every rank if the fabric offers
    no multicast."""

    def __init__(self, ctx, slot_bytes, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.ctx, self.lib, self.group = ctx, ctx.lib, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.slot_bytes = (int(slot_bytes) + 255) // 256 * 256
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        err, self.buf, self.hdl, self.mc = None, None, None, 0
        self.chunks = int(os.environ.get("MR_MCAST_CHUNKS", "4"))
        self.mode = os.environ.get("MR_MCAST_MODE", "sm")        # "sm": multimem.st kernel, "ce": copy engines
        try:
            self.buf = symm.empty(self.slot_bytes * self.world, dtype=torch.uint8, device=self.device)
            self.hdl = symm.rendezvous(self.buf, group=(group or dist.group.WORLD))
            self.mc = int(self.hdl.multicast_ptr or 0)
            if not self.mc:
                err = "no multicast pointer (NVSwitch multicast unsupported on this fabric / driver)"
        except Exception as e:  # noqa: BLE001
            err = f"{type(e).__name__}: {e}"
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            raise RuntimeError("multicast exchange unavailable: " + (err or "failed on another rank"))
        self.ptr = self.buf.data_ptr()
        self.send = torch.zeros(self.slot_bytes, dtype=torch.uint8, device=self.device)

    def slot(self, r, shape, dtype=torch.float32, offset_bytes=0):
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        return torch.as_tensor(_DevMem(self.ptr + r * self.slot_bytes + offset_bytes, shape, typestr), device=self.device)

    def out(self, shape, dtype=torch.float32, offset_bytes=0):
        """View of the send buffer: hand it to the path as its output buffer."""
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        return torch.as_tensor(_DevMem(self.send.data_ptr() + offset_bytes, shape, typestr), device=self.device)

    def push(self, nbytes=None):
        import ctypes as C
        nbytes = self.slot_bytes if nbytes is None else int(nbytes)
        dst, src = self.mc + self.rank * self.slot_bytes, self.send.data_ptr()
        if self.mode == "sm":
            # a few CTAs of 128-bit multimem.st on the high-priority push stream (mr_xchg_push_mcast)
            self.ctx.check(self.lib.mr_xchg_push_mcast(self.ctx.h, C.c_void_p(dst), C.c_void_p(src), (nbytes + 15) // 16 * 16))
            return
        # copy engines: one reaches ~100 GB/s into a multicast address, two or more ~330 GB/s on idle GPUs
        # (scripts/mcast_push_probe.py) -- but only ~35 GB/s while the path's kernels run
        step = (nbytes // self.chunks + 255) // 256 * 256
        for o in range(0, nbytes, step):
            self.ctx.check(self.lib.mr_xchg_push(self.ctx.h, C.c_void_p(dst + o), C.c_void_p(src + o), min(step, nbytes - o)))

    def signal_stream(self):
        return torch.cuda.ExternalStream(self.lib.mr_xchg_stream(self.ctx.h), device=self.device)

    def close(self):
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier(group=self.group)
        self.hdl = self.buf = None
