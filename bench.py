#!/usr/bin/env python
"""bench.py -- headline benchmark of the dense-correspondence hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU arithmetic (oracle)

Metric: Mpix/s matched + triangulated = pixels of all processed (main, side) pairs / time.
Workload (config 4 of BASELINE.json, the headline): synthetic 1920x1080 300-frame sequence,
adjacent-pair matching (main i, side i+1, S = 1); `--config 5`: synthetic 3840x2160 600-frame
sequence, multi-baseline (S = 4, sides i-2, i-1, i+1, i+2).  One "step" = `--pairs` main frames
per GPU (S pairs each) through the whole path (depth raster -> shadow raster + dilation -> reproject + mixBackground -> variational
refinement -> cubic remap -> pyramid compare -> Newton triangulation -> PCA normals), plus, for
N > 1, the NCCL all-gather of the point rows.  Frame pairs shard across ranks (weak scaling).

* `value`  : frames already resident in HBM when the timed region starts, point rows left in HBM.
* `e2e`    : same call (`mr_submit_main_frame` through the ctypes binding) with HOST buffers: frames
             in pinned host memory (H2D inside the timed region) and every pair's point rows + count
             copied back to pinned host memory (D2H inside the timed region; the host reads step s-1's
             results while step s runs, the last step is drained before the region ends).
* `roofline`: dominant kernel stage measured live with CUDA events on the library's stream.
* `cpu_baseline`: the CPU oracle (cv2 for the OpenCV-owned arithmetic + C restatement) on a
             bounded sample of the same workload, on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PATH = lambda S: 33 + 21 * S          # per main-frame pixel  # noqa: E731
STAGE_ALG_BYTES = {                              # per pixel-pair, DESIGN.md "Kernels"
    "raster": 4 + 4, "shade_mix": 1 + 4 + 1 + 4 + 1, "variational_refinement": 1 + 1 + 8, "cubic_remap": 8 + 1 + 1,
    "pyramid_compare": 1 + 1 + 4, "triangulate": 16 + 4 + 20, "normals": 20 + 28,
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons (NVML, every few ms) while the timed region runs.  NVML is initialised in the
    constructor and the thread is started BEFORE the warm-up (a cold nvmlInit can take longer than a short timed region);
    samples are only kept while `recording` is set."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.recording, self.sm, self.reasons, self.max_sm = index, False, False, [], 0, None
        self.nv = self.h = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def sample(self):
        nv, h = self.nv, self.h
        self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        try:
            self.reasons |= nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            self.reasons |= nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)

    def run(self):
        if self.nv is None:
            return
        try:
            while not self.stop_flag:
                if self.recording:
                    self.sample()
                time.sleep(0.003)
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["unavailable: " + getattr(self, "err", "no samples")]}
        sm = sorted(self.sm)
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_sm,
                "reasons": [n for b, n in bits.items() if self.reasons & b], "samples": len(sm)}


def sides_of(a, S):
    """side frames of main frame a: a+1 (S = 1, config 4), a-1 / a+1 (S = 2), a-2, a-1, a+1, a+2 (S = 4, config 5)"""
    return [a + 1] if S == 1 else ([a - 1, a + 1] if S == 2 else [a - 2, a - 1, a + 1, a + 2])


def main_range(n_local, S):
    """main frames of a block of n_local frames whose side frames all lie inside the block"""
    h = S // 2
    return (h, n_local - h) if S > 1 else (0, n_local - 1)


def bind_to_gpu_numa(local, diag):
    """Pin this rank to the CPUs of its GPU's NUMA node BEFORE any pinned buffer is allocated: first-touch then places
    the row / frame buffers in the memory next to the GPU's PCIe root, so that the D2H row traffic of the ranks does not
    cross the socket interconnect (e2e flattens at ~89 GB/s aggregate from 2 GPUs on)."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(local)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]                      # sysfs uses a 4-digit PCI domain
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read().strip())
        cpus = []
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if part:
                a, _, b = part.partition("-")
                cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        diag["numa_node"] = node
        if node >= 0 and allowed:
            os.sched_setaffinity(0, allowed)
            diag["numa_bound_cpus"] = len(allowed)
    except Exception as e:  # noqa: BLE001
        diag["numa_bind_error"] = str(e)[:120]


def cpu_reference_pairs(scene, frames, pairs, threads, farneback=False):
    """Times the CPU oracle (reference arithmetic) on the given (main, [sides]) jobs (a bare side index = one side)."""
    import cv2
    from oracle.pipeline import process_main_frame
    from oracle.render import RenderOracle
    cv2.setNumThreads(threads)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    r = RenderOracle(scene.width, scene.height)
    r.loadMesh(scene.vertices, scene.faces)
    t0 = time.perf_counter()
    pts = 0
    for fa, fb in pairs:
        tri = process_main_frame(r, frames, scene.cameras, fa, list(fb) if isinstance(fb, (list, tuple)) else [fb], use_farneback=farneback)
        pts += len(tri)
    return time.perf_counter() - t0, pts


def workload_name(args):
    W, H, S = args.width, args.height, args.sides
    if args.config == 5:
        return f"synthetic {W}x{H} {args.frames}-frame sequence, multi-baseline matching (S={S}: sides i-2, i-1, i+1, i+2) + triangulation (BASELINE config 5); surface height z0 = {args.scene_z0:g}"
    return f"synthetic {W}x{H} {args.frames}-frame sequence, adjacent-pair matching + triangulation, S={S} (BASELINE config 4); surface height z0 = {args.scene_z0:g}"


def run_reference(args):
    """--impl reference: the reference's own CPU arithmetic for the path.  The upstream binary
    cannot be built here (OpenCV C++/GLX/CGAL absent), so this is the oracle port: the real OpenCV
    (cv2) for VariationalRefinement/remap/pyramids/Sobel + oracle/recon_oracle.c, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mesh_reconstruction_b200 import synth
    W, H = args.width, args.height
    cores = os.cpu_count() or 1
    S = args.sides
    nfr = 4 + 2 * (S // 2) if S > 1 else 4
    scene = synth.make_scene(W, H, nfr, step=args.cam_step, mesh_err=args.mesh_err, z0=args.scene_z0)
    frames = {i: scene.frame(i) for i in range(nfr)}
    lo, hi = main_range(nfr, S)
    pairs_cycle = [(a, sides_of(a, S)) for a in range(lo, hi)]
    if args.warmup > 0:      # one real warm-up main frame is enough (each is seconds of CPU work)
        cpu_reference_pairs(scene, frames, [pairs_cycle[0]], cores, args.farneback)
    t_total = 0.0
    for k in range(args.steps):
        t, _ = cpu_reference_pairs(scene, frames, [pairs_cycle[k % len(pairs_cycle)]], cores, args.farneback)
        t_total += t
    pix = args.steps * W * H * S
    val = pix / t_total / 1e6
    line = {
        "impl": "reference", "metric": "Mpix/s matched+triangulated", "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "main_frames_per_step": 1, "side_frames": S,
                   "sample": f"1 main frame ({S} frame pair{'s' if S > 1 else ''}) per step (bounded sample of the {args.frames}-frame workload)"},
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} main frames x {S} side frames of {W}x{H}, cv2 {__import__('cv2').__version__} with {cores} threads + C restatement (OpenMP)"},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[4, 5],
                    help="BASELINE.json config: 4 = synthetic 1920x1080 x300, adjacent pairs (S = 1, the headline); "
                         "5 = synthetic 3840x2160 x600, multi-baseline (S = 4, side frames i-2, i-1, i+1, i+2)")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--sides", type=int, default=None, choices=[1, 2, 4], help="side cameras per main frame (default: 1 for config 4, 4 for config 5)")
    ap.add_argument("--pairs", type=int, default=None, help="MAIN frames per GPU per step (each with S side frames); default 8 (config 4) / 4 (config 5)")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--no-filter-bench", action="store_true", help="skip the filterPoints measurement (N = 1, after the timed regions)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin this rank (and its pinned host buffers) to the CPUs local to its GPU")
    ap.add_argument("--link-probe-s", type=float, default=0.3, help="seconds of the all-ranks D2H probe that measures the host link ceiling (0 = skip)")
    ap.add_argument("--scene-z0", type=float, default=-2.7,
                    help="world height z0 of the synthetic surface z = z0 + a sin(fx x) sin(fy y) (SURVEY 8(d)).  Default: the world origin sits "
                         "at the camera path, like in the reference's own tracks (tracks/*.yaml: bundle depths -3.9 .. -2.3, x and y around 0); "
                         "0 puts the surface ON the z = 0 plane -- the worst case of the bit-exact window-PCA kernel (every window near a zero "
                         "of z is accumulated sample by sample), the setting of rounds 1 and 2 up to profiles/bench_r2_final_*")
    ap.add_argument("--cam-step", type=float, default=0.006)
    ap.add_argument("--mesh-err", type=float, default=0.02)
    ap.add_argument("--cpu-pairs", type=int, default=2, help="frame pairs timed for cpu_baseline (rank 0, N=1)")
    ap.add_argument("--vr-impl", type=int, default=None)
    ap.add_argument("--contexts", type=int, default=3, help="library contexts (streams) per GPU; main frames alternate between them so that "
                    "one pair's kernel tails / low-occupancy phases overlap the other's (measured +15 %% at 2)")
    ap.add_argument("--exchange", choices=["auto", "mcast", "p2p", "nccl"], default="auto",
                    help="N > 1: how the point rows reach every rank -- mcast: ONE push per rank to an NVSwitch multicast address (a few CTAs "
                         "of multimem.st; lands in every rank's buffer); p2p: one copy-engine push per peer over NVLink (CUDA IPC), "
                         "no SMs; nccl: all_gather_into_tensor; auto = p2p (measured at 8 GPUs: 13.2 ms per step against 13.4 for mcast -- what "
                         "slows the step is the 7 x 464 MB ARRIVING at every GPU, which multicast does not reduce)")
    ap.add_argument("--xchg-repeat", type=int, default=1,
                    help="debug: push every slot this many times (emulates the per-GPU exchange volume of a larger world on few GPUs)")
    ap.add_argument("--no-e2e-balance", action="store_true",
                    help="N > 1: give every rank the same number of main frames per e2e step even when their host links differ")
    ap.add_argument("--xchg-from", type=int, default=-1,
                    help="debug: only this rank pushes its slot (separates the cost of SENDING rows from that of RECEIVING them: "
                         "compare diag.ms_by_rank); the content check is skipped")
    ap.add_argument("--graphs", type=int, default=None, choices=[0, 1, 2],
                    help="mr_set_use_graphs mode (default: the library's: graph replay when rows go to the host or with --farneback)")
    ap.add_argument("--farneback", action="store_true", help="run the reference's -f branch (not the headline configuration)")
    args = ap.parse_args()
    cfg5 = args.config == 5
    args.width = args.width or (3840 if cfg5 else 1920)
    args.height = args.height or (2160 if cfg5 else 1080)
    args.frames = args.frames or (600 if cfg5 else 300)
    args.sides = args.sides or (4 if cfg5 else 1)
    args.pairs = args.pairs or (4 if cfg5 else 8)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import mesh_reconstruction_b200 as mr
    from mesh_reconstruction_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    diag0 = {}
    affinity0 = os.sched_getaffinity(0)
    if not args.no_numa_bind:
        bind_to_gpu_numa(local, diag0)          # before the first pinned allocation
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    W, H, B, K, Wm, S = args.width, args.height, args.pairs, args.steps, args.warmup, args.sides
    N = W * H
    lib = mr.load_library()
    if args.vr_impl is not None:
        lib.mr_set_vr_impl(args.vr_impl)

    # ---- synthetic sequence: this rank's contiguous block of main frames --------------------
    scene = synth.make_scene(W, H, args.frames, step=args.cam_step, mesh_err=args.mesh_err, z0=args.scene_z0)
    block = max(2, args.frames // world)
    start = rank * block
    n_local = min(block, B * (K + Wm) + 1 + 2 * (S // 2), args.frames - start)
    idx = [start + i for i in range(n_local)]
    frames_dev = [scene.frame_torch(i, dev).contiguous() for i in idx]
    frames_pin = [f.cpu().pin_memory() for f in frames_dev]
    cams = scene.cameras
    nctx = max(1, args.contexts)
    renders = [mr.Render(W, H, ctx=mr.api.Context(W, H, local)) for _ in range(nctx)]
    for r_ in renders:
        r_.loadMesh(scene.vertices, scene.faces)
        if args.farneback:
            lib.mr_set_use_farneback(r_.ctx.h, 1)
        if args.graphs is not None:
            r_.ctx.set_use_graphs(args.graphs)
    render, ctx = renders[0], renders[0].ctx
    lib_streams = [torch.cuda.ExternalStream(r_.ctx.stream, device=dev) for r_ in renders]
    lib_stream = lib_streams[0]

    def join_streams():
        # everything queued on the other contexts' streams becomes a dependency of stream 0
        for st in lib_streams[1:]:
            lib_stream.wait_stream(st)

    nbuf = 2 if world > 1 else 1
    use_p2p = world > 1 and args.exchange != "nccl"
    use_mcast = False
    xch = None
    if use_p2p:
        # kernel-free exchange: per buffer set, every rank owns a receive buffer with one slot per rank (rows of B pairs at
        # full capacity, then the B counts).  p2p: the buffer is mapped into every peer over CUDA IPC, the normals kernel writes
        # this rank's rows straight into its own slot, which is then DMA'd into the same slot of every peer (mr_xchg_push);
        # mcast: the buffers are bound to an NVSwitch multicast object and ONE DMA per step reaches all of them
        from mesh_reconstruction_b200.shard import McastExchange, PeerExchange
        rows_bytes = (B * N * 28 + 255) // 256 * 256
        kinds = {"auto": ["p2p"], "mcast": ["mcast", "p2p"], "p2p": ["p2p"]}[args.exchange]
        for kind in kinds:
            xch = []
            try:
                for _ in range(nbuf):
                    xch.append((McastExchange if kind == "mcast" else PeerExchange)(ctx, rows_bytes + B * 4, dev))   # collective; raises on every rank or on none
                ok = 1
            except Exception as e:  # noqa: BLE001  (multicast / CUDA IPC / peer access unavailable on this box)
                print(f"bench.py: rank {rank}: {kind} exchange unavailable ({e})", file=sys.stderr)
                for x in xch:                    # an exchange that was already built must not stay mapped in the peers
                    x.close()
                ok = 0
            okt = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)           # every rank must take the same path
            if int(okt.item()) == 1:
                use_mcast = kind == "mcast"
                break
            xch = None
        if xch is None:
            print(f"bench.py: rank {rank}: using the NCCL all-gather", file=sys.stderr)
            use_p2p = False
    if use_p2p and use_mcast:
        rows_dev = [x.out((B, N, 7)) for x in xch]
        counts_dev = [x.out((B,), torch.int32, offset_bytes=rows_bytes) for x in xch]
        xflag = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(nbuf)]
    elif use_p2p:
        rows_dev = [x.slot(rank, (B, N, 7)) for x in xch]
        counts_dev = [x.slot(rank, (B,), torch.int32, offset_bytes=rows_bytes) for x in xch]
        for c_ in counts_dev:
            c_.zero_()
        xflag = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(nbuf)]
    else:
        rows_dev = [torch.empty((B, N, 7), dtype=torch.float32, device=dev) for _ in range(nbuf)]   # normals kernel writes straight into the send buffer
        counts_dev = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(nbuf)]
    # every pair of a step lands on the host; two alternating sets so that the host reads step s-1 while step s runs
    # (N > 1: a rank behind a faster host link takes up to 1.5 B main frames per e2e step, see balance_e2e below)
    Bmax = B if world == 1 or args.no_e2e_balance else B + (B + 1) // 2
    rows_pin = [[torch.empty((N, 7), dtype=torch.float32).pin_memory() for _ in range(Bmax)] for _ in range(2)]
    counts_pin = [torch.zeros(Bmax, dtype=torch.int32).pin_memory() for _ in range(2)]
    e2e_B = {"mine": B, "all": [B] * world}                     # main frames per e2e step of this rank / of every rank
    counts = torch.zeros(B, dtype=torch.int64)
    rows_flat = [r.view(B * N, 7) for r in rows_dev]
    gather_rows = [torch.empty((world * B * N, 7), dtype=torch.float32, device=dev) for _ in range(nbuf)] if world > 1 and not use_p2p else None
    gather_counts = [torch.empty((world * B,), dtype=torch.int32, device=dev) for _ in range(nbuf)] if world > 1 and not use_p2p else None
    pending = [None] * nbuf

    m_lo, m_hi = main_range(n_local, S)

    def pair(j):                      # j-th main frame of this rank (cycling inside its block) and its side frames
        a = m_lo + j % (m_hi - m_lo)
        return a, sides_of(a, S)

    def wait_pending(k):
        # make the LIBRARY stream (not torch's current stream) wait for the collective that still reads buffer k
        if pending[k] is not None:
            for st in lib_streams:
                with torch.cuda.stream(st):
                    for h in pending[k]:
                        h.wait()
            pending[k] = None

    def step_resident(s):
        # fully asynchronous: mr_submit_main_frame never waits for the GPU, the host queues pairs (and steps) back to back.
        # The normals kernel writes each pair's rows straight into the all-gather send buffer (slot b of the step).
        k = s % nbuf
        wait_pending(k)
        for b in range(B):
            a, cs = pair(s * B + b)
            mr.submit_main_frame(renders[b % nctx], frames_dev[a], cams[idx[a]], [frames_dev[c] for c in cs], [cams[idx[c]] for c in cs],
                                 out=rows_dev[k][b], out_count=counts_dev[k][b:b + 1])
        if world > 1:
            join_streams()
            if use_p2p:
                # the path's one exchange step without kernels: this rank's slot (rows + counts of the step) is DMA'd
                # into every peer's buffer over NVLink, stream-ordered after this step's kernels; a 4-byte all-reduce entered
                # after the pushes is the completion signal (when it is done everywhere, every slot of set k has landed)
                for _ in range(max(1, args.xchg_repeat)):
                    if args.xchg_from < 0 or args.xchg_from == rank:
                        xch[k].push()
                with torch.cuda.stream(xch[k].signal_stream()):
                    pending[k] = (dist.all_reduce(xflag[k], async_op=True),)
            else:
                # NCCL all-gather of rows + counts, ASYNC so that it overlaps the next step's compute (double-buffered send /
                # receive buffers, stream-ordered after this step's kernels)
                torch.cuda.current_stream().wait_stream(lib_stream)
                h1 = dist.all_gather_into_tensor(gather_counts[k], counts_dev[k], async_op=True)
                h2 = dist.all_gather_into_tensor(gather_rows[k], rows_flat[k], async_op=True)
                pending[k] = (h1, h2)

    def step_e2e(s):
        # host frames in (pinned; H2D inside the call), every pair's point rows + count out to pinned host memory by the
        # library's copy-engine DMA (overlapping the next pairs' compute).  The host queues step s, then waits until every
        # result of step s-1 has landed (mr_wait_copies_until: all but this step's copies) and reads it -- a two-deep
        # pipeline of pinned buffer sets, the way a long-running reconstruction consumes its main frames.
        k = s & 1
        Be = e2e_B["mine"]
        for b in range(Be):
            a, cs = pair(s * Be + b)
            mr.submit_main_frame(renders[b % nctx], frames_pin[a], cams[idx[a]], [frames_pin[c] for c in cs], [cams[idx[c]] for c in cs],
                                 out=rows_pin[k][b], out_count=counts_pin[k][b:b + 1])
        for c_, r_ in enumerate(renders):
            r_.ctx.wait_copies_until(len(range(c_, Be, nctx)))       # all but this step's copies of that context
        return int(counts_pin[k ^ 1].sum())

    def drain_e2e():
        for r_ in renders:
            r_.ctx.synchronize()          # the last step's rows are on the host before the timed region ends

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    diag = dict(diag0)

    def timed(fn, steps, first, drain=None, sampler=None):
        for k in range(nbuf):
            wait_pending(k)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = sum(r_.ctx.launches for r_ in renders)
        e0.record(lib_stream)
        t_host = time.perf_counter()
        for s in range(steps):
            fn(first + s)
        diag["host_ms_per_step"] = (time.perf_counter() - t_host) * 1e3 / steps     # CPU time to queue a step
        if sampler is not None and sampler.nv is not None:
            try:
                sampler.sample()          # the queue is still draining here: at least one sample under load, whatever the thread's timing
            except Exception:  # noqa: BLE001
                pass
        if drain is not None:
            drain()
        if world > 1:
            for k in range(nbuf):
                wait_pending(k)           # the exchange of every timed step completes inside the timed region
        join_streams()
        e1.record(lib_stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, diag["host_ms_per_step"]], dtype=torch.float64, device=dev)
            allt = torch.empty((world, 2), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allt, t)
            diag["ms_by_rank"] = [round(float(v) / steps, 3) for v in allt[:, 0]]
            diag["host_ms_by_rank"] = [round(float(v), 3) for v in allt[:, 1]]
            ms = float(allt[:, 0].max())
        return ms, sum(r_.ctx.launches for r_ in renders) - l0

    # ---- host link: every rank copies device -> pinned host at once for ~0.3 s (outside the timed regions) ----------------
    def probe_link():
        for r_ in renders:
            r_.ctx.synchronize()
        src_rows = rows_dev[0].view(-1)[: N * 7 * min(B, 4)]
        dst_rows = [rows_pin[0][b].view(-1) for b in range(min(B, 4))]
        chunk = N * 7
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

        def burst(reps):
            for i in range(reps):
                with torch.cuda.stream(streams[i & 1]):
                    j = i % len(dst_rows)
                    dst_rows[j].copy_(src_rows[j * chunk:(j + 1) * chunk], non_blocking=True)
        burst(4)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(8, int(args.link_probe_s * 50e9 / (chunk * 4)))
        e0.record(torch.cuda.current_stream())
        for st_ in streams:
            st_.wait_stream(torch.cuda.current_stream())
        burst(reps)
        for st_ in streams:
            torch.cuda.current_stream().wait_stream(st_)
        e1.record(torch.cuda.current_stream())
        barrier()
        ms_probe = e0.elapsed_time(e1)
        mine_gbs = reps * chunk * 4 / (ms_probe * 1e-3) / 1e9
        if world > 1:
            t = torch.tensor([ms_probe, mine_gbs], dtype=torch.float64, device=dev)
            allp = torch.empty((world, 2), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allp, t)
            # equal_split_gbs: what the ranks deliver together when each has the same bytes to move (the slowest link sets the
            # time); sum_gbs: when each moves bytes in proportion to its own rate
            return {"equal_split_gbs": world * reps * chunk * 4 / (float(allp[:, 0].max()) * 1e-3) / 1e9,
                    "sum_gbs": float(allp[:, 1].sum()), "by_rank_gbs": [round(float(v), 1) for v in allp[:, 1]]}
        return {"equal_split_gbs": mine_gbs, "sum_gbs": mine_gbs, "by_rank_gbs": [round(mine_gbs, 1)]}

    def balance_e2e(link, ms_frame_compute):
        """e2e at N > 1 is bound by the host links, and on this pool's boxes they are not equal (four GPUs reach ~12 GB/s, four
        ~18 GB/s when all copy at once).  Main frames are independent, so every rank takes them in proportion to the rate it
        can sustain: one frame costs it max(compute time, row bytes / its link rate)."""
        if world == 1 or args.no_e2e_balance or not link:
            return
        rates = link["by_rank_gbs"]
        if min(rates) <= 0:
            return
        ms_frame_link = [(N * 28 + 4) / (r_ * 1e9) * 1e3 for r_ in rates]
        w = [1.0 / max(ms_frame_compute, t_) for t_ in ms_frame_link]
        e2e_B["all"] = [min(Bmax, max(1, int(round(B * world * w_ / sum(w))))) for w_ in w]
        e2e_B["mine"] = e2e_B["all"][rank]

    # ---- warm-up, then the timed regions ---------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for s in range(Wm):
        step_resident(s)
    sampler.recording = True
    ms_res, launches = timed(step_resident, K, Wm, sampler=sampler)
    sampler.recording = False
    diag_res = dict(diag)
    sampler.stop_flag = True
    link = probe_link() if args.link_probe_s > 0 else None
    balance_e2e(link, ms_res / K / B)
    for s in range(max(Wm - 2, 1)):
        step_e2e(s)
    drain_e2e()
    ms_e2e, _ = timed(step_e2e, K, Wm, drain=drain_e2e)
    sampler.join(timeout=2)

    # ---- per-stage breakdown of one more step (events inside the library) -----------------------
    import ctypes as C
    lib.mr_profile_enable(ctx.h, 1)
    nctx_save, nctx = nctx, 1            # stage breakdown: one context, kernels back to back (no overlap between pairs)
    for s in range(2):
        step_resident(Wm + K + s)
    for k in range(nbuf):
        wait_pending(k)
    nctx = nctx_save
    msb, lb = (C.c_double * 8)(), (C.c_uint64 * 8)()
    ns = lib.mr_profile_read(ctx.h, msb, lb, 8)
    lib.mr_profile_enable(ctx.h, 0)
    prof_pairs = 2 * B
    stages = {lib.mr_stage_name(i).decode(): {"ms_per_pair": msb[i] / prof_pairs, "launches_per_pair": lb[i] / prof_pairs} for i in range(ns)}
    if world > 1:
        # per-rank health for the scaling runs: kernel time per pair (one context, no overlap), SM clock under load
        cs = sampler.summary()
        mine = torch.tensor([sum(v["ms_per_pair"] for v in stages.values()), float(cs.get("sm_mhz") or 0), float(cs.get("sm_min_mhz") or 0),
                             float(sampler.reasons)], dtype=torch.float64, device=dev)
        allh = torch.empty((world, 4), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allh, mine)
        diag_res["kernel_ms_per_pair_by_rank"] = [round(float(v), 4) for v in allh[:, 0]]
        diag_res["sm_mhz_by_rank"] = [int(v) for v in allh[:, 1]]
        diag_res["sm_min_mhz_by_rank"] = [int(v) for v in allh[:, 2]]
        diag_res["clock_reason_bits_by_rank"] = [int(v) for v in allh[:, 3]]
    dom = max(stages, key=lambda k: stages[k]["ms_per_pair"])
    peak, peak_src = load_peaks()
    stage_bytes = {"raster": 4 + 4 * S, "shade_mix": 11 * S, "variational_refinement": 10 * S, "cubic_remap": 10 * S, "pyramid_compare": 6 * S,
                   "triangulate": 16 * S + 4 + 20, "normals": 20 + 28}             # per main-frame pixel, DESIGN.md "Kernels"
    dom_ms_launch = stages[dom]["ms_per_pair"] / max(stages[dom]["launches_per_pair"], 1)
    dom_bytes_launch = stage_bytes[dom] * N / max(stages[dom]["launches_per_pair"], 1)
    ach = dom_bytes_launch / (dom_ms_launch * 1e-3) / 1e9
    path_bytes_step = ALG_BYTES_PATH(S) * N * B                                     # (33 + 21 S) B per main-frame pixel
    path_ach = path_bytes_step * K * 1.0 / (ms_res * 1e-3) / 1e9                   # per GPU

    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "kernels_r2.json")))
    except Exception:
        pass
    same_shape = prof.get("shape") == [W, H, S]
    traffic = prof.get("dram_bytes_per_main_frame") if same_shape else None
    clocks = sampler.summary()
    compute = None
    if same_shape and prof.get("warp_inst_per_main_frame") and clocks.get("sm_mhz"):
        # issue-slot roofline: warp instructions of one main frame (ncu smsp__inst_executed.sum of every launch, committed
        # profile) x main frames per second, against 4 schedulers x 148 SMs x the SM clock sampled in this run
        inst_rate = prof["warp_inst_per_main_frame"] * B * K / (ms_res * 1e-3)
        issue_peak = 148 * 4 * clocks["sm_mhz"] * 1e6
        compute = {"bound": "issue", "achieved": inst_rate / 1e12, "peak": issue_peak / 1e12, "unit": "T warp-inst/s", "frac": inst_rate / issue_peak,
                   "warp_inst_per_main_frame": prof["warp_inst_per_main_frame"], "source": prof.get("source"),
                   "pipes_pct_of_peak": prof.get("pipes_pct_of_peak"),
                   "note": "the path reproduces the reference's IEEE arithmetic operation by operation (exact div / sqrt, FP64 islands, 50 Newton "
                           "iterations, 441-term sequential window sums): every large kernel is issue- or XU/FP64-pipe bound, DRAM throughput is 1-3 %"}
    pix_total = world * B * K * N * S
    value = pix_total / (ms_res * 1e-3) / 1e6
    e2e_val = sum(e2e_B["all"]) * K * N * S / (ms_e2e * 1e-3) / 1e6
    m_mean = float(counts_dev[(Wm + K - 1) % nbuf].float().mean())

    # ---- exchange content check (N > 1, outside the timed regions): every slot every rank received equals, bit for bit,
    # what its producer holds; and the compacted rank-ordered cloud (the reference's append order, recon.cpp:115-116)
    # built from the slots equals the NCCL all-gather of the compacted rows ------------------------------------------------
    xcheck = None
    if world > 1:
        for r_ in renders:
            r_.ctx.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        if use_p2p and args.xchg_from < 0:
            from mesh_reconstruction_b200 import shard
            bad = 0
            for k_ in range(nbuf):
                mine_sum = torch.stack([rows_dev[k_].view(torch.int32).sum(dtype=torch.int64), counts_dev[k_].sum(dtype=torch.int64)])
                allsum = torch.empty((world, 2), dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(allsum, mine_sum)
                for p_ in range(world):
                    got = torch.stack([xch[k_].slot(p_, (B, N, 7)).view(torch.int32).sum(dtype=torch.int64),
                                       xch[k_].slot(p_, (B,), torch.int32, offset_bytes=rows_bytes).sum(dtype=torch.int64)])
                    if not torch.equal(got, allsum[p_]):
                        bad += 1
            # compacted cloud of the last step from the peer slots vs the torch.distributed all-gather of this rank's compacted rows
            k_ = (Wm + K + 1) % nbuf                       # buffer set of the last step that ran (the two profiling steps)
            cnts = torch.stack([xch[k_].slot(p_, (B,), torch.int32, offset_bytes=rows_bytes) for p_ in range(world)]).cpu()
            cloud = torch.cat([xch[k_].slot(p_, (B, N, 7))[b, :int(cnts[p_, b])] for p_ in range(world) for b in range(B)], 0)
            own = torch.cat([rows_dev[k_][b, :int(cnts[rank, b])] for b in range(B)], 0)
            ref_cloud, _ = shard.allgather_points(own, len(own))
            same = bool(cloud.shape == ref_cloud.shape and torch.equal(cloud.view(torch.int32), ref_cloud.view(torch.int32)))
            okt = torch.tensor([0 if (bad or not same) else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            xcheck = {"slots_bit_identical_on_every_rank": bool(int(okt.item())), "cloud_rows": int(cloud.shape[0]),
                      "rank_ordered_cloud_equals_nccl_allgather": same}
            if not int(okt.item()):
                raise SystemExit(f"bench.py: rank {rank}: exchanged rows differ from their producer's ({bad} slots) or the compacted cloud differs from the all-gather")

    cpu = None
    if rank == 0 and world == 1 and args.cpu_pairs > 0:
        os.sched_setaffinity(0, affinity0)        # the CPU baseline uses every host core
        cores = os.cpu_count() or 1
        n_main = max(1, args.cpu_pairs // S)
        fr = {i: frames_pin[i].numpy() for i in range(min(n_local, n_main + 2 * (S // 2) + (1 if S == 1 else 0)))}
        sc4 = synth.make_scene(W, H, args.frames, step=args.cam_step, mesh_err=args.mesh_err, z0=args.scene_z0)
        sc4.cameras = cams[idx[0]:idx[0] + len(fr)]
        lo_, hi_ = main_range(len(fr), S)
        prs = [(a_, sides_of(a_, S)) for a_ in range(lo_, hi_)][:n_main]
        t, _ = cpu_reference_pairs(sc4, fr, prs, cores, args.farneback)
        import cv2
        cpu = {"value": len(prs) * S * N / t / 1e6, "unit": "Mpix/s", "cores": cores, "kind": "port",
               "sample": f"{len(prs)} main frame(s) x {S} side frame(s) of {W}x{H} (of the {args.frames}-frame workload), cv2 {cv2.__version__} x{cores} threads + C restatement (OpenMP)"}

    # ---- the step after the path, outside the timed regions: Heuristic::filterPoints on the rows of ONE main frame,
    # on the device, against its CPU restatement (bounded: one frame's cloud, radius = 3 pixel spacings) ---------------------
    filt = None
    if rank == 0 and world == 1 and args.cpu_pairs > 0 and not args.no_filter_bench:
        try:
            from oracle import filter as ofilter
            k_ = (Wm + K + 1) % nbuf
            m0 = int(counts_dev[k_][0])
            rows0 = rows_dev[k_][0, :m0].contiguous()
            fin = torch.isfinite(rows0).all(1)
            rows0 = rows0[fin].contiguous()
            d3 = rows0[:, :3] / rows0[:, 3:4]
            # pixel spacing on the surface from the 1 % .. 99 % range of x: a handful of rows are far outliers (pixels whose
            # Newton iteration ran away -- reference behaviour), and a radius derived from the raw extent would be absurd
            qs = torch.quantile(d3[::8, 0].contiguous(), torch.tensor([0.01, 0.99], device=dev))
            spacing = float(qs[1] - qs[0]) / (0.98 * W)
            radius = (3.0 * spacing) ** 2
            out_rows = torch.empty_like(rows0)
            mr.filter_rows(rows0, radius, ctx=ctx, out=out_rows)          # first call allocates the tables
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            kept_rows, keep = mr.filter_rows(rows0, radius, ctx=ctx, out=out_rows)
            torch.cuda.synchronize()
            t_gpu = time.perf_counter() - t0
            info = mr.api.filter_info(ctx)
            host_rows = rows0.cpu().numpy()
            t0 = time.perf_counter()
            ref = ofilter.filter_points(host_rows[:, :4], radius)
            t_cpu = time.perf_counter() - t0
            filt = {"what": "Heuristic::filterPoints (heuristic.cpp:55-176) on the rows of one main frame, device resident", "points": int(len(rows0)),
                    "neighbour_pairs": int(info["n_edges"]), "power_iterations": int(info["iters"]), "thinning_rounds": int(info["rounds"]),
                    "survivors": int(len(keep)), "gpu_ms": 1e3 * t_gpu, "cpu_ms": 1e3 * t_cpu, "cpu_kind": "port (oracle/filter_oracle.cpp, 1 core)",
                    "survivors_bit_identical_to_cpu": bool(np.array_equal(keep.cpu().numpy(), ref["keep"])),
                    "d2h_bytes_if_filtered_on_device": int(len(keep)) * 28, "d2h_bytes_unfiltered": int(len(rows0)) * 28}
        except Exception as e:  # noqa: BLE001
            filt = {"error": str(e)[:200]}

    if rank == 0:
        fr_step = sum(e2e_B["all"]) / world                      # main frames per step and GPU (mean over ranks)
        d2h_step = (N * 28 + 4) * fr_step
        e2e = {"value": e2e_val, "unit": "Mpix/s", "h2d_bytes_per_step": int((1 + S) * N * fr_step), "d2h_bytes_per_step": int(d2h_step),   # per GPU; the DMA moves the full row capacity (count unknown on the host without a sync)
               "rows_bytes_per_step": int(m_mean * 28 * fr_step), "ms_per_step": ms_e2e / K}
        if world > 1:
            e2e["main_frames_per_step_by_rank"] = e2e_B["all"]
        if link:
            # share of the probed host-link capacity that the e2e run's row traffic reaches.  host_link_gbs: every rank copying the
            # SAME number of bytes at once (the slowest link sets the time); the sum of the per-rank rates is an upper bound that
            # sustained traffic does not reach (measured: 97 of 122 GB/s with frames divided in proportion to the rates)
            e2e["host_link_gbs"] = link["equal_split_gbs"]
            e2e["host_link_sum_of_rank_rates_gbs"] = link["sum_gbs"]
            e2e["host_link_by_rank_gbs"] = link["by_rank_gbs"]
            e2e["d2h_achieved_gbs"] = world * d2h_step / (ms_e2e / K * 1e-3) / 1e9
            e2e["host_link_frac"] = e2e["d2h_achieved_gbs"] / link["equal_split_gbs"]
            e2e["bound"] = "host link" if e2e["host_link_frac"] > 0.85 else "compute"
        line = {
            "metric": "Mpix/s matched+triangulated", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "main_frames_per_step_per_gpu": B, "side_frames": S, "pairs_per_step_per_gpu": B * S,
                       "contexts_per_gpu": nctx, "flow": "farneback (-f)" if args.farneback else "variational refinement (reference default)", "mesh_faces": int(len(scene.faces)), "points_per_main_frame": m_mean,
                       "l2": f"working set per step ({B} main frames x ~{((16 * 4 + 24) * S + 40) * N / 1e6:.0f} MB of planes) exceeds the 126 MB L2; no explicit flush",
                       "exchange": ("none (single GPU)" if world == 1 else
                                    "ONE push of rows + counts per rank and step to an NVSwitch multicast address (mr_xchg_push_mcast: a few CTAs of 128-bit multimem.st on a high-priority stream; the switch replicates it into every rank's buffer), overlapped with the next step" if use_mcast else
                                    "copy-engine pushes of rows + counts into every peer's buffer over NVLink (CUDA IPC, mr_xchg_push; no SMs), overlapped with the next step" if use_p2p else
                                    "async nccl all_gather_into_tensor of point rows + counts per step, overlapped with the next step")},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "whole main-frame step (every kernel of the path)", "achieved": path_ach, "peak": peak, "unit": "GB/s",
                         "frac": path_ach / peak, "traffic": traffic, "peak_source": peak_src, "alg_bytes_per_main_frame_pixel": ALG_BYTES_PATH(S),
                         "alg_bytes_per_step": path_bytes_step, "ms_per_step": ms_res / K,
                         "note": "SURVEY 8(d) contract figure (33 + 21 S bytes per main-frame pixel); the path is not HBM-bound, see roofline_compute"},
            "roofline_kernel": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "alg_bytes_per_pixel": stage_bytes[dom], "ms_per_launch": dom_ms_launch, "launches_per_main_frame": stages[dom]["launches_per_pair"]},
            "roofline_compute": compute,
            "stages_ms_per_main_frame": {k: round(v["ms_per_pair"], 4) for k, v in stages.items()},
            "clocks": clocks,
            "diag": diag_res,      # host time to queue a step; per-rank device time per step (value uses the max)
        }
        if xcheck:
            line["exchange_check"] = xcheck
        if filt:
            line["filter_points"] = filt
        if cpu:
            line["cpu_baseline"] = cpu
        try:
            line["normals_stats"] = dict(zip(["tiles", "tiles_1_sampled_coord", "tiles_2_sampled_coords", "tiles_3_sampled_coords", "residual_pixels"], ctx.normals_stats()))
        except Exception:
            pass
        print(json.dumps(line), flush=True)
    if world > 1:
        if use_p2p:
            for x in xch:
                x.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
