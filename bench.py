#!/usr/bin/env python
"""bench.py -- headline benchmark of the dense-correspondence hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU arithmetic (oracle)

Metric: Mpix/s matched + triangulated = pixels of all processed (main, side) pairs / time.
Workload (config 4 of BASELINE.json): synthetic 1920x1080 300-frame sequence, adjacent-pair
matching (main i, side i+1, S = 1).  One "step" = `--pairs` frame pairs per GPU through the whole
path (depth raster -> shadow raster + dilation -> reproject + mixBackground -> variational
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
    """Samples SM clocks / throttle reasons (NVML, every few ms) while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm = index, False, [], 0, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.reasons |= nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    self.reasons |= nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                time.sleep(0.004)
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["unavailable: " + getattr(self, "err", "no samples")]}
        sm = sorted(self.sm)
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_sm,
                "reasons": [n for b, n in bits.items() if self.reasons & b], "samples": len(sm)}


def cpu_reference_pairs(scene, frames, pairs, threads, farneback=False):
    """Times the CPU oracle (reference arithmetic) on the given (main, side) index pairs."""
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
        tri = process_main_frame(r, frames, scene.cameras, fa, [fb], use_farneback=farneback)
        pts += len(tri)
    return time.perf_counter() - t0, pts


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
    scene = synth.make_scene(W, H, 4, step=args.cam_step, mesh_err=args.mesh_err)
    frames = {i: scene.frame(i) for i in range(4)}
    pairs_cycle = [(0, 1), (1, 2), (2, 3)]
    if args.warmup > 0:      # one real warm-up pair is enough (each pair is seconds of CPU work)
        cpu_reference_pairs(scene, frames, [pairs_cycle[0]], cores, args.farneback)
    t_total = 0.0
    for k in range(args.steps):
        t, _ = cpu_reference_pairs(scene, frames, [pairs_cycle[k % 3]], cores, args.farneback)
        t_total += t
    pix = args.steps * W * H
    val = pix / t_total / 1e6
    line = {
        "impl": "reference", "metric": "Mpix/s matched+triangulated", "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic {W}x{H} sequence, adjacent-pair matching + triangulation, S=1",
                   "pairs_per_step": 1, "sample": "1 frame pair per step (bounded sample of the 299-pair workload)"},
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} frame pairs of {W}x{H}, cv2 {__import__('cv2').__version__} with {cores} threads + C restatement (OpenMP)"},
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
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--pairs", type=int, default=8, help="frame pairs per GPU per step")
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--cam-step", type=float, default=0.006)
    ap.add_argument("--mesh-err", type=float, default=0.02)
    ap.add_argument("--cpu-pairs", type=int, default=2, help="frame pairs timed for cpu_baseline (rank 0, N=1)")
    ap.add_argument("--vr-impl", type=int, default=None)
    ap.add_argument("--contexts", type=int, default=3, help="library contexts (streams) per GPU; main frames alternate between them so that "
                    "one pair's kernel tails / low-occupancy phases overlap the other's (measured +15 %% at 2)")
    ap.add_argument("--exchange", choices=["p2p", "nccl"], default="p2p",
                    help="N > 1: how the point rows reach every rank -- p2p: copy-engine pushes into the peers' buffers over NVLink "
                         "(CUDA IPC, no SMs); nccl: all_gather_into_tensor")
    ap.add_argument("--xchg-repeat", type=int, default=1,
                    help="debug: push every slot this many times (emulates the per-GPU exchange volume of a larger world on few GPUs)")
    ap.add_argument("--graphs", type=int, default=None, choices=[0, 1, 2],
                    help="mr_set_use_graphs mode (default: the library's: graph replay when rows go to the host or with --farneback)")
    ap.add_argument("--farneback", action="store_true", help="run the reference's -f branch (not the headline configuration)")
    args = ap.parse_args()
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
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    W, H, B, K, Wm = args.width, args.height, args.pairs, args.steps, args.warmup
    N = W * H
    lib = mr.load_library()
    if args.vr_impl is not None:
        lib.mr_set_vr_impl(args.vr_impl)

    # ---- synthetic sequence: this rank's contiguous block of main frames --------------------
    scene = synth.make_scene(W, H, args.frames, step=args.cam_step, mesh_err=args.mesh_err)
    block = max(2, args.frames // world)
    start = rank * block
    n_local = min(block, B * (K + Wm) + 1, args.frames - start)
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
    use_p2p = world > 1 and args.exchange == "p2p"
    xch = None
    if use_p2p:
        # kernel-free exchange: per buffer set, every rank owns a receive buffer with one slot per rank (rows of B pairs at
        # full capacity, then the B counts), mapped into every peer over CUDA IPC; the normals kernel writes this rank's rows
        # straight into its own slot, which is then DMA'd into the same slot of every peer (mr_xchg_push)
        from mesh_reconstruction_b200.shard import PeerExchange
        rows_bytes = (B * N * 28 + 255) // 256 * 256
        try:
            xch = [PeerExchange(ctx, rows_bytes + B * 4, dev) for _ in range(nbuf)]
            ok = 1
        except Exception as e:  # noqa: BLE001  (CUDA IPC / peer access unavailable on this box)
            print(f"bench.py: rank {rank}: peer-memory exchange unavailable ({e}); using the NCCL all-gather", file=sys.stderr)
            ok = 0
        okt = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)           # every rank must take the same path
        if int(okt.item()) == 0:
            use_p2p, xch = False, None
    if use_p2p:
        rows_dev = [x.slot(rank, (B, N, 7)) for x in xch]
        counts_dev = [x.slot(rank, (B,), torch.int32, offset_bytes=rows_bytes) for x in xch]
        for c_ in counts_dev:
            c_.zero_()
        xflag = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(nbuf)]
    else:
        rows_dev = [torch.empty((B, N, 7), dtype=torch.float32, device=dev) for _ in range(nbuf)]   # normals kernel writes straight into the send buffer
        counts_dev = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(nbuf)]
    # every pair of a step lands on the host; two alternating sets so that the host reads step s-1 while step s runs
    rows_pin = [[torch.empty((N, 7), dtype=torch.float32).pin_memory() for _ in range(B)] for _ in range(2)]
    counts_pin = [torch.zeros(B, dtype=torch.int32).pin_memory() for _ in range(2)]
    per_ctx = [len(range(c, B, nctx)) for c in range(nctx)]     # pairs each context gets per step
    counts = torch.zeros(B, dtype=torch.int64)
    rows_flat = [r.view(B * N, 7) for r in rows_dev]
    gather_rows = [torch.empty((world * B * N, 7), dtype=torch.float32, device=dev) for _ in range(nbuf)] if world > 1 and not use_p2p else None
    gather_counts = [torch.empty((world * B,), dtype=torch.int32, device=dev) for _ in range(nbuf)] if world > 1 and not use_p2p else None
    pending = [None] * nbuf

    def pair(j):                      # j-th pair of this rank, cycling inside its block
        a = j % (n_local - 1)
        return a, a + 1

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
            a, c = pair(s * B + b)
            mr.submit_main_frame(renders[b % nctx], frames_dev[a], cams[idx[a]], [frames_dev[c]], [cams[idx[c]]],
                                 out=rows_dev[k][b], out_count=counts_dev[k][b:b + 1])
        if world > 1:
            join_streams()
            if use_p2p:
                # the path's one exchange step without kernels: this rank's slot (rows + counts of the step) is DMA'd
                # into every peer's buffer over NVLink, stream-ordered after this step's kernels; a 4-byte all-reduce entered
                # after the pushes is the completion signal (when it is done everywhere, every slot of set k has landed)
                for _ in range(max(1, args.xchg_repeat)):
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
        for b in range(B):
            a, c = pair(s * B + b)
            mr.submit_main_frame(renders[b % nctx], frames_pin[a], cams[idx[a]], [frames_pin[c]], [cams[idx[c]]],
                                 out=rows_pin[k][b], out_count=counts_pin[k][b:b + 1])
        for c_, r_ in enumerate(renders):
            r_.ctx.wait_copies_until(per_ctx[c_])
        return int(counts_pin[k ^ 1].sum())

    def drain_e2e():
        for r_ in renders:
            r_.ctx.synchronize()          # the last step's rows are on the host before the timed region ends

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    diag = {}

    def timed(fn, steps, first, drain=None):
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

    # ---- warm-up, then the timed regions ---------------------------------------------------
    for s in range(Wm):
        step_resident(s)
    sampler = ClockSampler(local)
    sampler.start()
    ms_res, launches = timed(step_resident, K, Wm)
    diag_res = dict(diag)
    sampler.stop_flag = True
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
    dom_ms_launch = stages[dom]["ms_per_pair"] / max(stages[dom]["launches_per_pair"], 1)
    dom_bytes_launch = STAGE_ALG_BYTES[dom] * N / max(stages[dom]["launches_per_pair"], 1)
    ach = dom_bytes_launch / (dom_ms_launch * 1e-3) / 1e9
    path_ach = ALG_BYTES_PATH(1) * N * B * K * 1.0 / (ms_res * 1e-3) / 1e9     # per GPU

    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if (W, H) == (1920, 1080) and dom in tj and tj[dom].get("dram_bytes_per_launch"):
            traffic = tj[dom]["dram_bytes_per_launch"]
    except Exception:
        pass
    pix_total = world * B * K * N
    value = pix_total / (ms_res * 1e-3) / 1e6
    e2e_val = pix_total / (ms_e2e * 1e-3) / 1e6
    m_mean = float(counts_dev[(Wm + K - 1) % nbuf].float().mean())

    cpu = None
    if rank == 0 and world == 1 and args.cpu_pairs > 0:
        cores = os.cpu_count() or 1
        fr = {i: frames_pin[i].numpy() for i in range(min(n_local, args.cpu_pairs + 1))}
        sc4 = synth.make_scene(W, H, args.frames, step=args.cam_step, mesh_err=args.mesh_err)
        sc4.cameras = cams[idx[0]:idx[0] + len(fr)]
        prs = [(i, i + 1) for i in range(len(fr) - 1)]
        t, _ = cpu_reference_pairs(sc4, fr, prs, cores, args.farneback)
        import cv2
        cpu = {"value": len(prs) * N / t / 1e6, "unit": "Mpix/s", "cores": cores, "kind": "port",
               "sample": f"{len(prs)} frame pairs of {W}x{H} (of the 299-pair workload), cv2 {cv2.__version__} x{cores} threads + C restatement (OpenMP)"}

    if rank == 0:
        line = {
            "metric": "Mpix/s matched+triangulated", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"synthetic {W}x{H} {args.frames}-frame sequence, adjacent-pair matching + triangulation, S=1 (BASELINE config 4)",
                       "pairs_per_step_per_gpu": B, "contexts_per_gpu": nctx, "flow": "farneback (-f)" if args.farneback else "variational refinement (reference default)", "mesh_faces": int(len(scene.faces)), "points_per_pair": m_mean,
                       "l2": f"working set per step ({B} pairs x ~{(16 * 4 + 40) * N / 1e6:.0f} MB of planes) exceeds the 126 MB L2; no explicit flush",
                       "exchange": ("none (single GPU)" if world == 1 else
                                    "copy-engine pushes of rows + counts into every peer's buffer over NVLink (CUDA IPC, mr_xchg_push; no SMs), overlapped with the next step" if use_p2p else
                                    "async nccl all_gather_into_tensor of point rows + counts per step, overlapped with the next step")},
            "e2e": {"value": e2e_val, "unit": "Mpix/s", "h2d_bytes_per_step": 2 * N * B, "d2h_bytes_per_step": (N * 28 + 4) * B,   # the DMA moves the full row capacity (count unknown on the host without a sync)
                    "rows_bytes_per_step": int(m_mean * 28 * B),
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "note": "compute-bound stage (IEEE-exact div/sqrt, FP64 islands): HBM fraction is low by construction, see DESIGN.md section 4", "alg_bytes_per_pixel": STAGE_ALG_BYTES[dom],
                         "ms_per_launch": dom_ms_launch, "launches_per_pair": stages[dom]["launches_per_pair"]},
            "roofline_path": {"alg_bytes_per_pixel_pair": ALG_BYTES_PATH(1), "achieved": path_ach, "peak": peak, "unit": "GB/s",
                              "frac": path_ach / peak},
            "stages_ms_per_pair": {k: round(v["ms_per_pair"], 4) for k, v in stages.items()},
            "clocks": sampler.summary(),
            "diag": diag_res,      # host time to queue a step; per-rank device time per step (value uses the max)
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        if use_p2p:
            # every rank must hold every rank's counts of the last steps (exchange sanity, outside the timed regions)
            for r_ in renders:
                r_.ctx.synchronize()
            torch.cuda.synchronize()
            dist.barrier()
            for x in xch:
                got = torch.stack([x.slot(p, (B,), torch.int32, offset_bytes=rows_bytes) for p in range(world)])
                if int(got.min()) <= 0:
                    raise SystemExit(f"bench.py: rank {rank} is missing exchanged counts: {got.tolist()}")
            for x in xch:
                x.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
