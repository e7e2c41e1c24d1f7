"""Which stages slow down while the copy engine moves the previous pair's rows to the host?  One context, stage
events inside the library (mr_profile_*), rows to device memory vs. to pinned host memory."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import synth
W, H, B = 1920, 1080, 16
N = W * H
sc = synth.make_scene(W, H, 300, step=0.006, mesh_err=0.02)
dev = torch.device("cuda", 0)
fd = [sc.frame_torch(i, dev).contiguous() for i in range(B + 1)]
r = mr.Render(W, H, ctx=mr.api.Context(W, H, 0)); r.loadMesh(sc.vertices, sc.faces)
lib = r.ctx.lib
rows_dev = torch.empty((N, 7), dtype=torch.float32, device=dev)
cnt_dev = torch.zeros(1, dtype=torch.int32, device=dev)
rows_pin = [torch.empty((N, 7), dtype=torch.float32).pin_memory() for _ in range(4)]
cnt_pin = torch.zeros(4, dtype=torch.int32).pin_memory()
def run(host_out):
    r.ctx.synchronize()
    lib.mr_profile_enable(r.ctx.h, 1)
    t0 = time.perf_counter()
    for b in range(B):
        if host_out:
            mr.submit_main_frame(r, fd[b], sc.cameras[b], [fd[b + 1]], [sc.cameras[b + 1]], out=rows_pin[b & 3], out_count=cnt_pin[(b & 3):(b & 3) + 1])
        else:
            mr.submit_main_frame(r, fd[b], sc.cameras[b], [fd[b + 1]], [sc.cameras[b + 1]], out=rows_dev, out_count=cnt_dev)
    r.ctx.synchronize()
    dt = (time.perf_counter() - t0) / B * 1e3
    ms, lb = (C.c_double * 8)(), (C.c_uint64 * 8)()
    ns = lib.mr_profile_read(r.ctx.h, ms, lb, 8)
    lib.mr_profile_enable(r.ctx.h, 0)
    return dt, {lib.mr_stage_name(i).decode(): round(ms[i] / B, 4) for i in range(ns)}
for ho in (False, True, False, True):
    run(ho)
    print("host rows" if ho else "dev rows ", run(ho))

# same, rows to device memory, but an UNRELATED 58 MB D2H issued from another stream at every submit
bg_src = torch.empty((N, 7), dtype=torch.float32, device=dev)
bg_stream = torch.cuda.Stream()
def run_bg(kind):
    r.ctx.synchronize(); torch.cuda.synchronize()
    lib.mr_profile_enable(r.ctx.h, 1)
    for b in range(B):
        mr.submit_main_frame(r, fd[b], sc.cameras[b], [fd[b + 1]], [sc.cameras[b + 1]], out=rows_dev, out_count=cnt_dev)
        with torch.cuda.stream(bg_stream):
            if kind == "d2h": rows_pin[b & 3].copy_(bg_src, non_blocking=True)
            elif kind == "h2d": bg_src.copy_(rows_pin[b & 3], non_blocking=True)
            elif kind == "d2d": bg_src.copy_(rows_dev, non_blocking=True)
    r.ctx.synchronize(); torch.cuda.synchronize()
    ms, lb = (C.c_double * 8)(), (C.c_uint64 * 8)()
    ns = lib.mr_profile_read(r.ctx.h, ms, lb, 8)
    lib.mr_profile_enable(r.ctx.h, 0)
    return {lib.mr_stage_name(i).decode(): round(ms[i] / B, 4) for i in range(ns)}
for kind in ("d2h", "h2d", "d2d"):
    run_bg(kind)
    print("background", kind, run_bg(kind))
