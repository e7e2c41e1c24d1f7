import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import synth
W, H, B = 1920, 1080, 24
N = W * H
sc = synth.make_scene(W, H, 300, step=0.006, mesh_err=0.02)
dev = torch.device("cuda", 0)
fd = [sc.frame_torch(i, dev).contiguous() for i in range(B + 1)]
fp = [f.cpu().pin_memory() for f in fd]
r = mr.Render(W, H, ctx=mr.api.Context(W, H, 0)); r.loadMesh(sc.vertices, sc.faces)
rows_dev = torch.empty((N, 7), dtype=torch.float32, device=dev)
rows_pin = [torch.empty((N, 7), dtype=torch.float32).pin_memory() for _ in range(2)]
def run(frames, host_out, async_copy):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for b in range(B):
        f0 = frames[b] if isinstance(frames[b], torch.Tensor) and frames[b].is_cuda else frames[b].numpy()
        f1 = frames[b + 1] if isinstance(frames[b + 1], torch.Tensor) and frames[b + 1].is_cuda else frames[b + 1].numpy()
        if host_out:
            mr.process_main_frame(r, f0, sc.cameras[b], [f1], [sc.cameras[b + 1]], out=rows_pin[b & 1].numpy(), want_host=True, async_copy=async_copy)
        else:
            mr.process_main_frame(r, f0, sc.cameras[b], [f1], [sc.cameras[b + 1]], out=rows_dev, want_host=False)
    r.ctx.wait_copies(); r.ctx.synchronize(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / B * 1e3
for name, fr, ho, ac in [("dev frames, dev rows", fd, False, False), ("host frames, dev rows", fp, False, False),
                         ("dev frames, host rows sync", fd, True, False), ("dev frames, host rows async", fd, True, True),
                         ("host frames, host rows async", fp, True, True)]:
    run(fr, ho, ac)
    print(f"{name:32s} {run(fr, ho, ac):.3f} ms/pair")

rows_pinB = [torch.empty((N, 7), dtype=torch.float32).pin_memory() for _ in range(B)]
cnt_pin = torch.zeros(B, dtype=torch.int32).pin_memory()
cnt_dev = torch.zeros(B, dtype=torch.int32, device=dev)
def run_submit(frames, host_out):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for b in range(B):
        if host_out:
            mr.submit_main_frame(r, frames[b], sc.cameras[b], [frames[b + 1]], [sc.cameras[b + 1]], out=rows_pinB[b], out_count=cnt_pin[b:b+1])
        else:
            mr.submit_main_frame(r, frames[b], sc.cameras[b], [frames[b + 1]], [sc.cameras[b + 1]], out=rows_dev, out_count=cnt_dev[b:b+1])
    r.ctx.synchronize(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / B * 1e3
for name, fr, ho in [("submit dev frames, dev rows", fd, False), ("submit dev frames, host rows", fd, True), ("submit host frames, host rows", fp, True)]:
    run_submit(fr, ho)
    print(f"{name:32s} {run_submit(fr, ho):.3f} ms/pair")
