"""Where does e2e lose against resident?  2 contexts, submit API, toggling host/device inputs and outputs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import synth
W, H, B, STEPS = 1920, 1080, 8, 6
N = W * H
nctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc = synth.make_scene(W, H, 300, step=0.006, mesh_err=0.02)
dev = torch.device("cuda", 0)
fd = [sc.frame_torch(i, dev).contiguous() for i in range(B + 1)]
fp = [f.cpu().pin_memory() for f in fd]
rs = []
for _ in range(nctx):
    r = mr.Render(W, H, ctx=mr.api.Context(W, H, 0)); r.loadMesh(sc.vertices, sc.faces); rs.append(r)
rows_dev = [torch.empty((N, 7), dtype=torch.float32, device=dev) for _ in range(B)]
cnt_dev = torch.zeros(B, dtype=torch.int32, device=dev)
rows_pin = [[torch.empty((N, 7), dtype=torch.float32).pin_memory() for _ in range(B)] for _ in range(2)]
cnt_pin = [torch.zeros(B, dtype=torch.int32).pin_memory() for _ in range(2)]
per = [len(range(c, B, nctx)) for c in range(nctx)]
def run(frames, host_out):
    for r in rs: r.ctx.synchronize()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for s in range(STEPS):
        k = s & 1
        for b in range(B):
            if host_out:
                mr.submit_main_frame(rs[b % nctx], frames[b], sc.cameras[b], [frames[b + 1]], [sc.cameras[b + 1]], out=rows_pin[k][b], out_count=cnt_pin[k][b:b+1])
            else:
                mr.submit_main_frame(rs[b % nctx], frames[b], sc.cameras[b], [frames[b + 1]], [sc.cameras[b + 1]], out=rows_dev[b], out_count=cnt_dev[b:b+1])
        if host_out:
            for c, r in enumerate(rs): r.ctx.wait_copies_until(per[c])
    for r in rs: r.ctx.synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / (B * STEPS) * 1e3
for name, fr, ho in [("dev frames, dev rows", fd, False), ("host frames, dev rows", fp, False), ("dev frames, host rows", fd, True), ("host frames, host rows", fp, True)]:
    run(fr, ho)
    print(f"nctx={nctx} {name:28s} {run(fr, ho):.3f} ms/pair")

# Is it contention or dependencies?  Resident run with an UNRELATED D2H stream moving the same volume in the background.
bg_src = torch.empty((N, 7), dtype=torch.float32, device=dev)
bg_dst = [torch.empty((N, 7), dtype=torch.float32).pin_memory() for _ in range(2)]
bg_stream = torch.cuda.Stream()
def run_bg():
    for r in rs: r.ctx.synchronize()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for s in range(STEPS):
        for b in range(B):
            mr.submit_main_frame(rs[b % nctx], fd[b], sc.cameras[b], [fd[b + 1]], [sc.cameras[b + 1]], out=rows_dev[b], out_count=cnt_dev[b:b+1])
            with torch.cuda.stream(bg_stream):
                bg_dst[b & 1].copy_(bg_src, non_blocking=True)
    for r in rs: r.ctx.synchronize()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / (B * STEPS) * 1e3, (time.perf_counter() - t0) / (B * STEPS) * 1e3
run_bg()
print("resident + unrelated background D2H (compute done, all done): %.3f %.3f ms/pair" % run_bg())
