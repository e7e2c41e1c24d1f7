import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import synth
W, H, B = 1920, 1080, 32
N = W * H
sc = synth.make_scene(W, H, 300, step=0.006, mesh_err=0.02)
dev = torch.device("cuda", 0)
fd = [sc.frame_torch(i, dev).contiguous() for i in range(B + 1)]
for nctx in (1, 2, 3):
    rs = []
    for _ in range(nctx):
        r = mr.Render(W, H, ctx=mr.api.Context(W, H, 0)); r.loadMesh(sc.vertices, sc.faces); rs.append(r)
    rows = [torch.empty((N, 7), dtype=torch.float32, device=dev) for _ in range(nctx)]
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    def run():
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for b in range(B):
            k = b % nctx
            mr.submit_main_frame(rs[k], fd[b], sc.cameras[b], [fd[b + 1]], [sc.cameras[b + 1]], out=rows[k], out_count=cnt[b:b+1])
        for r in rs: r.ctx.synchronize()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / B * 1e3
    run()
    print(f"contexts={nctx}: {run():.3f} ms/pair  -> {N / run() / 1e3:.0f} Mpix/s")
    del rs
