"""Repro: filterPoints on the rows of one 4K main frame of the bench scene (z0 = -2.7 by default)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import synth
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
z0 = float(sys.argv[3]) if len(sys.argv) > 3 else -2.7
S = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = torch.device("cuda", 0)
sc = synth.make_scene(W, H, 600, step=0.006, mesh_err=0.02, z0=z0)
fa = 2
sides = [fa + 1] if S == 1 else [fa - 2, fa - 1, fa + 1, fa + 2]
fr = {i: sc.frame_torch(i, dev).contiguous() for i in [fa] + sides}
r = mr.Render(W, H, ctx=mr.api.Context(W, H, 0))
r.loadMesh(sc.vertices, sc.faces)
rows = torch.empty((W * H, 7), dtype=torch.float32, device=dev)
cnt = torch.zeros(1, dtype=torch.int32, device=dev)
mr.submit_main_frame(r, fr[fa], sc.cameras[fa], [fr[s] for s in sides], [sc.cameras[s] for s in sides], out=rows, out_count=cnt)
r.ctx.synchronize()
m0 = int(cnt[0])
rows0 = rows[:m0].contiguous()
fin = torch.isfinite(rows0).all(1)
rows0 = rows0[fin].contiguous()
d3 = rows0[:, :3] / rows0[:, 3:4]
spacing = float((d3[:, 0].max() - d3[:, 0].min())) / W
radius = (3.0 * spacing) ** 2
print("points", len(rows0), "radius", radius, "bbox", d3.min(0).values.tolist(), d3.max(0).values.tolist(), flush=True)
out_rows = torch.empty_like(rows0)
os.environ["MR_FILTER_TIMING"] = "1"
kept, keep = mr.filter_rows(rows0, radius, ctx=r.ctx, out=out_rows)
torch.cuda.synchronize()
print("survivors", len(keep), mr.api.filter_info(r.ctx), flush=True)
