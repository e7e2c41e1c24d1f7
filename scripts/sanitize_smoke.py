import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import synth
for (W, H, S) in [(224, 208, 2), (101, 67, 1)]:   # 224x208: border AND interior VR tiles (TMA path); 101x67: plain loads
    sc = synth.make_scene(W, H, S + 1, step=0.15, mesh_err=0.03, mesh_res=6)
    frames = sc.frames()
    r = mr.Render(W, H, ctx=mr.api.Context(W, H)); r.loadMesh(sc.vertices, sc.faces)
    sides = list(range(1, S + 1))
    tri = mr.process_main_frame(r, frames[0], sc.cameras[0], [frames[s] for s in sides], [sc.cameras[s] for s in sides])
    fl = mr.calculateFlow(frames[0], frames[1], useFarneback=True)
    q = r.depthSamples(sc.cameras[:1], np.array([[3, 5]]), np.array([[7, 9]]))
    print(W, H, S, tri.shape, float(np.abs(fl).max()), q)

# submit path: plain first run, then CUDA-graph replays, rows by DMA into pinned buffers
import torch
W, H = 224, 208
sc = synth.make_scene(W, H, 4, step=0.15, mesh_err=0.03, mesh_res=6)
frames = [torch.from_numpy(f).pin_memory() for f in sc.frames()]
r = mr.Render(W, H, ctx=mr.api.Context(W, H)); r.loadMesh(sc.vertices, sc.faces)
rows = [torch.empty((W * H, 7), dtype=torch.float32).pin_memory() for _ in range(3)]
cnt = torch.zeros(3, dtype=torch.int32).pin_memory()
for i in range(3):
    mr.submit_main_frame(r, frames[i], sc.cameras[i], [frames[i + 1]], [sc.cameras[i + 1]], out=rows[i], out_count=cnt[i:i + 1])
r.ctx.wait_copies_until(0)
r.ctx.synchronize()
print("submit", cnt.tolist(), r.ctx.graph_launches)

# round 2: filterPoints (grid, segmented sorts, seqsum, thinning rounds), frame ingest, and a translated scene (integer-moment
# route of the normals kernel; the scenes above sit on the z = 0 plane and take the sample loops)
from tests.test_oracle_filter import cloud
p = cloud(6000, seed=3)
ctx = mr.api.Context(16, 16)
op, on, keep = mr.filterPoints(p, np.zeros((len(p), 3), np.float32), 0.004, ctx=ctx)
print("filter", len(keep), mr.api.filter_info(ctx))
bgr = np.random.default_rng(0).integers(0, 256, (96 * 3, 128 * 3, 3)).astype(np.uint8)
g = mr.api.ingest_frame(mr.api.Context(128, 96), bgr)
g1 = mr.api.ingest_frame(mr.api.Context(128 * 3, 96 * 3), bgr)
print("ingest", g.shape, int(g.sum()), g1.shape)
W, H = 224, 208
sc = synth.make_scene(W, H, 3, step=0.15, mesh_err=0.03, mesh_res=6)
T = np.eye(4); T[:3, 3] = (40.0, 25.0, -30.0)
verts = (sc.vertices.astype(np.float64) @ T.T).astype(np.float32)
cams = [(c.astype(np.float64) @ np.linalg.inv(T)).astype(np.float32) for c in sc.cameras]
fr = sc.frames()
r = mr.Render(W, H, ctx=mr.api.Context(W, H)); r.loadMesh(verts, sc.faces)
tri = mr.process_main_frame(r, fr[1], cams[1], [fr[0], fr[2]], [cams[0], cams[2]])
print("translated", tri.shape, r.ctx.normals_stats())
