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

# round 2, second half: the rasteriser's work split (small boxes / lists + persistent warps / row spans, a viewer INSIDE the mesh,
# the direct grid of small meshes), the ray-query depth shots, the general INTER_AREA ingest and the exposure mix
from tests.test_gpu_parity_r2 import face_camera
W, H = 200, 150
for res in (6, 48):                                   # 72 faces: direct grid; 4608 faces: lists
    sc = synth.make_scene(W, H, 2, mesh_res=res, mesh_err=0.05, amp=0.3)
    r = mr.Render(W, H, ctx=mr.api.Context(W, H)); r.loadMesh(sc.vertices, sc.faces)
    P = face_camera(sc.vertices, sc.faces, len(sc.faces) // 2, 10.0, 0.5, 0.3, 0.4)
    d_out, d_in = r.depth(sc.cameras[0]), r.depth(P)
    q = r.depthSamples(np.stack([sc.cameras[0], P]), np.tile(np.arange(0, 150, 10, dtype=np.int32), (2, 1)), np.tile(np.arange(0, 195, 13, dtype=np.int32), (2, 1)))
    big = r.depthSamples(sc.cameras[:1], np.random.default_rng(1).integers(0, H, (1, 5000)).astype(np.int32), np.random.default_rng(2).integers(0, W, (1, 5000)).astype(np.int32))
    print("raster", len(sc.faces), float((d_out != 1).mean()), float((d_in != 1).mean()), q.shape, big.shape)
bgr = np.random.default_rng(3).integers(0, 256, (108, 192, 3)).astype(np.uint8)
c = mr.api.Context(128, 72)
print("ingest 1.5x", int(mr.api.ingest_frame(c, bgr).sum()), int(mr.api.ingest_frame(c, bgr, exposure=(0.4, 0.3, 0.3)).sum()))
