"""Rasteriser cost against mesh size: Render::depth at 1080p for proxy meshes of
10^3 .. 10^6 faces (the reference's second outer iteration renders a Poisson mesh), plus one
viewer INSIDE the scene (triangles crossing the camera plane get a full-screen bounding box).
Run on a GPU box:  python scripts/raster_scaling_probe.py [W H]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mesh_reconstruction_b200 as mr  # noqa: E402
from mesh_reconstruction_b200 import synth  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
dev = torch.device("cuda", 0)
print(f"# {W}x{H}; ms per Render::depth call (CUDA events, 20 calls after 3 warm-ups), device-resident output")
for res in [int(v) for v in os.environ.get("RASTER_PROBE_RES", "24,71,224,500,707").split(",")]:
    sc = synth.make_scene(W, H, 3, mesh_res=res)
    r = mr.Render(W, H, ctx=mr.api.Context(W, H, 0))
    r.loadMesh(sc.vertices, sc.faces)
    out = torch.empty((H, W), dtype=torch.float32, device=dev)
    st = torch.cuda.ExternalStream(r.ctx.stream, device=dev)
    cams = {"outside": sc.cameras[1]}
    # a viewer sitting inside the surface's bounding box, looking along it: many triangles cross the w = 0 plane
    c = synth.look_at((0.0, -1.0, 0.02), (0.0, 1.0, 0.0))
    cams["inside"] = (synth.perspective_matrix(sc.fov, W / H, 0.001, 10.0) @ np.linalg.inv(c)).astype(np.float32)
    line = [f"faces {len(sc.faces):8d}"]
    for name, P in cams.items():
        for _ in range(3):
            r.depth(P, out=out)
        r.ctx.synchronize()
        with torch.cuda.stream(st):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(20):
                r.depth(P, out=out)
            e1.record(st)
        r.ctx.synchronize()
        t0 = time.perf_counter()
        cov = float((out != 1.0).float().mean())
        line.append(f"{name}: {e0.elapsed_time(e1) / 20:8.3f} ms (covered {100 * cov:5.1f} %)")
    print("  ".join(line), flush=True)
    r.ctx.close()
