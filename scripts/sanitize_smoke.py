import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import synth
for (W, H, S) in [(160, 128, 2), (101, 67, 1)]:
    sc = synth.make_scene(W, H, S + 1, step=0.15, mesh_err=0.03, mesh_res=6)
    frames = sc.frames()
    r = mr.Render(W, H, ctx=mr.api.Context(W, H)); r.loadMesh(sc.vertices, sc.faces)
    sides = list(range(1, S + 1))
    tri = mr.process_main_frame(r, frames[0], sc.cameras[0], [frames[s] for s in sides], [sc.cameras[s] for s in sides])
    fl = mr.calculateFlow(frames[0], frames[1], useFarneback=True)
    q = r.depthSamples(sc.cameras[:1], np.array([[3, 5]]), np.array([[7, 9]]))
    print(W, H, S, tri.shape, float(np.abs(fl).max()), q)
