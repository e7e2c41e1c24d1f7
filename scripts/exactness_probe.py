import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import synth
from oracle.pipeline import process_main_frame
from oracle.render import RenderOracle
for (W, H, S, step) in [(320, 240, 1, 0.12), (640, 480, 2, 0.05), (640, 480, 4, 0.01), (960, 540, 1, 0.006)]:
    n = 2 * S + 1
    sc = synth.make_scene(W, H, n, seed=W + S, step=step, mesh_err=0.03, mesh_res=14)
    frames = sc.frames(); fa = n // 2; sides = [i for i in range(n) if i != fa][:S]
    ro = RenderOracle(W, H); ro.loadMesh(sc.vertices, sc.faces)
    ref = process_main_frame(ro, frames, sc.cameras, fa, sides)
    r = mr.Render(W, H, ctx=mr.api.Context(W, H)); r.loadMesh(sc.vertices, sc.faces)
    got = mr.process_main_frame(r, frames[fa], sc.cameras[fa], [frames[s] for s in sides], [sc.cameras[s] for s in sides])
    same_xyzw = (got[:, :4].view(np.uint32) == ref[:, :4].view(np.uint32)).all(1)
    same_all = (got.view(np.uint32) == ref.view(np.uint32)).all(1)
    nan = np.isnan(ref).any(1)
    print(f"{W}x{H} S={S} step={step}: rows {len(ref)}, xyzw bit-identical {same_xyzw.mean():.6f}, all 7 floats bit-identical {same_all.mean():.6f}, nan rows {nan.sum()}")
