"""Generates the committed golden vectors from the CPU oracle.

    python tests/golden/make_golden.py

The reference has no golden vectors for this path and cannot be
built here, so the vectors come from the oracle: OpenCV-owned stages are executed by
the cv2 binary, reference-owned stages by oracle/recon_oracle.c.  Inputs are fully
determined by the seeds below; the .npz files are small enough to commit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mesh_reconstruction_b200 import synth  # noqa: E402
from oracle.pipeline import process_main_frame  # noqa: E402
from oracle.render import RenderOracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def scene_case(name, W, H, n, fa, sides, **kw):
    sc = synth.make_scene(W, H, n, **kw)
    frames = sc.frames()
    r = RenderOracle(W, H)
    r.loadMesh(sc.vertices, sc.faces)
    tri, inter = process_main_frame(r, frames, sc.cameras, fa, sides, keep=True)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        W=W, H=H, fa=fa, sides=np.asarray(sides), frames=np.stack(frames), cameras=sc.cameras,
        vertices=sc.vertices, faces=sc.faces, scale=sc.scale,
        depth0=inter["depth0"], depth=inter["depth"], projected=np.stack(inter["projected"]),
        mixed=np.stack(inter["mixed"]), flows=np.stack(inter["flows"]), tri=tri)
    print(name, "points", tri.shape, "nan rows", int(np.isnan(tri).any(1).sum()))


def glx_case():
    """The reference's only fixture on this path: render_glx.cpp:407-410."""
    W, H = 160, 120
    r = RenderOracle(W, H)
    r.loadMesh(synth.TEST_GLX_POINTS, synth.TEST_GLX_FACES)
    rng = np.random.default_rng(7)
    yy, xx = np.mgrid[0:H, 0:W]
    grid = (((xx // 8 + yy // 8) % 2) * 160 + 40 + rng.integers(0, 16, (H, W))).astype(np.uint8)
    depth = r.depth(synth.TEST_GLX_MVP)
    proj = r.projected(synth.TEST_GLX_MVP, grid, synth.TEST_GLX_SIDE_MVP)
    shadow = r.shadow_map(synth.TEST_GLX_SIDE_MVP)
    np.savez_compressed(os.path.join(HERE, "test_glx.npz"), W=W, H=H, grid=grid, depth=depth, projected=proj,
                        shadow=shadow)
    print("test_glx depth range", depth[depth != 1].min(), depth[depth != 1].max(), "covered", (depth != 1).mean())


if __name__ == "__main__":
    scene_case("scene_s2_96x72", 96, 72, 4, 1, [0, 2], step=0.2, mesh_err=0.03, mesh_res=8)
    scene_case("scene_s1_128x96", 128, 96, 3, 1, [2], step=0.15, mesh_err=0.03, mesh_res=12, seed=3)
    glx_case()
