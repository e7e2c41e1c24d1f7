"""Rasteriser work split (small / big triangles, tight boxes for triangles crossing the camera plane, row spans) and the
ray-query form of the chooseCameras depth shots: results must stay bit-identical to the oracle's whole-box walk
(oracle/recon_oracle.c orc_raster), whatever the triangle soup looks like."""
import numpy as np
import pytest

import mesh_reconstruction_b200 as mr
from mesh_reconstruction_b200 import synth
from tests.test_gpu_parity_r2 import face_camera

pytestmark = pytest.mark.gpu
f32 = np.float32


def _random_camera(rng, W, H, near):
    eye = rng.uniform(-1.0, 1.0, 3)
    target = eye + rng.normal(size=3)
    c = synth.look_at(eye, target, up=rng.normal(size=3))
    return (synth.perspective_matrix(rng.uniform(0.4, 1.6), W / H, near, 10.0) @ np.linalg.inv(c)).astype(f32)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_soup_matches_oracle(seed):
    """Random triangles all around (and through) randomly placed cameras with a tiny near plane: slivers, triangles
    spanning the camera plane, back-facing and degenerate ones, on an odd-sized image."""
    from oracle.render import RenderOracle
    W, H = 97, 61
    rng = np.random.default_rng(seed)
    nt = 400
    centres = rng.uniform(-1.5, 1.5, (nt, 1, 3))
    size = np.exp(rng.uniform(np.log(0.01), np.log(3.0), (nt, 1, 1)))
    tri = centres + size * rng.normal(size=(nt, 3, 3))
    tri[::17, 2] = tri[::17, 1]                              # degenerate (zero area)
    tri[5::23, :, 2] = 0.25                                  # a few coplanar horizontal ones
    verts = np.concatenate([tri.reshape(-1, 3), np.ones((nt * 3, 1))], 1).astype(f32)
    faces = np.arange(nt * 3, dtype=np.int32).reshape(nt, 3)
    r = mr.Render(W, H, ctx=mr.api.Context(W, H))
    r.loadMesh(verts, faces)
    ro = RenderOracle(W, H)
    ro.loadMesh(verts, faces)
    hits = []
    for k in range(8):
        P = _random_camera(rng, W, H, 0.001 if k % 2 else 0.2)
        if k == 7:                                           # camera centre exactly ON a vertex (w == 0 there)
            eye = verts[3 * 11, :3].astype(np.float64)
            c = synth.look_at(eye, eye + np.array([0.3, 0.9, 0.1]))
            P = (synth.perspective_matrix(0.9, W / H, 0.001, 10.0) @ np.linalg.inv(c)).astype(f32)
        d, d_ref = r.depth(P), ro.depth(P)
        assert np.array_equal(d, d_ref), (seed, k, int((d != d_ref).sum()))
        hits.append(float((d_ref != 1.0).mean()))
    assert max(hits) > 0.3
    r.ctx.close()


def test_fine_mesh_face_cameras_and_queries():
    """A few thousand faces seen from viewers sitting on faces (heuristic.cpp:193-247): every triangle around the viewer
    crosses its camera plane.  Depth maps and the batched single-pixel queries (ray-query path) against the oracle."""
    from oracle.render import RenderOracle
    W, H = 200, 150
    sc = synth.make_scene(W, H, 2, mesh_res=40, mesh_err=0.05, amp=0.3)
    r = mr.Render(W, H, ctx=mr.api.Context(W, H))
    r.loadMesh(sc.vertices, sc.faces)
    ro = RenderOracle(W, H)
    ro.loadMesh(sc.vertices, sc.faces)
    rng = np.random.default_rng(5)
    cams, maps = [], []
    for face in rng.integers(0, len(sc.faces), 6):
        u1, u2 = rng.random(2)
        P = face_camera(sc.vertices, sc.faces, int(face), 10.0, 0.5, u1, u2)
        d_ref = ro.depth(P)
        d = r.depth(P)
        assert np.array_equal(d, d_ref), (int(face), int((d != d_ref).sum()))
        cams.append(P)
        maps.append(d_ref)
    assert max(float((m != 1.0).mean()) for m in maps) > 0.05
    n = 173
    rows = rng.integers(-2, H + 2, (len(cams), n)).astype(np.int32)
    cols = rng.integers(-2, W + 2, (len(cams), n)).astype(np.int32)
    cols[1, :4] = W
    got = r.depthSamples(np.stack(cams), rows, cols)
    for i, m in enumerate(maps):
        exp = np.full(n, 1.0, f32)
        ok = (rows[i] >= 0) & (rows[i] < H) & (cols[i] >= 0) & (cols[i] <= W)
        idx = np.minimum(rows[i].astype(np.int64) * W + cols[i], W * H - 1)
        exp[ok] = m.ravel()[idx[ok]]
        assert np.array_equal(got[i], exp), i
    r.ctx.close()


def test_depth_samples_many_queries_per_viewer_render_the_map():
    """More than 4096 queries per viewer take the render-then-index route; same answers as the ray queries."""
    W, H = 160, 120
    sc = synth.make_scene(W, H, 3, mesh_res=12)
    r = mr.Render(W, H, ctx=mr.api.Context(W, H))
    r.loadMesh(sc.vertices, sc.faces)
    rng = np.random.default_rng(9)
    n = 5000
    rows = rng.integers(0, H, (2, n)).astype(np.int32)
    cols = rng.integers(0, W + 1, (2, n)).astype(np.int32)
    got = r.depthSamples(sc.cameras[:2], rows, cols)
    few = r.depthSamples(sc.cameras[:2], rows[:, :100], cols[:, :100])
    assert np.array_equal(got[:, :100], few)
    for i in range(2):
        flat = r.depth(sc.cameras[i]).ravel()
        idx = np.minimum(rows[i].astype(np.int64) * W + cols[i], W * H - 1)
        assert np.array_equal(got[i], flat[idx])
    r.ctx.close()


def test_big_mesh_is_rasterised_at_all_scales():
    """10^5 faces at 1080p (the second outer iteration's Poisson mesh): the warp-per-triangle path against the
    oracle on a crop-free full frame (the oracle needs ~1 s), and the map of a coarser mesh of the SAME surface for sanity."""
    from oracle.render import RenderOracle
    W, H = 640, 360
    sc = synth.make_scene(W, H, 2, mesh_res=224, mesh_err=0.0)
    r = mr.Render(W, H, ctx=mr.api.Context(W, H))
    r.loadMesh(sc.vertices, sc.faces)
    ro = RenderOracle(W, H)
    ro.loadMesh(sc.vertices, sc.faces)
    d, d_ref = r.depth(sc.cameras[0]), ro.depth(sc.cameras[0])
    assert len(sc.faces) > 100000 and np.array_equal(d, d_ref)
    r.ctx.close()
