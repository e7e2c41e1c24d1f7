"""The oracle against a cv2-level transliteration of util.cpp:62-329 (tests/golden/make_cv2_transliteration.py): every
cv::Mat expression of triangulatePixels / triangulatePixel evaluated by the real OpenCV binary in the reference's
statement order.  oracle/recon_oracle.c must reproduce its committed output BIT FOR BIT -- points, pdf-scaled normals,
K < 3 fallback rows and NaN rows -- which pins the evaluation rules the restatement (and the CUDA path, which shares
them) had only read off the OpenCV sources."""
import os

import numpy as np

from oracle.tri import triangulate_pixels

HERE = os.path.dirname(os.path.abspath(__file__))


def _same(a, b):
    return a.shape == b.shape and bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


def _load():
    g = np.load(os.path.join(HERE, "golden", "scene_s2_96x72.npz"))
    t = np.load(os.path.join(HERE, "golden", "cv2_translit_s2_96x72.npz"))
    fa, sides = int(g["fa"]), [int(s) for s in g["sides"]]
    return g, t, fa, sides


def test_oracle_reproduces_cv2_transliteration():
    g, t, fa, sides = _load()
    cams = g["cameras"]
    ref, ev = triangulate_pixels(list(g["flows"]), cams[fa], [cams[s] for s in sides], g["depth"], return_evals=True)
    assert len(ref) > 6000 and _same(ref, t["tri"])
    assert np.array_equal(ref, g["tri"], equal_nan=True)                      # and the older golden of the oracle itself
    # S = 1 with isolated pixels (K < 3 fallback) and a zero-variance pixel (NaN row)
    ref1, ev1 = triangulate_pixels([t["flow1"]], cams[fa], [cams[sides[0]]], t["depth1"], return_evals=True)
    assert set(int(k) for k in ev1[:, 3]) >= {1, 2} and np.isnan(ref1).any()
    assert _same(ref1, t["tri1"])


def test_oracle_reproduces_cv2_transliteration_s4_and_s3():
    """Four side cameras make `projectionW * k` (util.cpp:105) a 4x4 cv::gemm -- OpenCV's FLOAT small-matrix kernel -- while
    any other S takes the generic kernel with double accumulators.  The transliteration (cv2.gemm itself) pins both."""
    t = np.load(os.path.join(HERE, "golden", "cv2_translit_s34_40x30.npz"))
    cams, fa = t["cameras"], int(t["fa"])
    for name in ("s4", "s3"):
        sides = [int(s) for s in t["sides_" + name]]
        ref = triangulate_pixels(list(t["flows_" + name]), cams[fa], [cams[s] for s in sides], t["depth_" + name])
        assert len(ref) > 900 and _same(ref, t["tri_" + name]), name


def test_live_transliteration_on_random_small_cases():
    """Beyond the committed goldens: the cv2-level transliteration is run LIVE on small random inputs (random sub-pixel flows,
    variances incl. exact zeros, holes in the depth map, S = 1, 2, 4, 5) and the oracle must agree bit for bit every time."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("make_cv2_transliteration", os.path.join(HERE, "golden", "make_cv2_transliteration.py"))
    tl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tl)
    from mesh_reconstruction_b200 import synth
    from oracle.render import RenderOracle
    W, H = 14, 10
    for seed, S in [(0, 1), (1, 2), (2, 4), (3, 5), (4, 4)]:
        rng = np.random.default_rng(100 + seed)
        sc = synth.make_scene(W, H, S + 1, seed=seed, step=0.1, mesh_res=4, z0=(-2.7 if seed % 2 else 0.0))
        ro = RenderOracle(W, H)
        ro.loadMesh(sc.vertices, sc.faces)
        depth = ro.depth(sc.cameras[0]).copy()
        depth[rng.random((H, W)) < 0.15] = 1.0                                   # holes: goodSample fallbacks, K < 3 windows
        flows = []
        for _ in range(S):
            f = np.zeros((H, W, 4), np.float32)
            f[..., :2] = rng.normal(0, 0.4, (H, W, 2))
            f[..., 2] = rng.uniform(0.05, 3.0, (H, W))
            f[..., 2][rng.random((H, W)) < 0.03] = 0.0                           # zero variance -> NaN rows (quirk C11)
            flows.append(f)
        cams = [sc.cameras[i].astype(np.float32) for i in range(S + 1)]
        exp = tl.triangulate_pixels(flows, cams[0], cams[1:], depth)
        got = triangulate_pixels(flows, cams[0], cams[1:], depth)
        assert len(exp) > 40 and _same(got, exp), (seed, S, got.shape, exp.shape)
