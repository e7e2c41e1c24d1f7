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
