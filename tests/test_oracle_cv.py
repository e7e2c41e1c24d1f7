"""Pins the oracle's restated OpenCV primitives against the real OpenCV binary (cv2).
CPU-only.  These are the known-answer checks available for the third-party arithmetic
on the path: the reference's own tests pin nothing."""
import cv2
import numpy as np
import pytest

from oracle import cvprims as P
from oracle import native

f32 = np.float32


def _pair(H, W, seed=0, shift=(0.3, -0.2)):
    rng = np.random.default_rng(seed)
    base = cv2.resize(rng.normal(128, 45, (H // 8 + 2, W // 8 + 2)).astype(f32), (W + 8, H + 8),
                      interpolation=cv2.INTER_CUBIC)
    i0 = np.clip(base[4:4 + H, 4:4 + W], 0, 255).astype(np.uint8)
    M = np.float32([[1, 0, shift[0]], [0, 1, shift[1]]])
    i1 = np.clip(cv2.warpAffine(base, M, (W + 8, H + 8), flags=cv2.INTER_CUBIC)[4:4 + H, 4:4 + W], 0, 255).astype(np.uint8)
    return i0, i1


@pytest.mark.parametrize("H,W,seed", [(61, 83, 0), (64, 64, 1), (37, 50, 2), (120, 161, 3)])
def test_vr_restatement_bit_exact(H, W, seed):
    i0, i1 = _pair(H, W, seed)
    ref = cv2.VariationalRefinement_create().calc(i0, i1, np.zeros((H, W, 2), f32))
    mine = P.variational_refinement(i0, i1)
    assert np.array_equal(ref, mine)


def test_vr_restatement_large_motion_and_flat():
    i0, i1 = _pair(48, 64, 5, shift=(2.5, 1.5))
    ref = cv2.VariationalRefinement_create().calc(i0, i1, np.zeros((48, 64, 2), f32))
    assert np.array_equal(ref, P.variational_refinement(i0, i1))
    flat = np.full((20, 30), 77, np.uint8)
    ref = cv2.VariationalRefinement_create().calc(flat, flat, np.zeros((20, 30, 2), f32))
    assert np.array_equal(ref, P.variational_refinement(flat, flat))


@pytest.mark.parametrize("H,W", [(61, 83), (120, 160)])
def test_cubic_remap_bit_exact(H, W):
    rng = np.random.default_rng(0)
    img = (rng.random((H, W)) * 255).astype(np.uint8)
    fl = (rng.normal(size=(H, W, 2)) * 3).astype(f32)
    fl[0, 0] = (-50, -50)  # far outside -> 0
    m = np.stack([fl[..., 0] + np.arange(W, dtype=f32)[None], fl[..., 1] + np.arange(H, dtype=f32)[:, None]], -1)
    ref = cv2.remap(img, m, None, cv2.INTER_CUBIC)
    assert np.array_equal(ref, P.flow_remap(fl, img))


@pytest.mark.parametrize("H,W", [(61, 83), (64, 64), (135, 240), (480, 640), (37, 50), (5, 7), (9, 3), (17, 30), (3, 4), (2, 3), (3, 2), (2, 2)])
def test_pyramids_and_compare_bit_exact(H, W):
    rng = np.random.default_rng(1)
    a = (rng.random((H, W)) * 255).astype(f32)
    assert np.array_equal(cv2.pyrDown(a), P.pyr_down(a))
    d = P.pyr_down(a)
    h, w = d.shape
    for (HH, WW) in {(2 * h, 2 * w), (H, W)}:
        assert np.array_equal(cv2.pyrUp(d, dstsize=(WW, HH)), P.pyr_up(d, (HH, WW))), (HH, WW)
    ua = a.astype(np.uint8)
    ub = np.clip(ua.astype(int) + rng.integers(-6, 6, (H, W)), 0, 255).astype(np.uint8)
    ref = P.compare(ua, ub, cv2.pyrDown, lambda s, sz: cv2.pyrUp(s, dstsize=(sz[1], sz[0])))
    assert np.array_equal(ref, P.compare(ua, ub))


def test_compare_level_count():
    assert len(P.compare_levels(480, 640)) == 9
    assert len(P.compare_levels(1080, 1920)) == 11 - 1 or len(P.compare_levels(1080, 1920)) == 10
    assert len(P.compare_levels(2160, 3840)) == 11


def test_sobel_bit_exact():
    rng = np.random.default_rng(2)
    for W in list(range(3, 40)) + [50, 83, 96, 101, 640, 641]:
        for H in (3, 40):
            a = (rng.normal(size=(H, W)) * 0.3).astype(f32)
            g = P.sobel_gradient(a)
            assert np.array_equal(g[..., 0], cv2.Sobel(a, cv2.CV_32F, 1, 0)), (H, W)
            assert np.array_equal(g[..., 1], cv2.Sobel(a, cv2.CV_32F, 0, 1)), (H, W)


def test_small_linalg_bit_exact():
    rng = np.random.default_rng(3)
    out = np.empty(16, f32)
    for _ in range(200):
        m = rng.normal(size=(4, 4)).astype(f32)
        ref = cv2.invert(m)[1]
        assert np.array_equal(ref, P.lu_inv4(m))
        native.lib().orc_lu_inv4(np.ascontiguousarray(m.ravel()), out)
        assert np.array_equal(ref.ravel(), out)
        s = rng.normal(size=(2, 2)).astype(f32)
        s = (s @ s.T).astype(f32)
        assert np.array_equal(cv2.invert(s)[1], P.inv2(s))
    for shp in [((4, 4), (4, 4)), ((4, 4), (4, 1)), ((2, 3), (3, 3)), ((2, 3), (3, 2)), ((2, 4), (4, 1)),
                ((1, 4), (4, 4)), ((2, 2), (2, 1))]:
        for _ in range(50):
            a = rng.normal(size=shp[0]).astype(f32)
            b = rng.normal(size=shp[1]).astype(f32)
            assert np.array_equal(cv2.gemm(a, b, 1.0, None, 0.0), P.gemm_f32(a, b)), shp


def test_pca_matches_cv2():
    """orc_pca_normal restates cv::PCA (cv::reduce mean, cv::mulTransposed covariance, cv::eigen = JacobiImpl_ with
    OpenCV's own hypot): BIT-IDENTICAL to cv2.PCACompute2 -- smallest eigenvector and all eigenvalues."""
    rng = np.random.default_rng(4)
    L = native.lib()
    worst = 0.0
    for t in range(200):
        K = int(rng.integers(3, 441))
        basis = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        sig = np.array([1.0, 0.6, 0.02]) * 10 ** rng.uniform(-2, 0)
        pts = ((rng.normal(size=(K, 3)) * sig) @ basis.T + rng.normal(size=3) * 3).astype(f32)
        mean, evec, evals = cv2.PCACompute2(pts, None)
        n = np.empty(3, f32)
        ev = np.empty(3, f32)
        L.orc_pca_normal(np.ascontiguousarray(pts), K, n, ev)
        assert np.array_equal(n, evec[2]), (t, K, n, evec[2])
        assert np.array_equal(ev, evals.ravel()), (t, K, ev, evals.ravel())
        # and stage by stage: mean (cv::reduce), covariance (cv::calcCovarMatrix -> mulTransposed)
        cov, mean2 = cv2.calcCovarMatrix(pts, None, cv2.COVAR_NORMAL | cv2.COVAR_SCALE | cv2.COVAR_ROWS, ctype=cv2.CV_32F)
        assert np.array_equal(mean2.ravel().astype(f32), mean.ravel())
    assert worst == 0.0


def test_camera_center_matches_decompose_and_yaml_convention():
    from mesh_reconstruction_b200 import synth
    sc = synth.make_scene(64, 48, 5)
    L = native.lib()
    for i in range(5):
        P4 = sc.cameras[i]
        T = cv2.decomposeProjectionMatrix(np.ascontiguousarray(P4[[0, 1, 3]]))[2]
        c_ref = (T[:3] / T[3]).ravel()
        c = np.empty(3, f32)
        L.orc_camera_center(np.ascontiguousarray(P4.ravel()), c)
        assert np.allclose(c, c_ref, rtol=2e-5, atol=2e-6)
        assert np.allclose(c, sc.cam2world[i][:3, 3], rtol=1e-4, atol=1e-5)


def test_float_reciprocal_equals_double_rounded_reciprocal():
    """`Mat /= s` multiplies by (float)(1.0/(double)s).  The CUDA kernels use the correctly rounded
    float reciprocal instead; the two are identical for every float (double rounding is harmless
    for reciprocals).  Exhaustive over all 2^23 mantissas of one binade plus random exponents."""
    m = np.arange(1 << 23, dtype=np.uint32) | np.uint32(0x3F800000)
    s = m.view(np.float32)
    a = (1.0 / s.astype(np.float64)).astype(np.float32)           # reference form
    # correctly rounded float reciprocal, computed exactly with integers: round(2^47 / M) etc. -> use float division
    b = np.float32(1.0) / s                                         # IEEE float division, correctly rounded
    assert np.array_equal(a, b)
    rng = np.random.default_rng(0)
    s2 = (rng.integers(0x00800000, 0x7F000000, 1 << 20, dtype=np.uint32)).view(np.float32)
    with np.errstate(all="ignore"):
        assert np.array_equal((1.0 / s2.astype(np.float64)).astype(np.float32), np.float32(1.0) / s2)


def test_farneback_restatement_matches_cv2():
    """oracle/farneback_np.py (the formulas csrc/farneback.cu is written from) vs the cv2 binary."""
    from oracle import farneback_np as fb
    rng = np.random.default_rng(0)
    H, W = 96, 128
    base = cv2.resize(rng.normal(128, 45, (H // 8 + 3, W // 8 + 3)).astype(f32), (W + 16, H + 16), interpolation=cv2.INTER_CUBIC)
    p = np.clip(base[8:8 + H, 8:8 + W], 0, 255).astype(np.uint8)
    M = np.float32([[1, 0, 1.3], [0, 1, -0.7]])
    n = np.clip(cv2.warpAffine(base, M, (W + 16, H + 16), flags=cv2.INTER_CUBIC)[8:8 + H, 8:8 + W], 0, 255).astype(np.uint8)
    ps = (H + W) / 1000.0
    ref = cv2.FarnebackOpticalFlow_create(10, 0.8, False, (H + W) // 100, 7, 5 if ps < 1.5 else 7, ps, 0).calc(p, n, None)
    mine = fb.farneback(p, n)
    assert np.abs(ref - mine).max() < 1e-4
    assert abs(np.median(ref[..., 0]) - 1.3) < 0.1 and abs(np.median(ref[..., 1]) + 0.7) < 0.1


@pytest.mark.parametrize("sw,sh,W,H", [(96, 72, 64, 48), (100, 70, 40, 28), (97, 61, 33, 20), (50, 50, 49, 49), (120, 54, 48, 36), (64, 90, 64, 36)])
def test_general_inter_area_restatement_bit_exact(sw, sh, W, H):
    """cv::resize INTER_AREA for the fractional shrink factors a `-s 1.5` run produces (configuration.cpp:160-163, 232-233)."""
    rng = np.random.default_rng(sw + W)
    src = rng.integers(0, 256, (sh, sw, 3)).astype(np.uint8)
    src[: sh // 3] = rng.integers(0, 2, (sh // 3, sw, 3)) * 255
    assert np.array_equal(P.resize_area_8u(src, W, H), cv2.resize(src, (W, H), interpolation=cv2.INTER_AREA))
    assert np.array_equal(P.resize_area_8u(src[..., 0], W, H), cv2.resize(np.ascontiguousarray(src[..., 0]), (W, H), interpolation=cv2.INTER_AREA))
