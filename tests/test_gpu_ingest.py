"""GPU parity of mr_ingest_frame (configuration.cpp:226-245: cv::resize INTER_AREA + cv::cvtColor BGR2GRAY) against the
OpenCV binary itself (cv2): byte-exact."""
import numpy as np
import pytest

import mesh_reconstruction_b200 as mr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H,f", [(640, 480, 1), (641, 479, 1), (7, 5, 1), (1920, 1080, 1), (640, 360, 2), (640, 360, 3), (320, 240, 4), (33, 21, 5), (3840, 2160, 1)])
def test_ingest_matches_cv2(W, H, f):
    import cv2
    import torch
    rng = np.random.default_rng(W + f)
    bgr = rng.integers(0, 256, (H * f, W * f, 3)).astype(np.uint8)
    bgr[: H * f // 3] = rng.integers(0, 4, (H * f // 3, W * f, 3)) * 85        # saturated / flat areas: rounding ties
    ref = cv2.cvtColor(cv2.resize(bgr, (W, H), interpolation=cv2.INTER_AREA) if f > 1 else bgr, cv2.COLOR_BGR2GRAY)
    ctx = mr.api.Context(W, H)
    got = mr.api.ingest_frame(ctx, bgr)
    assert np.array_equal(got, ref), (np.abs(got.astype(int) - ref.astype(int)).max(), (got != ref).mean())
    # pinned host frame in, device gray frame out (what a streaming decoder would do), asynchronous
    pin = torch.from_numpy(bgr).pin_memory()
    dev = torch.empty((H, W), dtype=torch.uint8, device="cuda")
    mr.api.ingest_frame(ctx, pin, out=dev)
    ctx.synchronize()
    assert np.array_equal(dev.cpu().numpy(), ref)
    # the 14-bit coefficients of OpenCV 3.0 - 3.4.5
    ctx.check(ctx.lib.mr_set_gray_shift(ctx.h, 14))
    src = (cv2.resize(bgr, (W, H), interpolation=cv2.INTER_AREA) if f > 1 else bgr).astype(np.int64)
    ref14 = ((src[..., 0] * 1868 + src[..., 1] * 9617 + src[..., 2] * 4899 + (1 << 13)) >> 14).astype(np.uint8)
    assert np.array_equal(mr.api.ingest_frame(ctx, bgr), ref14)


def test_ingest_argument_errors():
    ctx = mr.api.Context(64, 48)
    with pytest.raises(mr.MeshReconError):
        mr.api.ingest_frame(ctx, np.zeros((40, 64, 3), np.uint8))      # smaller than the render size: INTER_AREA only shrinks here
    with pytest.raises(mr.MeshReconError):
        mr.api.ingest_frame(ctx, np.zeros((48, 63, 3), np.uint8))
    with pytest.raises(mr.MeshReconError):
        ctx.check(ctx.lib.mr_set_gray_shift(ctx.h, 13))


@pytest.mark.parametrize("sw,sh,W,H", [(960, 540, 640, 360), (1000, 700, 400, 280), (97, 61, 33, 20), (1920, 1080, 1280, 720), (50, 50, 49, 49),
                                       (120, 54, 48, 36), (64, 90, 64, 36), (192, 144, 64, 72), (64, 96, 64, 48), (3840, 2160, 1536, 864)])
def test_ingest_fractional_and_mixed_factors_match_cv2(sw, sh, W, H):
    """`-s 1.5`, `-s 2.5` ... (configuration.cpp:160-163): cv::resize's GENERAL area path (float cell weights), and integer
    factors that differ in x and y (OpenCV's fast path with fx != fy): byte-exact against the cv2 binary, gray and exposure mix."""
    import cv2
    rng = np.random.default_rng(sw * 3 + H)
    bgr = rng.integers(0, 256, (sh, sw, 3)).astype(np.uint8)
    bgr[: sh // 3] = rng.integers(0, 4, (sh // 3, sw, 3)) * 85
    small = cv2.resize(bgr, (W, H), interpolation=cv2.INTER_AREA)
    ctx = mr.api.Context(W, H)
    got = mr.api.ingest_frame(ctx, bgr)
    ref = cv2.cvtColor(small, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(got, ref), (np.abs(got.astype(int) - ref.astype(int)).max(), (got != ref).mean())
    e = (0.4, 0.35, 0.3)
    assert np.array_equal(mr.api.ingest_frame(ctx, bgr, exposure=e), _exposure_ref(small, e))


def _exposure_ref(bgr, e):
    """configuration.cpp:417-425 with the OpenCV binary: frame = zeros; frame += channel[c] * exposure[c] (8-bit Mat
    arithmetic: convertTo with a float scale, then a saturating add).  cv2.convertScaleAbs == convertTo for weights >= 0;
    a negative weight saturates every product to 0."""
    import cv2
    ref = np.zeros(bgr.shape[:2], np.uint8)
    for c in range(3):
        t = cv2.convertScaleAbs(np.ascontiguousarray(bgr[..., c]), alpha=float(np.float32(e[c]))) if e[c] >= 0 else np.zeros_like(ref)
        ref = cv2.add(ref, t)
    return ref


@pytest.mark.parametrize("W,H,f,e", [(640, 480, 1, (0.31, 0.36, 0.29)), (321, 243, 1, (0.5, 0.5, 0.5)), (640, 360, 2, (0.9, 0.7, 0.2)),
                                     (160, 120, 3, (0.114, 0.587, 0.299)), (64, 48, 1, (1.7, -0.4, 0.3)), (64, 48, 1, (0.0, 0.0, 3.0e7))])
def test_ingest_with_exposure_matches_cv2(W, H, f, e):
    """Configuration::estimateExposure's frame normalisation (its estimate is host work; the per-pixel mix is the device part)."""
    import cv2
    rng = np.random.default_rng(W * 7 + f)
    bgr = rng.integers(0, 256, (H * f, W * f, 3)).astype(np.uint8)
    bgr[: H * f // 4] = rng.integers(0, 4, (H * f // 4, W * f, 3)) * 85
    small = cv2.resize(bgr, (W, H), interpolation=cv2.INTER_AREA) if f > 1 else bgr
    ref = _exposure_ref(small, e)
    if e[2] > 1e6:        # 255 * 3e7 >= 2^31: cvRound gives the integer indefinite, which saturates to 0 (x86 cvtps2dq + packs)
        ref = np.where(small[..., 2].astype(np.float32) * np.float32(e[2]) >= 2147483648.0, 0, ref).astype(np.uint8)
    ctx = mr.api.Context(W, H)
    got = mr.api.ingest_frame(ctx, bgr, exposure=e)
    assert np.array_equal(got, ref), (np.abs(got.astype(int) - ref.astype(int)).max(), (got != ref).mean())
