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
    bad = np.zeros((48 * 2, 64 * 3, 3), np.uint8)          # different factors in x and y
    with pytest.raises(mr.MeshReconError):
        mr.api.ingest_frame(ctx, bad)
    with pytest.raises(mr.MeshReconError):
        mr.api.ingest_frame(ctx, np.zeros((72, 96, 3), np.uint8))      # factor 1.5
    with pytest.raises(mr.MeshReconError):
        ctx.check(ctx.lib.mr_set_gray_shift(ctx.h, 13))
