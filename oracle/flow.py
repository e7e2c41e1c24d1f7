"""Oracle for ``calculateFlow`` (flow.cpp:19-42), ``flowRemap`` (util.cpp:390-403),
``compare`` (util.cpp:332-361) and ``imageGradient`` (util.cpp:465-479).
TEST INFRASTRUCTURE ONLY.

The arithmetic here belongs to OpenCV (an un-vendored dependency of the
reference), so the oracle calls the real OpenCV binary ``cv2``; the NumPy
restatements in :mod:`oracle.cvprims` are pinned against it in the tests."""
import cv2
import numpy as np

from . import cvprims

f32 = np.float32


def flow_remap(flow, image):
    H, W = image.shape
    m = np.empty((H, W, 2), f32)
    m[..., 0] = flow[..., 0] + np.arange(W, dtype=f32)[None, :]
    m[..., 1] = flow[..., 1] + np.arange(H, dtype=f32)[:, None]
    return cv2.remap(image, m, None, cv2.INTER_CUBIC)


def compare(prev, nxt):
    return cvprims.compare(prev, nxt, cv2.pyrDown, lambda s, sz: cv2.pyrUp(s, dstsize=(sz[1], sz[0])))


def image_gradient(img):
    gx = cv2.Sobel(img, cv2.CV_32F, 1, 0)
    gy = cv2.Sobel(img, cv2.CV_32F, 0, 1)
    return np.ascontiguousarray(np.stack([gx, gy], -1), f32)


def calculate_flow(prev, nxt, use_farneback=False):
    """Returns H x W x 4 float32 ``(u, v, variance, 0)``.  Quirk C1: the initial
    flow handed to VariationalRefinement is defined as zeros."""
    H, W = prev.shape
    if use_farneback:
        poly_sigma = (H + W) / 1000.0
        algo = cv2.FarnebackOpticalFlow_create(10, 0.8, False, (H + W) // 100, 7, 5 if poly_sigma < 1.5 else 7,
                                               poly_sigma, 0)
        flow = algo.calc(prev, nxt, None)
    else:
        algo = cv2.VariationalRefinement_create()
        flow = algo.calc(prev, nxt, np.zeros((H, W, 2), f32))
    var = compare(prev, flow_remap(flow, nxt))
    out = np.zeros((H, W, 4), f32)
    out[..., 0:2] = flow
    out[..., 2] = var
    return out
