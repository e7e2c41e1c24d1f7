"""Oracle for ``triangulatePixels`` (util.cpp:167-329).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np

from . import native
from .flow import image_gradient

f32 = np.float32


def triangulate_dense(flows, main_camera, cameras, depth, gradient=None, want_iters=False):
    """Pass 1 for every pixel: returns (dense H x W x 5 [X,Y,Z,W,pdf], valid H x W u8)."""
    H, W = depth.shape
    S = len(flows)
    flows = [np.ascontiguousarray(f, f32) for f in flows]
    ptrs = (C.c_void_p * S)(*[f.ctypes.data for f in flows])
    cams = np.ascontiguousarray(np.stack(cameras), f32).reshape(S, 16)
    depth = np.ascontiguousarray(depth, f32)
    grad = image_gradient(depth) if gradient is None else np.ascontiguousarray(gradient, f32)
    dense = np.empty((H, W, 5), f32)
    valid = np.empty((H, W), np.uint8)
    iters = np.empty((H, W), np.int32) if want_iters else None
    native.lib().orc_triangulate_dense(ptrs, S, np.ascontiguousarray(main_camera, f32).reshape(16), cams, depth, grad,
                                       W, H, dense, valid, iters.ctypes.data if want_iters else None)
    if want_iters:
        return dense, valid, iters
    return dense, valid


def triangulate_pixels(flows, main_camera, cameras, depth, gradient=None, return_evals=False):
    """``triangulatePixels``: M x 7 float32 rows (x, y, z, w, nx, ny, nz) in
    row-major pixel order.  With ``return_evals`` also the PCA eigenvalues and neighbour count
    of every row (M x 4: l0 >= l1 >= l2, K), a diagnostic the parity tests use to tell
    well-conditioned normals from ill-conditioned ones."""
    H, W = depth.shape
    S = len(flows)
    dense, valid = triangulate_dense(flows, main_camera, cameras, depth, gradient)
    cams = np.ascontiguousarray(np.stack(cameras), f32).reshape(S, 16)
    out = np.empty((int(valid.sum()), 7), f32)
    evals = np.zeros((max(len(out), 1), 4), f32) if return_evals else None
    m = native.lib().orc_normals_compact(dense, valid, W, H, np.ascontiguousarray(main_camera, f32).reshape(16), cams, S,
                                         out if len(out) else np.empty((1, 7), f32),
                                         evals.ctypes.data if return_evals else None)
    assert m == len(out)
    if return_evals:
        return out, evals[:len(out)]
    return out
