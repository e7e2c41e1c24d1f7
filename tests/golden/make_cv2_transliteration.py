#!/usr/bin/env python
"""Golden generator: a cv2-LEVEL TRANSLITERATION of the reference's triangulatePixels / triangulatePixel
(util.cpp:62-329, 33-53, 438-461) -- every cv::Mat expression of the reference is evaluated, in the reference's
statement order, by the REAL OpenCV binary (cv2.gemm / invert / divide / subtract / multiply / determinant / norm /
PCACompute2 / decomposeProjectionMatrix / Sobel), not by our restatement.  The only things written out by hand are
what has no cv2 entry point: Mat::dot on 2- and 3-vectors (OpenCV accumulates it in double), the scalar float / double
statements, and goodSample / sampleImage<T> (plain C++ in the reference).

Purpose ("reference-side pinning of the oracle"): oracle/recon_oracle.c and the CUDA path share the host code
that decides WHICH products accumulate in double, how `Mat /= s` rounds, how cv::PCA forms its mean and covariance ...;
this script pins those decisions to the OpenCV binary.  tests/test_oracle_cv2_transliteration.py requires
recon_oracle.c to reproduce the committed output bit for bit.

    python tests/golden/make_cv2_transliteration.py        # writes tests/golden/cv2_translit_s2_96x72.npz (~1 minute)
                                                           # and tests/golden/cv2_translit_s34_40x30.npz (S = 4 and S = 3)

The second file pins a rule that only shows with FOUR side cameras: `projectionW * k` (util.cpp:105) is then a 4x4 by 4x1
cv::gemm, which OpenCV evaluates with its hand-written small-matrix kernels in FLOAT, whereas the S x 4 product of any
other S goes through the generic kernel with DOUBLE accumulators -- the oracle's and the CUDA kernel's `S == 4` branch.

Quirk decisions are the documented ones: C5 continuous-memory reads past a row end, C8 `dot` starts at 0,
C9 the neighbourhood is cleared for every pixel.
"""
import math
import os
import struct
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32
backgroundDepth = f32(1.0)


def F(x):
    return f32(x)


def gemm(a, b, flags=0):
    """MatExpr a * b on CV_32F matrices"""
    return cv2.gemm(np.ascontiguousarray(a, f32), np.ascontiguousarray(b, f32), 1.0, None, 0.0, flags=flags)


def scale(m, s):
    """MatExpr m * s (s a double) is Mat::convertTo(dst, type, s), which for CV_32F multiplies by (float)s in float
    (cvtScale_<float, float, float>).  convertTo has no Python binding and cv2.multiply(m, Scalar) works in double, so the
    float product is taken with the element-wise cv2.multiply against a matrix filled with (float)s."""
    m = np.ascontiguousarray(m, f32)
    return cv2.multiply(m, np.full_like(m, f32(s))).reshape(m.shape)


def ddiv(a, b):
    """IEEE double division (inf / NaN instead of Python's ZeroDivisionError)"""
    with np.errstate(all="ignore"):
        return float(np.float64(a) / np.float64(b))


def div_scalar(m, s):
    """m / s and m /= s with s a float: MatExpr scaling by the DOUBLE 1. / s"""
    with np.errstate(all="ignore"):
        return scale(m, ddiv(1.0, s))


def dot(a, b):
    """Mat::dot (dotProd_32f): double accumulation of the float products"""
    r = 0.0
    for x, y in zip(np.asarray(a, f32).ravel(), np.asarray(b, f32).ravel()):
        r += float(x) * float(y)
    return r


def fmod1(x):
    return f32(math.fmod(float(x), 1.0))


def good_sample(image, x, y):                                  # util.cpp:44-53
    ix, iy = int(x), int(y)
    H, W = image.shape[:2]
    if ix <= 0 or ix >= W - 1 or iy <= 0 or iy >= H - 1:
        return False
    return (image[iy, ix] != backgroundDepth and image[iy, ix + 1] != backgroundDepth and
            image[iy + 1, ix] != backgroundDepth and image[iy + 1, ix + 1] != backgroundDepth)


def at_flat(image, y, x):
    """image.at<T>(y, x) with float arguments truncated to int, on a CONTINUOUS Mat (C5: x == cols reads the next row)"""
    H, W = image.shape[:2]
    idx = int(y) * W + int(x)
    idx = min(max(idx, 0), H * W - 1)
    return image.reshape(H * W, -1)[idx]


def sample_float(image, x, y):                                 # sampleImage<float>, util.cpp:438-461
    lw, tw = fmod1(x), fmod1(y)
    rw, bw = f32(1) - lw, f32(1) - tw
    a, b = at_flat(image, y, x)[0], at_flat(image, y, F(x) + f32(1))[0]
    c, d = at_flat(image, F(y) + f32(1), x)[0], at_flat(image, F(y) + f32(1), F(x) + f32(1))[0]
    if rw == 0:
        return a if bw == 0 else f32(f32(a * tw) + f32(c * bw))
    if bw == 0:
        return f32(f32(a * lw) + f32(b * rw))
    return f32(f32(f32(f32(a * lw) + f32(b * rw)) * tw) + f32(f32(f32(c * lw) + f32(d * rw)) * bw))


def cv_round(v):
    """saturate_cast<int>(float) = cvRound: round half to even; out of range -> INT_MIN (x86 cvtss2si)"""
    v = float(v)
    if not (-2147483648.0 < v < 2147483648.0):
        return -2147483648
    return int(np.rint(v))


def wrap32(v):
    return (v + 2 ** 31) % 2 ** 32 - 2 ** 31


def sample_point(grad, x, y):
    """sampleImage<cv::Point> on a CV_32FC2 gradient (util.cpp:215-217): the float bits are read as ints (C4), Point * float
    saturate_casts each product, Point + Point wraps"""
    def bits(p):
        return [struct.unpack("<i", struct.pack("<f", float(v)))[0] for v in p]

    def pmul(p, w):
        return [cv_round(f32(f32(c) * w)) for c in p]

    def padd(p, q):
        return [wrap32(c + d) for c, d in zip(p, q)]

    lw, tw = fmod1(x), fmod1(y)
    rw, bw = f32(1) - lw, f32(1) - tw
    a, b = bits(at_flat(grad, y, x)), bits(at_flat(grad, y, F(x) + f32(1)))
    c, d = bits(at_flat(grad, F(y) + f32(1), x)), bits(at_flat(grad, F(y) + f32(1), F(x) + f32(1)))
    if rw == 0:
        r = a if bw == 0 else padd(pmul(a, tw), pmul(c, bw))
    elif bw == 0:
        r = padd(pmul(a, lw), pmul(b, rw))
    else:
        r = padd(pmul(padd(pmul(a, lw), pmul(b, rw)), tw), pmul(padd(pmul(c, lw), pmul(d, rw)), bw))
    return [struct.unpack("<f", struct.pack("<i", v))[0] for v in r]      # written back as bits into the float slots of D


def image_gradient(depth):                                    # util.cpp:465-479
    gx = cv2.Sobel(depth, cv2.CV_32F, 1, 0)
    gy = cv2.Sobel(depth, cv2.CV_32F, 0, 1)
    return np.ascontiguousarray(np.stack([gx, gy], -1), f32)


def extract_camera_center(camera):                             # util.cpp:33-41
    projection = np.ascontiguousarray(np.concatenate([camera[0:2], camera[3:4]], 0), f32)
    return cv2.decomposeProjectionMatrix(projection)[2].astype(f32)      # T, 4x1 (the function returns it in the input's depth)


def triangulate_pixel(x, y, measuredPoints, icovars, mainCameraInv, cameras, depth):     # util.cpp:62-164
    S = len(cameras)
    k = np.array([[x], [y], [depth], [1]], f32)
    p = np.zeros((2, S), f32)
    delta_p = np.zeros((2, S), f32)
    projectionDerivatives = np.zeros((2, S), f32)
    projectionW = np.zeros((S, 4), f32)
    for i, camera in enumerate(cameras):
        projectionDerivatives[:, i:i + 1] = gemm(camera[0:2], mainCameraInv[:, 2:3])
        projectionW[i] = camera[3]
    projectionW = gemm(projectionW, mainCameraInv)
    it = 0
    while True:
        for i, camera in enumerate(cameras):
            estimatedPoint = gemm(gemm(camera, mainCameraInv), k)
            estimatedPoint = div_scalar(estimatedPoint, estimatedPoint[3, 0])
            p[:, i] = estimatedPoint[0:2, 0]
        pointsW = np.ascontiguousarray(gemm(projectionW, k).T)
        delta_p[0:1] = cv2.divide(np.ascontiguousarray(projectionDerivatives[0:1]), pointsW)
        delta_p[1:2] = cv2.divide(np.ascontiguousarray(projectionDerivatives[1:2]), pointsW)
        firstDz, secondDz = 0.0, 0.0
        difference = cv2.subtract(p, measuredPoints)
        for i in range(S):
            transformed = gemm(icovars[i], delta_p[:, i:i + 1])
            firstDz += dot(difference[:, i], transformed)
            secondDz += dot(delta_p[:, i], transformed)
        with np.errstate(all="ignore"):
            delta_z = float(-np.float64(firstDz) / np.float64(secondDz))
        eps = 1e-7
        if it >= 50 or (delta_z < eps and delta_z > -eps):
            exponent, product_ivar = 0.0, 1.0
            for i in range(S):
                transformed = gemm(icovars[i], difference[:, i:i + 1])
                exponent -= dot(difference[:, i], transformed)
                product_ivar *= cv2.determinant(np.ascontiguousarray(icovars[i]))
            with np.errstate(all="ignore"):
                pdf = f32(np.float64(0.159) * np.float64(product_ivar) * np.exp(np.float64(0.5 * exponent)))
            break
        k[2, 0] = f32(float(k[2, 0]) + delta_z)
        it += 1
    return gemm(mainCameraInv, k), pdf


def triangulate_pixels(flows, mainCamera, cameras, depth):                                  # util.cpp:167-329
    H, W = depth.shape
    S = len(cameras)
    rows = []
    mainCameraInv = cv2.invert(np.ascontiguousarray(mainCamera, f32))[1]
    gradient = image_gradient(depth)
    pixelIndices = -np.ones((H, W), np.int32)
    for row in range(H):
        for col in range(W):
            if depth[row, col] == backgroundDepth:
                continue
            okay = True
            centerX, centerY = f32(W / 2.0), f32(H / 2.0)
            scaleX, scaleY = f32(2.0 / W), f32(2.0 / H)
            x = f32(f32(f32(col) - centerX) * scaleX)
            y = f32(f32(centerY - f32(row)) * scaleY)
            measuredPoints = np.zeros((2, S), f32)
            icovars = []
            for i, (camera, flow) in enumerate(zip(cameras, flows)):
                flx, fly, variance = flow[row, col, 0], flow[row, col, 1], flow[row, col, 2]
                sx, sy = f32(f32(col) + flx), f32(f32(row) + fly)
                good = good_sample(depth, sx, sy)
                z = sample_float(depth[..., None], sx, sy) if good else depth[row, col]
                vec = np.array([[f32(x + f32(flx * scaleX))], [f32(y + f32(fly * scaleY))], [z], [1]], f32)
                measuredPoint = gemm(gemm(camera, mainCameraInv), vec)
                D = np.eye(3, 2, dtype=f32)
                D[2] = sample_point(gradient, sx, sy) if good else sample_point(gradient, f32(col), f32(row))
                with np.errstate(all="ignore"):
                    A = gemm(gemm(camera[0:2, 0:3], mainCameraInv[0:3, 0:3]), D)
                    A = div_scalar(A, measuredPoint[3, 0])
                    icovarMatrix = div_scalar(cv2.invert(gemm(A, A, cv2.GEMM_2_T))[1], variance)
                    icovars.append(np.ascontiguousarray(icovarMatrix.reshape(2, 2)))
                    measuredPoint = div_scalar(measuredPoint, measuredPoint[3, 0])
                if measuredPoint[2, 0] < -1:
                    okay = False
                    break
                measuredPoints[:, i] = measuredPoint[0:2, 0]
            if okay:
                point, pdf = triangulate_pixel(x, y, measuredPoints, icovars, mainCameraInv, cameras, depth[row, col])
                pixelIndices[row, col] = len(rows)
                rows.append(list(point[:, 0]) + [pdf, 0, 0])
    points = np.array(rows, f32).reshape(-1, 7)
    # ---- normals ----
    radius = 10
    centers = [extract_camera_center(mainCamera)] + [extract_camera_center(c) for c in cameras]
    centers = [div_scalar(np.ascontiguousarray(c[0:3].T), c[3, 0]) for c in centers]
    out = points.copy()
    for row in range(H):
        for col in range(W):
            pid = pixelIndices[row, col]
            if pid < 0:
                continue
            pdf = points[pid, 4]
            if S > 1:
                with np.errstate(all="ignore"):
                    pdf = f32(np.power(np.float64(pdf), np.float64(1.0 / S)))        # C pow(): NaN for a negative base, no exception
            nb = []
            for ny in range(row - radius, row + radius + 1):
                if ny < 0 or ny >= H:
                    continue
                for nx in range(col - radius, col + radius + 1):
                    if nx < 0 or nx >= W or pixelIndices[ny, nx] < 0:
                        continue
                    q = pixelIndices[ny, nx]
                    nb.append(div_scalar(np.ascontiguousarray(points[q:q + 1, 0:3]), points[q, 3])[0])
            if len(nb) >= 3:
                _, evecs, _ = cv2.PCACompute2(np.ascontiguousarray(np.array(nb, f32)), mean=None)
                normal = np.ascontiguousarray(evecs[2:3], f32)
                d = f32(0)                                                        # C8
                me = div_scalar(np.ascontiguousarray(points[pid:pid + 1, 0:3]), points[pid, 3])
                for c in centers:
                    with np.errstate(all="ignore"):
                        d = f32(float(d) + ddiv(1.0, dot(normal, cv2.subtract(c, me))))
                if d < 0:
                    normal = -normal
            else:
                normal = np.zeros((1, 3), f32)
                for c in centers:
                    vec = cv2.subtract(c, np.ascontiguousarray(points[pid:pid + 1, 0:3]))
                    with np.errstate(all="ignore"):
                        # normal += vec / vec.dot(vec): MatOp::augAssignAdd evaluates the expression into a temporary
                        # (convertTo with alpha = 1 / dot), then cv::add -- two rounded operations (cv2.scaleAdd of the 4.x
                        # binary fuses them on FMA hardware and differs from this in the last bit of one K = 2 row)
                        normal = cv2.add(normal, div_scalar(vec, dot(vec, vec)))
            with np.errstate(all="ignore"):
                out[pid, 4:7] = scale(normal, float(pdf) * ddiv(1.0, cv2.norm(normal)))[0]
    return out


def main():
    g = np.load(os.path.join(HERE, "scene_s2_96x72.npz"))
    fa, sides = int(g["fa"]), [int(s) for s in g["sides"]]
    cams = g["cameras"].astype(f32)
    flows = [np.ascontiguousarray(f, f32) for f in g["flows"]]
    depth = np.ascontiguousarray(g["depth"], f32)
    tri = triangulate_pixels(flows, cams[fa], [cams[s] for s in sides], depth)
    # a second case with S = 1, isolated pixels (K < 3 branch) and zero variance (NaN rows)
    d1 = np.ones_like(depth)
    keep = np.zeros(depth.shape, bool)
    keep[10:14, 20:30] = True
    keep[40, 50] = keep[40, 70] = keep[41, 71] = True
    keep[60:66, 5:12] = True
    keep &= depth != 1.0
    d1[keep] = depth[keep]
    fl1 = flows[0].copy()
    fl1[12, 22, 2] = 0.0
    tri1 = triangulate_pixels([fl1], cams[fa], [cams[sides[0]]], d1)
    out = os.path.join(HERE, "cv2_translit_s2_96x72.npz")
    np.savez_compressed(out, tri=tri, tri1=tri1, depth1=d1, flow1=fl1, cv2_version=cv2.__version__)
    print("wrote", out, tri.shape, tri1.shape, "cv2", cv2.__version__)
    main_s34()


def main_s34():
    """S = 4 (BASELINE config 5's schedule: sides fa-2, fa-1, fa+1, fa+2) and S = 3 on a 40 x 30 frame.  The inputs (flows,
    depth after the cumulative mixBackground masking, cameras) come from the oracle's front half -- cv2's
    VariationalRefinement etc. -- and are stored next to the transliteration's rows."""
    root = os.path.dirname(os.path.dirname(HERE))
    sys.path.insert(0, root)
    from mesh_reconstruction_b200 import synth
    from oracle.pipeline import process_main_frame
    from oracle.render import RenderOracle
    W, H = 40, 30
    sc = synth.make_scene(W, H, 5, seed=11, step=0.12, mesh_err=0.03, mesh_res=6)
    frames = sc.frames()
    r = RenderOracle(W, H)
    r.loadMesh(sc.vertices, sc.faces)
    fa = 2
    res = {}
    for name, sides in (("s4", [0, 1, 3, 4]), ("s3", [1, 3, 4])):
        _, inter = process_main_frame(r, frames, sc.cameras, fa, sides, keep=True)
        flows = [np.ascontiguousarray(f, f32) for f in inter["flows"]]
        depth = np.ascontiguousarray(inter["depth"], f32)
        tri = triangulate_pixels(flows, sc.cameras[fa].astype(f32), [sc.cameras[s].astype(f32) for s in sides], depth)
        res.update({f"flows_{name}": np.stack(flows), f"depth_{name}": depth, f"sides_{name}": np.asarray(sides), f"tri_{name}": tri})
        print(name, tri.shape, "nan rows", int(np.isnan(tri).any(1).sum()))
    out = os.path.join(HERE, "cv2_translit_s34_40x30.npz")
    np.savez_compressed(out, cameras=sc.cameras.astype(f32), fa=fa, cv2_version=cv2.__version__, **res)
    print("wrote", out)


if __name__ == "__main__":
    sys.exit(main())
