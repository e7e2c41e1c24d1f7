"""Restatements (NumPy, float32, no FMA) of the OpenCV primitives the reference's
hot path calls.  TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

OpenCV is an un-vendored, un-pinned dependency of the reference
(``Makefile:10`` links ``-lopencv_optflow -lopencv_video -lopencv_imgproc ...``;
API usage implies OpenCV 3.2-3.4 + contrib).  Its source is not in
``/root/reference``; these functions restate the published algorithms and are
pinned in ``tests/test_oracle_cv.py`` against the real OpenCV binary available
here (``cv2`` 4.13.0):

=====================  ==========================  ==============================
function               reference call site         agreement with cv2 4.13
=====================  ==========================  ==============================
variational_refinement flow.cpp:29,32              bit-exact (all sizes tested)
remap_cubic_8u         util.cpp:401                bit-exact
pyr_down / pyr_up      util.cpp:348-349,356        bit-exact (incl. OpenCV's SIMD/scalar column split)
sobel_gradient         util.cpp:473-474            bit-exact (incl. OpenCV's SIMD/scalar column split)
lu_inv4                util.cpp:174                bit-exact
inv2                   util.cpp:222                bit-exact
gemm_f32               util.cpp:86,89,99,209,219   bit-exact (small-matrix rule)
=====================  ==========================  ==============================

The CUDA kernels are written from the same formulas, in the same operation
order, compiled with ``-fmad=false`` and IEEE division / square root.
"""
import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------
# cv::optflow::createVariationalFlowRefinement()->calc(prev, next, flow=0)
# flow.cpp:29-32.  Defaults: fixedPoint 5, SOR 5, omega 1.6, alpha 20, delta 5,
# gamma 10, zeta 0.1, epsilon 1e-3.
# --------------------------------------------------------------------------
def _dx(a):
    p = np.pad(a, ((0, 0), (1, 1)), mode="edge")
    return p[:, 2:] - p[:, :-2]


def _dy(a):
    p = np.pad(a, ((1, 1), (0, 0)), mode="edge")
    return p[2:, :] - p[:-2, :]


def vr_derivatives(i0, i1):
    """Derivative fields of step 1-2 (initial flow is zero => warped I1 == I1).

    Central differences WITHOUT the 1/2 factor, replicated borders."""
    i0 = i0.astype(f32)
    i1 = i1.astype(f32)
    avg = f32(0.5) * i0 + f32(0.5) * i1
    iz = i1 - i0
    ix = _dx(avg)
    iy = _dy(avg)
    return dict(Ix=ix, Iy=iy, Iz=iz, Ixx=_dx(ix), Ixy=_dy(ix), Iyy=_dy(iy),
                Ixz=_dx(iz), Iyz=_dy(iz))


def variational_refinement(i0, i1, fixed_point=5, sor=5, omega=1.6, alpha=20.0,
                           delta=5.0, gamma=10.0, zeta=0.1, epsilon=0.001,
                           return_state=False):
    """Bit-exact restatement of ``VariationalRefinement::calc`` from a ZERO
    initial flow (quirk C1: the reference passes an uninitialised Mat; the
    oracle defines it as zeros).  Returns H x W x 2 float32."""
    d = vr_derivatives(i0, i1)
    Ix, Iy, Iz = d["Ix"], d["Iy"], d["Iz"]
    Ixx, Ixy, Iyy, Ixz, Iyz = d["Ixx"], d["Ixy"], d["Iyy"], d["Ixz"], d["Iyz"]
    H, W = Ix.shape
    du = np.zeros((H, W), f32)
    dv = np.zeros((H, W), f32)
    # NB: squared constants are float*float products, NOT float(double product)
    z2 = f32(zeta) * f32(zeta)
    e2 = f32(epsilon) * f32(epsilon)
    d2 = f32(delta / 2)
    g2 = f32(gamma / 2)
    a2 = f32(alpha / 2)
    om = f32(omega)
    yy, xx = np.mgrid[0:H, 0:W]
    red = ((xx + yy) % 2) == 0
    with np.errstate(all="ignore"):
        for _ in range(fixed_point):
            # -- data term -------------------------------------------------
            n = Ix * Ix + Iy * Iy + z2
            r = Iz + Ix * du + Iy * dv
            w = (d2 / np.sqrt(r * r / n + e2)) / n
            A11 = w * (Ix * Ix) + z2
            A12 = w * (Ix * Iy)
            A22 = w * (Iy * Iy) + z2
            b1 = -w * (Iz * Ix)
            b2 = -w * (Iz * Iy)
            n1 = Ixx * Ixx + Ixy * Ixy + z2
            n2 = Iyy * Iyy + Ixy * Ixy + z2
            rx = Ixz + Ixx * du + Ixy * dv
            ry = Iyz + Ixy * du + Iyy * dv
            w = g2 / np.sqrt(rx * rx / n1 + ry * ry / n2 + e2)
            A11 = A11 + w * (Ixx * Ixx / n1 + Ixy * Ixy / n2)
            A12 = A12 + w * (Ixx * Ixy / n1 + Ixy * Iyy / n2)
            A22 = A22 + w * (Ixy * Ixy / n1 + Iyy * Iyy / n2)
            b1 = b1 - w * (Ixx * Ixz / n1 + Ixy * Iyz / n2)
            b2 = b2 - w * (Ixy * Ixz / n1 + Iyy * Iyz / n2)
            # -- smoothness term (W == 0, so the b contributions vanish) ----
            ux = np.zeros((H, W), f32)
            vx = np.zeros((H, W), f32)
            uy = np.zeros((H, W), f32)
            vy = np.zeros((H, W), f32)
            ux[:, :-1] = du[:, 1:] - du[:, :-1]
            vx[:, :-1] = dv[:, 1:] - dv[:, :-1]
            uy[:-1, :] = du[1:, :] - du[:-1, :]
            vy[:-1, :] = dv[1:, :] - dv[:-1, :]
            ws = a2 / np.sqrt(ux * ux + vx * vx + uy * uy + vy * vy + e2)
            sR = ws.copy()
            sR[:, -1] = 0           # link p -> right(p)
            sD = ws.copy()
            sD[-1, :] = 0           # link p -> down(p)
            sL = np.zeros_like(ws)
            sL[:, 1:] = ws[:, :-1]  # link left(p) -> p
            sU = np.zeros_like(ws)
            sU[1:, :] = ws[:-1, :]  # link up(p) -> p
            # accumulation order differs by checkerboard colour (red-black
            # buffer traversal order inside OpenCV): verified bit-exact.
            A11 = np.where(red, (((A11 + sR) + sL) + sD) + sU, (((A11 + sL) + sR) + sU) + sD)
            A22 = np.where(red, (((A22 + sR) + sL) + sD) + sU, (((A22 + sL) + sR) + sU) + sD)
            # -- red-black SOR ----------------------------------------------
            for _s in range(sor):
                for color in (red, ~red):
                    P = np.pad(du, 1)
                    Q = np.pad(dv, 1)
                    su = sL * P[1:-1, :-2] + sR * P[1:-1, 2:] + sU * P[:-2, 1:-1] + sD * P[2:, 1:-1]
                    sv = sL * Q[1:-1, :-2] + sR * Q[1:-1, 2:] + sU * Q[:-2, 1:-1] + sD * Q[2:, 1:-1]
                    ndu = du + om * ((su + b1 - dv * A12) / A11 - du)
                    ndv = dv + om * ((sv + b2 - ndu * A12) / A22 - dv)
                    du = np.where(color, ndu, du)
                    dv = np.where(color, ndv, dv)
    out = np.stack([du, dv], -1)
    if return_state:
        return out, dict(A11=A11, A12=A12, A22=A22, b1=b1, b2=b2, ws=ws)
    return out


# --------------------------------------------------------------------------
# cv::remap(8UC1, INTER_CUBIC, BORDER_CONSTANT 0)   util.cpp:401
# --------------------------------------------------------------------------
def _cubic_coeffs(t):
    """Keys cubic, A = -0.75, evaluated in float32 exactly like OpenCV's
    ``interpolateCubic``."""
    A = f32(-0.75)
    t = f32(t)
    c0 = ((A * (t + f32(1)) - f32(5) * A) * (t + f32(1)) + f32(8) * A) * (t + f32(1)) - f32(4) * A
    c1 = ((A + f32(2)) * t - (A + f32(3))) * t * t + f32(1)
    u = f32(1) - t
    c2 = ((A + f32(2)) * u - (A + f32(3))) * u * u + f32(1)
    c3 = f32(1) - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], f32)


_CUBIC_TAB = None


def cubic_table_i16():
    """The 32x32 table of 4x4 fixed-point (Q15) bicubic weights used by
    ``cv::remap`` for 8-bit images (``initInterTab2D(INTER_CUBIC, fixpt=true)``).
    Index ``[ay*32+ax, ky, kx]``."""
    global _CUBIC_TAB
    if _CUBIC_TAB is not None:
        return _CUBIC_TAB
    one = [_cubic_coeffs(f32(i) / f32(32)) for i in range(32)]
    tab = np.zeros((1024, 4, 4), np.int16)
    for ay in range(32):
        for ax in range(32):
            t = np.outer(one[ay], one[ax]).astype(f32)           # cy (x) cx
            it = np.clip(np.rint(t * f32(32768)), -32768, 32767).astype(np.int32)
            isum = int(it.sum())
            if isum != 32768:
                diff = isum - 32768
                mk1 = mk2 = Mk1 = Mk2 = 2
                for k1 in range(2, 4):
                    for k2 in range(2, 4):
                        if it[k1, k2] < it[mk1, mk2]:
                            mk1, mk2 = k1, k2
                        elif it[k1, k2] > it[Mk1, Mk2]:
                            Mk1, Mk2 = k1, k2
                if diff < 0:
                    it[Mk1, Mk2] -= diff
                else:
                    it[mk1, mk2] -= diff
            tab[ay * 32 + ax] = it.astype(np.int16)
    _CUBIC_TAB = tab
    return tab


def remap_cubic_8u(img, mapx, mapy):
    """``cv::remap(img 8UC1, map 32FC2, INTER_CUBIC, BORDER_CONSTANT, 0)``.

    Coordinates are quantised to 1/32 px with round-half-to-even (cvRound of
    ``v*32``); samples outside the image contribute 0."""
    H, W = img.shape
    tab = cubic_table_i16().astype(np.int32)
    sx = np.rint(mapx.astype(f32) * f32(32)).astype(np.int64)
    sy = np.rint(mapy.astype(f32) * f32(32)).astype(np.int64)
    # saturate_cast<int> on cvRound result; then >>5 and saturate to short
    ix = np.clip(sx >> 5, -32768, 32767) - 1
    iy = np.clip(sy >> 5, -32768, 32767) - 1
    a = ((sy & 31) * 32 + (sx & 31)).astype(np.int64)
    wts = tab[a]                                            # H x W x 4 x 4
    acc = np.zeros(mapx.shape, np.int64)
    padded = np.zeros((H, W), np.int32) + img.astype(np.int32)
    for ky in range(4):
        yy = iy + ky
        vy = (yy >= 0) & (yy < H)
        yc = np.clip(yy, 0, H - 1)
        for kx in range(4):
            xx = ix + kx
            v = vy & (xx >= 0) & (xx < W)
            xc = np.clip(xx, 0, W - 1)
            acc += np.where(v, padded[yc, xc], 0) * wts[..., ky, kx]
    out = (acc + (1 << 14)) >> 15
    return np.clip(out, 0, 255).astype(np.uint8)


def flow_remap(flow, image):
    """``flowRemap`` util.cpp:390-403: map = flow[..., :2] + (x, y)."""
    H, W = image.shape
    xs = np.arange(W, dtype=f32)[None, :]
    ys = np.arange(H, dtype=f32)[:, None]
    return remap_cubic_8u(image, flow[..., 0].astype(f32) + xs, flow[..., 1].astype(f32) + ys)


# --------------------------------------------------------------------------
# cv::pyrDown / cv::pyrUp on float32  (util.cpp:348-349, 356)
# --------------------------------------------------------------------------
def _refl101(i, n):
    i = np.abs(np.asarray(i))
    if n == 1:
        return np.zeros_like(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def _refl101_scalar(i, n):
    if n == 1:
        return 0
    while i < 0 or i >= n:
        if i < 0:
            i = -i
        if i >= n:
            i = 2 * (n - 1) - i
    return i


def pyr_down(src):
    """``cv::pyrDown`` on float32: [1 4 6 4 1]/16 separable, BORDER_REFLECT_101, out ((W+1)/2, (H+1)/2).

    BIT-EXACT against cv2 4.13 (tests/test_oracle_cv.py), which requires reproducing which output columns
    OpenCV computes with its 4-lane SIMD body and which with its scalar code, because the two associate the
    five taps differently:
      horizontal  SIMD   c*6 + ((l + r)*4 + (ll + rr))      columns 1 .. 4*floor((width0-1)/4)
                  scalar ((c*6 + (l + r)*4) + ll) + rr      column 0, the remaining ones, and the right border
                                                            (width0 = min((W-3)/2 + 1, Wd))
      vertical    SIMD   ((r1 + r3 + r2)*4 + (r0 + r4 + (r2 + r2))) / 256     columns < 4*floor(Wd/4)
                  scalar (((r2*6 + (r1 + r3)*4) + r0) + r4) / 256             the rest"""
    src = src.astype(f32)
    H, W = src.shape
    Wd, Hd = (W + 1) // 2, (H + 1) // 2
    width0 = min(int((W - 3) / 2) + 1, Wd)            # C integer division truncates toward zero
    nvec_h = ((width0 - 1) // 4) * 4 if width0 - 1 >= 4 else 0
    xs = np.arange(Wd)
    idx = [np.array([_refl101_scalar(2 * x + k, W) for x in range(Wd)]) for k in (-2, -1, 0, 1, 2)]
    s0, s1, s2, s3, s4 = [src[:, i] for i in idx]
    simd = s2 * f32(6) + ((s1 + s3) * f32(4) + (s0 + s4))
    scal = ((s2 * f32(6) + (s1 + s3) * f32(4)) + s0) + s4
    row = np.where(((xs >= 1) & (xs < 1 + nvec_h))[None, :], simd, scal)
    idy = [np.array([_refl101_scalar(2 * y + k, H) for y in range(Hd)]) for k in (-2, -1, 0, 1, 2)]
    r0, r1, r2, r3, r4 = [row[i, :] for i in idy]
    simd = ((r1 + r3 + r2) * f32(4) + (r0 + r4 + (r2 + r2))) * f32(1.0 / 256)
    scal = (((r2 * f32(6) + (r1 + r3) * f32(4)) + r0) + r4) * f32(1.0 / 256)
    return np.where((xs < (Wd // 4) * 4)[None, :], simd, scal).astype(f32)


def pyr_up(src, dsize):
    """``cv::pyrUp`` on float32 to ``dsize=(H, W)`` (|H - 2h| <= 1, |W - 2w| <= 1).  BIT-EXACT against cv2 4.13.
      horizontal  even 2x : (s[x-1] + s[x]*6) + s[x+1]; x = 0: s[0]*6 + s[1]*2; x = w-1: s[w-2] + s[w-1]*7
                  odd 2x+1: (s[x] + s[x+1])*4;          x = w-1: s[w-1]*8;      w = 1: both s*8
      vertical    even 2y : ((r[y-1] + r[y]*6) + r[y+1]) / 64   (row -1 -> 1, row h -> h-1)
                  odd 2y+1: ((r[y] + r[y+1])*4) / 64
      odd sizes   W = 2w+1: last column repeats column 2w-1; H = 2h+1: last row repeats row 2h-2;
                  W = 2w-1 / H = 2h-1: the surplus odd column / row is dropped."""
    src = src.astype(f32)
    h, w = src.shape
    H, W = dsize
    row = np.zeros((h, max(W, 2 * w)), f32)
    if w == 1:
        row[:, 0] = src[:, 0] * f32(8)
        row[:, 1] = src[:, 0] * f32(8)
    else:
        a, b, c = src[:, :-2], src[:, 1:-1], src[:, 2:]
        row[:, 2:2 * w - 2:2] = a + b * f32(6) + c
        row[:, 3:2 * w - 1:2] = (b + c) * f32(4)
        row[:, 0] = src[:, 0] * f32(6) + src[:, 1] * f32(2)
        row[:, 1] = (src[:, 0] + src[:, 1]) * f32(4)
        row[:, 2 * w - 2] = src[:, w - 2] + src[:, w - 1] * f32(7)
        row[:, 2 * w - 1] = src[:, w - 1] * f32(8)
    if W > 2 * w:
        row[:, W - 1] = row[:, 2 * w - 1]
    yu = np.array([_refl101_scalar(2 * (y - 1), 2 * h) // 2 for y in range(h)])
    yd = np.array([_refl101_scalar(2 * (y + 1), 2 * h) // 2 for y in range(h)])
    r0, r1, r2 = row[yu], row, row[yd]
    out = np.zeros((max(H, 2 * h), row.shape[1]), f32)
    out[1:2 * h:2] = ((r1 + r2) * f32(4)) * f32(1.0 / 64)
    out[0:2 * h:2] = ((r0 + r1 * f32(6)) + r2) * f32(1.0 / 64)
    if H > 2 * h:
        out[2 * h] = out[2 * h - 2]
    return out[:H, :W].astype(f32)


def compare(prev, nxt, pyr_down_fn=pyr_down, pyr_up_fn=pyr_up):
    """``compare`` util.cpp:332-361: L1 difference summed over a Gaussian
    pyramid, up-sampled back to full resolution."""
    a = prev.astype(f32)
    b = nxt.astype(f32)
    size = min(prev.shape)
    pyr = []
    while True:
        pyr.append(np.abs(a - b))
        if size <= 2:
            break
        a = pyr_down_fn(a)
        b = pyr_down_fn(b)
        size //= 2
    for i in range(len(pyr) - 2, -1, -1):
        pyr[i] = pyr[i] + pyr_up_fn(pyr[i + 1], pyr[i].shape)
    return pyr[0]


def compare_levels(h, w):
    """Sizes of the pyramid levels used by ``compare`` for an h x w image."""
    size = min(h, w)
    out = [(h, w)]
    while size > 2:
        h, w = (h + 1) // 2, (w + 1) // 2
        out.append((h, w))
        size //= 2
    return out


# --------------------------------------------------------------------------
# imageGradient  util.cpp:465-479  (cv::Sobel 3x3, CV_32F, BORDER_REFLECT_101)
# --------------------------------------------------------------------------
def sobel_gradient(img):
    """Returns H x W x 2 float32 (gx, gy), BIT-EXACT against cv2.Sobel (ksize 3, CV_32F, BORDER_REFLECT_101).
    Row filter first, then the column filter.  The [1 2 1] smoothing is associated differently by OpenCV's 8-lane
    SIMD body and by its scalar tail:
      gx (column filter over the row differences d):  (d0 + d2) + 2*d1   columns < 8*floor(W/8)
                                                      (d0 + 2*d1) + d2   the tail columns
      gy (row filter l, c, r, then column difference): (l + r) + 2*c      columns < 8*floor(W/8) and a final
                                                                          unpaired tail column
                                                      (l + 2*c) + r      the tail columns taken in pairs"""
    img = img.astype(f32)
    H, W = img.shape
    p = np.pad(img, 1, mode="reflect") if min(H, W) > 1 else np.pad(img, 1, mode="edge")
    xs = np.arange(W)
    t0 = (W // 8) * 8
    tail = (xs >= t0)[None, :]
    dxr = p[:, 2:] - p[:, :-2]
    gx = np.where(tail, dxr[:-2] + dxr[1:-1] * f32(2) + dxr[2:], (dxr[:-2] + dxr[2:]) + dxr[1:-1] * f32(2))
    npair = ((W - t0) // 2) * 2
    tailp = ((xs >= t0) & (xs < t0 + npair))[None, :]
    l, c, r = p[:, :-2], p[:, 1:-1], p[:, 2:]
    sx = np.where(tailp, l + c * f32(2) + r, (l + r) + c * f32(2))
    gy = sx[2:] - sx[:-2]
    return np.stack([gx, gy], -1).astype(f32)


# --------------------------------------------------------------------------
# small dense linear algebra exactly as cv::Mat does it
# --------------------------------------------------------------------------
def lu_inv4(m):
    """``Mat::inv()`` (DECOMP_LU) of a 4x4 float32 matrix: OpenCV's built-in
    ``hal::LU32f`` (partial pivoting, float32 throughout). util.cpp:174."""
    n = 4
    A = np.array(m, dtype=f32).copy()
    b = np.eye(n, dtype=f32)
    for i in range(n):
        k = i
        for j in range(i + 1, n):
            if abs(A[j, i]) > abs(A[k, i]):
                k = j
        if abs(A[k, i]) < np.finfo(f32).eps:
            return np.zeros((4, 4), f32)
        if k != i:
            A[[i, k], i:] = A[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = f32(-1) / A[i, i]
        for j in range(i + 1, n):
            alpha = f32(A[j, i] * d)
            for kk in range(i + 1, n):
                A[j, kk] = f32(A[j, kk] + f32(alpha * A[i, kk]))
            for kk in range(n):
                b[j, kk] = f32(b[j, kk] + f32(alpha * b[i, kk]))
    for i in range(n - 1, -1, -1):
        for j in range(n):
            s = b[i, j]
            for kk in range(i + 1, n):
                s = f32(s - f32(A[i, kk] * b[kk, j]))
            b[i, j] = f32(s / A[i, i])
    return b


def inv2(m):
    """``Mat::inv()`` of a 2x2 float32 matrix (determinant in double, then the
    reciprocal cast to float and multiplied in float). util.cpp:222."""
    m = np.asarray(m, f32)
    det = float(m[0, 0]) * float(m[1, 1]) - float(m[0, 1]) * float(m[1, 0])
    if det == 0.0:
        return np.zeros((2, 2), f32)
    d = f32(1.0 / det)
    return np.array([[m[1, 1] * d, m[0, 1] * (-d)], [m[1, 0] * (-d), m[0, 0] * d]], f32)


def gemm_f32(a, b):
    """``cv::gemm`` on small float32 matrices as used by ``Mat * Mat``:
    if 2 <= K <= 4 and K equals the output width or height, products are
    accumulated sequentially in float32; otherwise in double."""
    a = np.asarray(a, f32)
    b = np.asarray(b, f32)
    M, K = a.shape
    N = b.shape[1]
    if 2 <= K <= 4 and (K == N or K == M):
        c = np.zeros((M, N), f32)
        for i in range(M):
            for j in range(N):
                s = f32(a[i, 0] * b[0, j])
                for k in range(1, K):
                    s = f32(s + f32(a[i, k] * b[k, j]))
                c[i, j] = s
        return c
    return (a.astype(np.float64) @ b.astype(np.float64)).astype(f32)


# ---------------------------------------------------------------------------
# cv::resize(..., INTER_AREA) on CV_8UC<cn> for a NON-integer shrink factor (configuration.cpp:232-233 with a fractional
# -s): OpenCV's computeResizeAreaTab + ResizeArea_Invoker<uchar, float>, restated; bit-exact vs cv2 (tests/test_oracle_cv.py).
# (Integer factors take resizeAreaFast_ instead; the CUDA ingest follows the same split.)
# ---------------------------------------------------------------------------
def _area_tab(ssize, dsize):
    import math
    scale = float(ssize) / dsize
    tab = []
    for d in range(dsize):
        fs1 = d * scale
        fs2 = fs1 + scale
        cell = min(scale, ssize - fs1)
        s1, s2 = math.ceil(fs1), math.floor(fs2)
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        if s1 - fs1 > 1e-3:
            tab.append((d, s1 - 1, f32((s1 - fs1) / cell)))
        for sx in range(s1, s2):
            tab.append((d, sx, f32(1.0 / cell)))
        if fs2 - s2 > 1e-3:
            tab.append((d, s2, f32(min(min(fs2 - s2, 1.0), cell) / cell)))
    return tab


def resize_area_8u(src, W, H):
    """``cv2.resize(src, (W, H), interpolation=cv2.INTER_AREA)`` for uint8 ``src`` (h x w or h x w x cn) being SHRUNK by a
    factor that is not an integer in both directions."""
    sh, sw = src.shape[:2]
    cn = src.shape[2] if src.ndim == 3 else 1
    s = src.reshape(sh, sw, cn)
    xtab, ytab = _area_tab(sw, W), _area_tab(sh, H)
    xd = np.array([t[0] for t in xtab])
    xs = np.array([t[1] for t in xtab])
    xa = np.array([t[2] for t in xtab], f32)
    dst = np.zeros((H, W, cn), np.uint8)
    acc = np.zeros((W, cn), f32)
    prev = ytab[0][0]
    for dy, sy, beta in ytab:
        buf = np.zeros((W, cn), f32)
        row = s[sy].astype(f32)
        for k in range(len(xtab)):                       # float accumulation in table order
            buf[xd[k]] = buf[xd[k]] + row[xs[k]] * xa[k]
        if dy != prev:
            dst[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
            acc = beta * buf
            prev = dy
        else:
            acc = acc + beta * buf
    dst[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return dst if src.ndim == 3 else dst[..., 0]
