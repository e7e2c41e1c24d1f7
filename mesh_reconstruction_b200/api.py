"""ctypes binding of ``libmeshrecon_b200.so`` mirroring the reference's C++ interface
(``recon.hpp``).  Every function takes NumPy arrays (host buffers) or torch CUDA tensors
(device buffers, zero-copy) and forwards raw pointers to the C ABI.

Reference interface -> here:
  spawnRender(hint)                     recon.hpp:100     -> spawnRender(width, height, device)
  Render::loadMesh/depth/projected      recon.hpp:93-99   -> Render.loadMesh/depth/projected
  calculateFlow(prev, next, farneback)  recon.hpp:40      -> calculateFlow
  mixBackground(image, bg, Mat &depth)  recon.hpp:49      -> mixBackground (depth mutated in place)
  triangulatePixels(flows, P, cams, d)  recon.hpp:44      -> triangulatePixels
  compare / flowRemap / imageGradient   recon.hpp:45,50,55
  extractCameraCenter                   recon.hpp:43
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MR_MAX_SIDE = 16
_ERRNAMES = {-1: "MR_EINVAL", -2: "MR_ENODEVICE", -3: "MR_ECUDA", -4: "MR_ENOMESH", -5: "MR_ENOMEM"}


class MeshReconError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_ERRNAMES.get(code, code)}: {msg}")
        self.code = code


def library_path():
    return os.path.join(_HERE, "libmeshrecon_b200.so")


def load_library():
    """Loads the CUDA library; raises (never falls back) if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise MeshReconError(-2, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(there is no CPU fallback for the hot path)")
    L = C.CDLL(path)
    vp, ip, fpp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p)
    L.mr_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int]
    L.mr_destroy.argtypes = [vp]
    L.mr_destroy.restype = None
    L.mr_last_error.argtypes = [vp]
    L.mr_last_error.restype = C.c_char_p
    L.mr_version.restype = C.c_int
    L.mr_stream.argtypes = [vp]
    L.mr_stream.restype = vp
    L.mr_synchronize.argtypes = [vp]
    L.mr_load_mesh.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.mr_depth.argtypes = [vp, vp, vp]
    L.mr_projected.argtypes = [vp, vp, vp, vp, vp]
    L.mr_depth_samples.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, vp]
    L.mr_mix_background.argtypes = [vp, vp, vp, vp, vp]
    L.mr_calculate_flow.argtypes = [vp, vp, vp, C.c_int, vp]
    L.mr_flow_remap.argtypes = [vp, vp, C.c_int, vp, vp]
    L.mr_compare.argtypes = [vp, vp, vp, vp]
    L.mr_image_gradient.argtypes = [vp, vp, vp]
    L.mr_triangulate_pixels.argtypes = [vp, fpp, C.c_int, vp, vp, vp, vp, ip]
    L.mr_extract_camera_center.argtypes = [vp, vp]
    L.mr_process_main_frame.argtypes = [vp, vp, vp, C.c_int, fpp, vp, vp, ip]
    L.mr_process_main_frame_async.argtypes = [vp, vp, vp, C.c_int, fpp, vp, vp, ip]
    L.mr_wait_copies.argtypes = [vp]
    L.mr_wait_copies_until.argtypes = [vp, C.c_int]
    L.mr_set_use_graphs.argtypes = [vp, C.c_int]
    L.mr_allgather_points.argtypes = [vp, vp, vp, C.c_int, vp, C.c_size_t, ip, C.POINTER(C.c_longlong)]
    L.mr_xchg_alloc.argtypes = [vp, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
    L.mr_xchg_free.argtypes = [vp, vp]
    L.mr_xchg_open.argtypes = [vp, C.c_char_p, C.POINTER(C.c_void_p)]
    L.mr_xchg_close.argtypes = [vp, vp]
    L.mr_xchg_push.argtypes = [vp, vp, vp, C.c_size_t]
    L.mr_xchg_push_mcast.argtypes = [vp, vp, vp, C.c_size_t]
    L.mr_xchg_stream.argtypes = [vp]
    L.mr_xchg_stream.restype = vp
    L.mr_graph_launch_count.argtypes = [vp]
    L.mr_graph_launch_count.restype = C.c_uint64
    L.mr_submit_main_frame.argtypes = [vp, vp, vp, C.c_int, fpp, vp, vp, vp]
    L.mr_points_device.argtypes = [vp, ip]
    L.mr_points_device.restype = vp
    L.mr_last_depth_device.argtypes = [vp]
    L.mr_last_depth_device.restype = vp
    L.mr_last_flow_device.argtypes = [vp, C.c_int]
    L.mr_last_flow_device.restype = vp
    L.mr_last_mixed_device.argtypes = [vp, C.c_int]
    L.mr_last_mixed_device.restype = vp
    L.mr_launch_count.argtypes = [vp]
    L.mr_launch_count.restype = C.c_uint64
    L.mr_set_vr_impl.argtypes = [C.c_int]
    L.mr_set_use_farneback.argtypes = [vp, C.c_int]
    L.mr_profile_enable.argtypes = [vp, C.c_int]
    L.mr_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]
    L.mr_normals_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.mr_ingest_frame.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.mr_ingest_frame_exposure.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(C.c_float), vp]
    L.mr_set_gray_shift.argtypes = [vp, C.c_int]
    L.mr_filter_points.argtypes = [vp, vp, vp, C.c_size_t, C.c_float, vp, vp, vp, C.POINTER(C.c_size_t)]
    L.mr_filter_rows.argtypes = [vp, vp, C.c_size_t, C.c_float, vp, vp, C.POINTER(C.c_size_t)]
    L.mr_filter_info.argtypes = [vp, C.POINTER(C.c_longlong), vp, vp]
    L.mr_debug_seqsum.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_double)]
    L.mr_stage_name.argtypes = [C.c_int]
    L.mr_stage_name.restype = C.c_char_p
    _LIB = L
    return L


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x, dtype=None, shape=None, name="argument"):
    """Raw pointer of a C-contiguous NumPy array or torch tensor (plus a keep-alive ref)."""
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError(f"{name}: tensor must be contiguous")
        return x.data_ptr(), x
    a = np.ascontiguousarray(x, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
    return a.ctypes.data, a


def _mat16(m):
    a = np.ascontiguousarray(np.asarray(m, np.float32).reshape(16))
    return a.ctypes.data, a


class Context:
    """Owns one ``mr_context`` (one per GPU and render size)."""

    def __init__(self, width, height, device=0):
        self.lib = load_library()
        self.W, self.H, self.device = int(width), int(height), int(device)
        h = C.c_void_p()
        rc = self.lib.mr_create(C.byref(h), self.device, self.W, self.H)
        if rc:
            raise MeshReconError(rc, self.lib.mr_last_error(None).decode())
        self.h = h

    def check(self, rc):
        if rc:
            raise MeshReconError(rc, self.lib.mr_last_error(self.h).decode())

    def synchronize(self):
        self.check(self.lib.mr_synchronize(self.h))

    def wait_copies(self):
        self.check(self.lib.mr_wait_copies(self.h))

    def wait_copies_until(self, max_in_flight):
        """Block until at most ``max_in_flight`` of the queued row copies are still outstanding."""
        self.check(self.lib.mr_wait_copies_until(self.h, int(max_in_flight)))

    @property
    def stream(self):
        return self.lib.mr_stream(self.h)

    def set_use_graphs(self, mode):
        """CUDA-graph replay of ``submit_main_frame``'s launch sequence: 0 never, 1 for host rows (default), 2 always."""
        self.check(self.lib.mr_set_use_graphs(self.h, int(mode)))

    @property
    def graph_launches(self):
        return int(self.lib.mr_graph_launch_count(self.h))

    def normals_stats(self):
        """(tiles with valid pixels, tiles with 1 / 2 / 3 coordinates on the sample-by-sample route, residual pixels) of
        the normals covariance kernel (``mr_normals_stats``)."""
        out = (C.c_uint64 * 5)()
        self.check(self.lib.mr_normals_stats(self.h, out))
        return tuple(int(v) for v in out)

    @property
    def launches(self):
        return int(self.lib.mr_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_CTX = {}


def _ctx(width, height, device=0):
    key = (int(width), int(height), int(device))
    if key not in _CTX:
        _CTX[key] = Context(*key)
    return _CTX[key]


class Render:
    """``class Render`` (recon.hpp:93-99) backed by the CUDA rasteriser."""

    def __init__(self, width, height, device=0, ctx=None):
        self.ctx = ctx or _ctx(width, height, device)
        self.W, self.H = self.ctx.W, self.ctx.H

    def loadMesh(self, vertices, faces=None):
        if faces is None:  # accept a (vertices, faces) Mesh tuple like the reference's struct
            vertices, faces = vertices
        pv, kv = _ptr(vertices, np.float32, name="vertices")
        pf, kf = _ptr(faces, np.int32, name="faces")
        nv = int(kv.shape[0])
        nf = int(kf.shape[0])
        self.ctx.check(self.ctx.lib.mr_load_mesh(self.ctx.h, pv, nv, pf, nf))

    def depth(self, camera, out=None):
        pc, kc = _mat16(camera)
        if out is None:
            out = np.empty((self.H, self.W), np.float32)
        po, _ = _ptr(out, np.float32)
        self.ctx.check(self.ctx.lib.mr_depth(self.ctx.h, pc, po))
        return out

    def depthSamples(self, cameras, rows, cols):
        """Batched ``depth(viewer).at<float>(row, col)`` queries (heuristic.cpp:306-311,456): ``cameras`` is
        m x 4 x 4, ``rows`` / ``cols`` are m x n int32; returns m x n float32 without reading back depth maps."""
        cams = np.ascontiguousarray(np.asarray(cameras, np.float32).reshape(-1, 16))
        rows = np.ascontiguousarray(np.asarray(rows, np.int32).reshape(len(cams), -1))
        cols = np.ascontiguousarray(np.asarray(cols, np.int32).reshape(len(cams), -1))
        assert rows.shape == cols.shape
        out = np.empty(rows.shape, np.float32)
        self.ctx.check(self.ctx.lib.mr_depth_samples(self.ctx.h, cams.ctypes.data, len(cams), rows.ctypes.data, cols.ctypes.data,
                                                     rows.shape[1], out.ctypes.data))
        return out

    def projected(self, camera, frame, projector, out=None):
        pc, kc = _mat16(camera)
        pp, kp = _mat16(projector)
        pf, kf = _ptr(frame, np.uint8, (self.H, self.W) if not _is_torch(frame) else None, "frame")
        if out is None:
            out = np.empty((self.H, self.W, 3), np.uint8)
        po, _ = _ptr(out, np.uint8)
        self.ctx.check(self.ctx.lib.mr_projected(self.ctx.h, pc, pf, pp, po))
        return out


def spawnRender(width, height, device=0):
    """``spawnRender(hint)`` (recon.hpp:100): the render size is the clip size
    (heuristic.cpp:548-551)."""
    return Render(width, height, device)


def _hw(x):
    return int(x.shape[0]), int(x.shape[1])


def mixBackground(image, background, depth, device=0):
    """util.cpp:366-387.  ``depth`` (float32 H x W, NumPy or CUDA tensor) is modified in place."""
    H, W = _hw(background)
    ctx = _ctx(W, H, device)
    if not _is_torch(depth):
        assert depth.dtype == np.float32 and depth.flags.c_contiguous, "depth must be a contiguous float32 array (in/out)"
    pi, ki = _ptr(image, np.uint8)
    pb, kb = _ptr(background, np.uint8)
    pd, kd = _ptr(depth, np.float32)
    out = np.empty((H, W), np.uint8)
    ctx.check(ctx.lib.mr_mix_background(ctx.h, pi, pb, pd, out.ctypes.data))
    return out


def calculateFlow(prev, next, useFarneback=False, device=0, out=None):
    """flow.cpp:19-42 -> H x W x 4 float32 ``(u, v, variance, 0)``."""
    H, W = _hw(prev)
    ctx = _ctx(W, H, device)
    pp, kp = _ptr(prev, np.uint8)
    pn, kn = _ptr(next, np.uint8)
    if out is None:
        out = np.empty((H, W, 4), np.float32)
    po, _ = _ptr(out, np.float32)
    ctx.check(ctx.lib.mr_calculate_flow(ctx.h, pp, pn, int(bool(useFarneback)), po))
    return out


def flowRemap(flow, image, device=0):
    """util.cpp:390-403."""
    H, W = _hw(image)
    ctx = _ctx(W, H, device)
    stride = int(flow.shape[2])
    pf, kf = _ptr(flow, np.float32)
    pi, ki = _ptr(image, np.uint8)
    out = np.empty((H, W), np.uint8)
    ctx.check(ctx.lib.mr_flow_remap(ctx.h, pf, stride, pi, out.ctypes.data))
    return out


def compare(prev, next, device=0):
    """util.cpp:332-361."""
    H, W = _hw(prev)
    ctx = _ctx(W, H, device)
    pp, kp = _ptr(prev, np.uint8)
    pn, kn = _ptr(next, np.uint8)
    out = np.empty((H, W), np.float32)
    ctx.check(ctx.lib.mr_compare(ctx.h, pp, pn, out.ctypes.data))
    return out


def imageGradient(image, device=0):
    """util.cpp:465-479 (single-channel float32 input)."""
    H, W = _hw(image)
    ctx = _ctx(W, H, device)
    pi, ki = _ptr(image, np.float32)
    out = np.empty((H, W, 2), np.float32)
    ctx.check(ctx.lib.mr_image_gradient(ctx.h, pi, out.ctypes.data))
    return out


def extractCameraCenter(camera):
    """util.cpp:33-41 (returned dehomogenised, 3 floats)."""
    lib = load_library()
    pc, kc = _mat16(camera)
    out = np.empty(3, np.float32)
    rc = lib.mr_extract_camera_center(pc, out.ctypes.data)
    if rc:
        raise MeshReconError(rc, "mr_extract_camera_center")
    return out


def triangulatePixels(flows, mainCamera, cameras, depth, device=0):
    """util.cpp:167-329 -> M x 7 float32 rows ``(x, y, z, w, nx, ny, nz)`` in row-major pixel order."""
    H, W = _hw(depth)
    ctx = _ctx(W, H, device)
    S = len(flows)
    if not (1 <= S <= MR_MAX_SIDE) or len(cameras) != S:
        raise MeshReconError(-1, "flows and cameras must have the same length in 1..16")
    keep = [_ptr(f, np.float32) for f in flows]
    arr = (C.c_void_p * S)(*[k[0] for k in keep])
    pm, km = _mat16(mainCamera)
    cams = np.ascontiguousarray(np.stack([np.asarray(c, np.float32).reshape(16) for c in cameras]))
    pd, kd = _ptr(depth, np.float32)
    out = np.empty((H * W, 7), np.float32)
    m = C.c_int(0)
    ctx.check(ctx.lib.mr_triangulate_pixels(ctx.h, arr, S, pm, cams.ctypes.data, pd, out.ctypes.data, C.byref(m)))
    return out[:m.value].copy()


def process_main_frame(render, main_frame, main_camera, side_frames, side_cameras, out=None, want_host=True, async_copy=False):
    """One iteration of the reference's outer loop (recon.cpp:65-119), fused and
    device-resident.  Returns the M x 7 rows (NumPy) or, with ``want_host=False``, just M
    (rows stay on the device: ``Context.lib.mr_points_device``).  With ``async_copy`` (host
    ``out``) the D2H copy of the rows overlaps the next call: alternate two pinned ``out`` buffers
    and call ``render.ctx.wait_copies()`` before reading them; the call then returns M."""
    ctx = render.ctx
    S = len(side_frames)
    keep = [_ptr(f, np.uint8) for f in side_frames]
    arr = (C.c_void_p * S)(*[k[0] for k in keep])
    pf, kf = _ptr(main_frame, np.uint8)
    pm, km = _mat16(main_camera)
    cams = np.ascontiguousarray(np.stack([np.asarray(c, np.float32).reshape(16) for c in side_cameras]))
    m = C.c_int(0)
    if out is not None:
        po, ko = _ptr(out, np.float32)
    elif want_host:
        out = np.empty((ctx.H * ctx.W, 7), np.float32)
        po = out.ctypes.data
    else:
        po = None
    fn = ctx.lib.mr_process_main_frame_async if async_copy else ctx.lib.mr_process_main_frame
    ctx.check(fn(ctx.h, pf, pm, S, arr, cams.ctypes.data, po, C.byref(m)))
    if want_host and not async_copy and not _is_torch(out):
        return out[:m.value]
    return m.value


def submit_main_frame(render, main_frame, main_camera, side_frames, side_cameras, out, out_count=None):
    """Fully asynchronous ``mr_submit_main_frame``: returns immediately.  ``out`` (N x 7 float32) and
    ``out_count`` (1 x int32) are torch tensors on the device or in pinned host memory; frames likewise (or NumPy
    views of pinned tensors).  Call ``render.ctx.synchronize()`` before reading the results; keep every buffer alive
    until then."""
    ctx = render.ctx
    S = len(side_frames)
    keep = [_ptr(f, np.uint8) for f in side_frames]
    arr = (C.c_void_p * S)(*[k[0] for k in keep])
    pf, kf = _ptr(main_frame, np.uint8)
    pm, km = _mat16(main_camera)
    cams = np.ascontiguousarray(np.stack([np.asarray(c, np.float32).reshape(16) for c in side_cameras]))
    po, ko = _ptr(out, np.float32)
    pc = _ptr(out_count, np.int32)[0] if out_count is not None else None
    ctx.check(ctx.lib.mr_submit_main_frame(ctx.h, pf, pm, S, arr, cams.ctypes.data, po, pc))


def filterPoints(points, normals, radius, device=0, ctx=None, want_info=False):
    """``Heuristic::filterPoints(Mat &points, Mat &normals)`` (heuristic.cpp:55-176) on the GPU.  ``points`` n x 4
    homogeneous, ``normals`` n x 3 (or None), NumPy or CUDA tensors; ``radius`` is ``alphaVals.back() / 4`` and, as in
    the reference, bounds SQUARED distances.  Returns (points, normals, keep) of the survivors in ascending index order
    -- NumPy for NumPy inputs, CUDA tensors (views of fresh buffers) for CUDA inputs -- plus an info dict if asked."""
    ctx = ctx or _ctx(16, 16, device)
    n = int(points.shape[0])
    cnt = C.c_size_t(0)
    pp, kp = _ptr(points, np.float32)
    pn, kn = (_ptr(normals, np.float32) if normals is not None else (None, None))
    if _is_torch(points):
        import torch
        op = torch.empty((max(n, 1), 4), dtype=torch.float32, device=points.device)
        on = torch.empty((max(n, 1), 3), dtype=torch.float32, device=points.device) if normals is not None else None
        ok = torch.empty(max(n, 1), dtype=torch.int32, device=points.device)
        ptrs = (op.data_ptr(), on.data_ptr() if on is not None else None, ok.data_ptr())
    else:
        op = np.empty((max(n, 1), 4), np.float32)
        on = np.empty((max(n, 1), 3), np.float32) if normals is not None else None
        ok = np.empty(max(n, 1), np.int32)
        ptrs = (op.ctypes.data, on.ctypes.data if on is not None else None, ok.ctypes.data)
    ctx.check(ctx.lib.mr_filter_points(ctx.h, pp, pn, n, float(radius), ptrs[0], ptrs[1], ptrs[2], C.byref(cnt)))
    m = cnt.value
    res = (op[:m], on[:m] if on is not None else None, ok[:m])
    return res + (filter_info(ctx),) if want_info else res


def filter_rows(rows7, radius, device=0, ctx=None, out=None):
    """``mr_filter_rows``: the same filter on n x 7 point rows (x, y, z, w, nx, ny, nz).  Returns (rows, keep)."""
    ctx = ctx or _ctx(16, 16, device)
    n = int(rows7.shape[0])
    cnt = C.c_size_t(0)
    pr, kr = _ptr(rows7, np.float32)
    if _is_torch(rows7):
        import torch
        orows = out if out is not None else torch.empty((max(n, 1), 7), dtype=torch.float32, device=rows7.device)
        ok = torch.empty(max(n, 1), dtype=torch.int32, device=rows7.device)
        po, pk = orows.data_ptr(), ok.data_ptr()
    else:
        orows = out if out is not None else np.empty((max(n, 1), 7), np.float32)
        ok = np.empty(max(n, 1), np.int32)
        po, pk = (orows.data_ptr() if _is_torch(orows) else orows.ctypes.data), ok.ctypes.data
    ctx.check(ctx.lib.mr_filter_rows(ctx.h, pr, n, float(radius), po, pk, C.byref(cnt)))
    return orows[:cnt.value], ok[:cnt.value]


def filter_info(ctx, want_arrays=False, n=None):
    """Facts about the context's last filter call: edges (j < i neighbour pairs), power iterations, thinning rounds."""
    info = (C.c_longlong * 3)()
    if want_arrays:
        density, score = np.empty(n, np.float32), np.empty(n, np.float32)
        ctx.check(ctx.lib.mr_filter_info(ctx.h, info, density.ctypes.data, score.ctypes.data))
        return {"n_edges": info[0], "iters": info[1], "rounds": info[2], "density": density, "score": score}
    ctx.check(ctx.lib.mr_filter_info(ctx.h, info, None, None))
    return {"n_edges": info[0], "iters": info[1], "rounds": info[2]}


def ingest_frame(ctx, bgr, out=None, exposure=None):
    """``mr_ingest_frame``: what configuration.cpp:226-245 does to a decoded frame -- ``cv::resize(INTER_AREA)`` to the
    context's size when the frame is larger (integer or fractional factors, OpenCV's fast / general area paths), then ``cv::cvtColor(BGR2GRAY)``.  ``bgr``: h x w x 3 uint8
    (NumPy / pinned or CUDA tensor); returns the H x W gray frame (NumPy, or ``out`` -- e.g. a CUDA tensor)."""
    h, w = int(bgr.shape[0]), int(bgr.shape[1])
    pb, kb = _ptr(bgr, np.uint8)
    if out is None:
        out = np.empty((ctx.H, ctx.W), np.uint8)
    po, ko = _ptr(out, np.uint8)
    if exposure is not None:      # estimateExposure's channel weights (B, G, R) of this frame, configuration.cpp:417-425
        ex = (C.c_float * 3)(*[float(v) for v in exposure])
        ctx.check(ctx.lib.mr_ingest_frame_exposure(ctx.h, pb, w, h, ex, po))
    else:
        ctx.check(ctx.lib.mr_ingest_frame(ctx.h, pb, w, h, po))
    return out
