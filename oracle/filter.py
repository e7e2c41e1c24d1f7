"""ctypes front end of ``oracle/filter_oracle.cpp`` -- the CPU restatement of ``Heuristic::filterPoints``
(heuristic.cpp:55-176).  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfilter_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "filter_oracle.cpp")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(_SO)
        vp = C.c_void_p
        L.orc_filter_points.argtypes = [vp, C.c_int, C.c_float, C.c_int, C.c_int, vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_longlong),
                                        vp, vp, vp, C.c_longlong]
        L.orc_filter_points.restype = C.c_int
        L.orc_sortidx_desc_stdsort.argtypes = [vp, C.c_int, vp]
        L.orc_sortidx_desc_stdsort.restype = None
        L.orc_seqsum.argtypes = [vp, C.c_longlong]
        L.orc_seqsum.restype = C.c_double
        _lib = L
    return _lib


def filter_points(points4, radius, tie_mode=0, brute=False, want_table=False):
    """Returns a dict: keep (ascending surviving indices), density, score (raw score of the last power iteration),
    iters, n_edges, and with ``want_table`` blocks / nb_idx / nb_w (the j < i neighbour table)."""
    p = np.ascontiguousarray(points4, np.float32).reshape(-1, 4)
    n = len(p)
    keep = np.empty(max(n, 1), np.int32)
    density = np.empty(max(n, 1), np.float32)
    score = np.empty(max(n, 1), np.float32)
    iters, edges = C.c_int(0), C.c_longlong(0)
    blocks = np.zeros(n + 1, np.int64)
    L = lib()
    m = L.orc_filter_points(p.ctypes.data, n, float(radius), int(tie_mode), int(bool(brute)), keep.ctypes.data, density.ctypes.data,
                            score.ctypes.data, C.byref(iters), C.byref(edges), blocks.ctypes.data, None, None, 0)
    if m < 0:
        raise ValueError("orc_filter_points: bad arguments")
    out = {"keep": keep[:m].copy(), "density": density[:n], "score": score[:n], "iters": iters.value, "n_edges": edges.value}
    if want_table:
        nb_idx = np.empty(max(edges.value, 1), np.int32)
        nb_w = np.empty(max(edges.value, 1), np.float32)
        L.orc_filter_points(p.ctypes.data, n, float(radius), int(tie_mode), int(bool(brute)), None, None, None, None, None, None,
                            nb_idx.ctypes.data, nb_w.ctypes.data, edges.value)
        out.update(blocks=blocks, nb_idx=nb_idx[:edges.value], nb_w=nb_w[:edges.value])
    return out


def sortidx_desc_stdsort(values):
    v = np.ascontiguousarray(values, np.float32)
    out = np.empty(len(v), np.int32)
    lib().orc_sortidx_desc_stdsort(v.ctypes.data, len(v), out.ctypes.data)
    return out


def seqsum(terms):
    t = np.ascontiguousarray(terms, np.float32)
    return float(lib().orc_seqsum(t.ctypes.data, len(t)))
