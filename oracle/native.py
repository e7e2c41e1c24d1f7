"""ctypes loader for ``oracle/recon_oracle.c`` (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librecon_oracle.so")
_lib = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "recon_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    L.orc_lu_inv4.argtypes = [f32p, f32p]
    L.orc_lu_inv4.restype = C.c_int
    L.orc_load_mesh.argtypes = [f32p, C.c_int, i32p, C.c_int, f32p]
    L.orc_raster.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, f32p, i32p]
    L.orc_dilate_shadow_gl.argtypes = [f32p, C.c_int, C.c_int]
    L.orc_projected.argtypes = [f32p, C.c_int, f32p, u8p, f32p, C.c_int, C.c_int, u8p]
    L.orc_mix_background.argtypes = [u8p, u8p, f32p, C.c_int, C.c_int, u8p]
    L.orc_triangulate_dense.argtypes = [C.POINTER(C.c_void_p), C.c_int, f32p, f32p, f32p, f32p, C.c_int, C.c_int,
                                        f32p, u8p, C.c_void_p]
    L.orc_camera_center.argtypes = [f32p, f32p]
    L.orc_jacobi3.argtypes = [f32p, f32p, f32p]
    L.orc_pca_normal.argtypes = [f32p, C.c_int, f32p, f32p]
    L.orc_normals_compact.argtypes = [f32p, u8p, C.c_int, C.c_int, f32p, f32p, C.c_int, f32p, C.c_void_p]
    L.orc_normals_compact.restype = C.c_int
    _lib = L
    return L
