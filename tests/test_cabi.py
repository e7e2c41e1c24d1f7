"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/meshrecon_b200.h declares; without a GPU it fails loudly (no CPU fallback)."""
import os
import re

import pytest

import mesh_reconstruction_b200 as mr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "meshrecon_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = mr.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/meshrecon_b200.h but not exported"
    assert lib.mr_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mr.MeshReconError) as e:
        mr.spawnRender(64, 48)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mesh_reconstruction_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "recon_oracle" not in txt or f.endswith((".cu", ".cuh")) and "oracle/recon_oracle.c" in txt, f


def test_extract_camera_center_is_host_side():
    import numpy as np
    from mesh_reconstruction_b200 import synth
    sc = synth.make_scene(64, 48, 3)
    c = mr.extractCameraCenter(sc.cameras[0])
    assert np.allclose(c, sc.cam2world[0][:3, 3], rtol=1e-4, atol=1e-5)


def _write_case(path, sc, frames, fa, sides):
    import numpy as np
    with open(path, "wb") as f:
        np.array([sc.width, sc.height, len(sc.vertices), len(sc.faces), len(sides)], np.int32).tofile(f)
        sc.vertices.astype(np.float32).tofile(f)
        sc.faces.astype(np.int32).tofile(f)
        sc.cameras[fa].astype(np.float32).tofile(f)
        for s in sides:
            sc.cameras[s].astype(np.float32).tofile(f)
        frames[fa].tofile(f)
        for s in sides:
            frames[s].tofile(f)


def test_cpp_host_mirror_fails_loudly_without_gpu(tmp_path):
    """The C++ mirror of recon.hpp (recon_b200.hpp) drives recon.cpp's loop through the C ABI; on a
    box without a GPU it must stop with the library's error, not compute on the CPU."""
    import subprocess
    import torch
    from mesh_reconstruction_b200 import synth
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = os.path.join(ROOT, "mesh_reconstruction_b200", "host_mirror_test")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    sc = synth.make_scene(64, 48, 3, step=0.1, mesh_res=4)
    frames = sc.frames()
    _write_case(tmp_path / "in.bin", sc, frames, 1, [0, 2])
    p = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert p.returncode == 3 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_matches_oracle(tmp_path):
    import subprocess
    import numpy as np
    from mesh_reconstruction_b200 import synth
    from oracle.pipeline import process_main_frame
    from oracle.render import RenderOracle
    exe = os.path.join(ROOT, "mesh_reconstruction_b200", "host_mirror_test")
    sc = synth.make_scene(160, 120, 3, step=0.15, mesh_err=0.03, mesh_res=8)
    frames = sc.frames()
    _write_case(tmp_path / "in.bin", sc, frames, 1, [0, 2])
    p = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    raw = np.fromfile(tmp_path / "out.bin", np.uint8)
    M = int(raw[:4].view(np.int32)[0])
    got = raw[4:4 + M * 28].view(np.float32).reshape(M, 7)
    rest = raw[4 + M * 28:]
    K = int(rest[:4].view(np.int32)[0])
    radius = float(rest[4:8].view(np.float32)[0])
    fp = rest[8:8 + K * 16].view(np.float32).reshape(K, 4)
    fn = rest[8 + K * 16:8 + K * 28].view(np.float32).reshape(K, 3)
    ro = RenderOracle(160, 120)
    ro.loadMesh(sc.vertices, sc.faces)
    ref = process_main_frame(ro, frames, sc.cameras, 1, [0, 2])
    assert got.shape == ref.shape
    ok = ~np.isnan(ref).any(1)
    assert np.array_equal(ok, ~np.isnan(got).any(1))
    err = np.abs(got[ok, :3] / got[ok, 3:4] - ref[ok, :3] / ref[ok, 3:4]).max()
    assert err <= 1e-4 * sc.scale
    assert np.array_equal(got, ref, equal_nan=True)
    # hint.filterPoints(points, normals) through the C++ mirror == the filter oracle on the same cloud
    from oracle import filter as ofilter
    keep = ofilter.filter_points(ref[:, :4], radius)["keep"]
    assert 0 < K < M and K == len(keep)
    assert np.array_equal(fp, ref[keep, :4], equal_nan=True) and np.array_equal(fn, ref[keep, 4:7], equal_nan=True)


def test_register_jacobi_equals_opencv_jacobi():
    """csrc/jacobi3.cuh (n = 3 specialisation held in registers, incl. the stale pivot bookkeeping)
    against the oracle's line-by-line restatement of OpenCV's generic JacobiImpl_ and cv2.eigen."""
    import ctypes as C
    import cv2
    import numpy as np
    from oracle import native
    lib = mr.load_library()
    lib.mr_debug_jacobi3.argtypes = [C.c_void_p] * 3
    rng = np.random.default_rng(0)
    L = native.lib()
    nbit = 0
    for t in range(3000):
        basis = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        sig = np.array([1.0, 10 ** rng.uniform(-3, 0), 10 ** rng.uniform(-6, 0)]) * 10 ** rng.uniform(-5, 1)
        if t % 7 == 0:
            sig[1] = sig[0]                      # repeated eigenvalues
        cov = ((basis * sig) @ basis.T).astype(np.float32)
        cov = ((cov + cov.T) * np.float32(0.5)).astype(np.float32)
        c6 = np.array([cov[0, 0], cov[0, 1], cov[0, 2], cov[1, 1], cov[1, 2], cov[2, 2]], np.float32)
        w, v = np.empty(3, np.float32), np.empty(9, np.float32)
        assert lib.mr_debug_jacobi3(c6.ctypes.data, w.ctypes.data, v.ctypes.data) == 0
        a = np.ascontiguousarray(cov.ravel().copy())
        w2, v2 = np.empty(3, np.float32), np.empty(9, np.float32)
        L.orc_jacobi3(a, w2, v2)
        nbit += int(np.array_equal(w, w2) and np.array_equal(v, v2))
        ok, evals, evecs = cv2.eigen(cov)                      # the OpenCV binary itself
        assert np.array_equal(evals.ravel(), w) and np.array_equal(evecs.ravel(), v), t
    assert nbit == 3000          # and bit-identical to the oracle's generic restatement
