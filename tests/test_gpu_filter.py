"""GPU parity tests of mr_filter_points / mr_filter_rows (Heuristic::filterPoints, heuristic.cpp:55-176) against the CPU
restatement oracle/filter_oracle.cpp: surviving indices, converged densities, raw scores, iteration and neighbour-pair
counts must all be bit-identical -- including the `sum` / `change` double accumulations over all pairs, which the CUDA
path evaluates in parallel (mr_debug_seqsum)."""
import ctypes as C
import time

import numpy as np
import pytest

import mesh_reconstruction_b200 as mr
from oracle import filter as ofilter
from tests.test_oracle_filter import cloud

pytestmark = pytest.mark.gpu
f32 = np.float32


def _seqsum_gpu(ctx, t):
    out = C.c_double(0)
    t = np.ascontiguousarray(t, f32)
    ctx.check(ctx.lib.mr_debug_seqsum(ctx.h, t.ctypes.data, len(t), C.byref(out)))
    return out.value


def test_seqsum_is_bit_identical_to_the_sequential_loop():
    ctx = mr.api.Context(16, 16)
    rng = np.random.default_rng(0)
    cases = [np.zeros(0, f32), np.array([0.3], f32), rng.random(5).astype(f32), rng.random(2048).astype(f32), rng.random(2049).astype(f32),
             rng.random(300000).astype(f32) * 4,                                   # like the `sum` terms
             (rng.random(1000003) ** 12).astype(f32),                              # many tiny terms: ties and sub-ulp terms
             (rng.random(400000).astype(f32) * 1e-6) ** 2,                         # like the `change` terms near convergence
             np.full(700000, 0.1, f32), np.full(70000, 2.0 ** -30, f32),           # constant terms: systematic ties
             np.concatenate([np.full(5000, 1e-20, f32), rng.random(5000).astype(f32), np.full(5000, 1e10, f32), rng.random(5000).astype(f32)]),
             np.where(rng.random(500000) < 0.5, 0, rng.random(500000)).astype(f32),
             (2.0 ** rng.integers(-40, 3, 200000)).astype(f32)]                    # powers of two: exact halves everywhere
    for k, t in enumerate(cases):
        ref = ofilter.seqsum(t)
        got = _seqsum_gpu(ctx, t)
        assert got == ref, (k, len(t), got, ref, got - ref)
    t = rng.random(10000).astype(f32)
    t[77] = np.nan
    assert np.isnan(_seqsum_gpu(ctx, t))
    t[77] = np.inf
    assert _seqsum_gpu(ctx, t) == np.inf


def _check(p, radius, normals=None, brute=False):
    ref = ofilter.filter_points(p, radius, brute=brute)
    ctx = mr.api.Context(16, 16)
    op, on, keep = mr.filterPoints(p, normals, radius, ctx=ctx)
    info = mr.api.filter_info(ctx, want_arrays=True, n=len(p)) if len(p) else {"n_edges": 0, "iters": 0}
    if len(p):
        assert info["n_edges"] == ref["n_edges"], (info["n_edges"], ref["n_edges"])
        assert info["iters"] == ref["iters"], (info["iters"], ref["iters"])
        assert np.array_equal(info["density"], ref["density"], equal_nan=True), np.sum(info["density"] != ref["density"])
        assert np.array_equal(info["score"], ref["score"], equal_nan=True)
    assert np.array_equal(keep, ref["keep"]), (len(keep), len(ref["keep"]))
    assert np.array_equal(op, p[ref["keep"]])
    if normals is not None:
        assert np.array_equal(on, normals[ref["keep"]])
    return ref, info


@pytest.mark.parametrize("n,radius", [(0, 0.01), (1, 0.01), (2, 10.0), (300, 0.02), (5000, 0.004), (20000, 0.0016), (20000, 1e-9), (3000, 100.0)])
def test_filter_points_small(n, radius):
    p = cloud(n, seed=n + 1)
    nrm = np.random.default_rng(n).normal(size=(n, 3)).astype(f32)
    ref, info = _check(p, radius, nrm, brute=n <= 5000)
    if n == 20000 and radius > 1e-6:
        assert 0 < len(ref["keep"]) < n and ref["n_edges"] > 5 * n


def test_filter_points_degenerate_inputs():
    rng = np.random.default_rng(4)
    p = cloud(4000, seed=9)
    p[10, 3] = 0.0                     # w = 0 -> inf / NaN coordinates
    p[11] = np.nan
    p[12, 0] = np.inf
    p[100:140] = p[99]                 # a pile of exact duplicates (distance 0, weight 1)
    p[200:230, :3] = 0.0               # points at the origin
    _check(p, 0.01, rng.normal(size=(4000, 3)).astype(f32), brute=True)
    # all points identical
    q = np.tile(np.array([[1, 2, 3, 1]], f32), (500, 1))
    _check(q, 0.5, None, brute=True)


def test_filter_points_million():
    """10^6 points on a surface with outliers: indices bit-identical to the oracle."""
    n = 1000000
    p = cloud(n, seed=5, outliers=0.03)
    radius = 5e-5                      # squared units: ~ 7e-3 Euclidean -> tens of neighbours per point
    t0 = time.perf_counter()
    ref = ofilter.filter_points(p, radius)
    t_cpu = time.perf_counter() - t0
    ctx = mr.api.Context(16, 16)
    import torch
    pd = torch.from_numpy(p).cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    op, _, keep = mr.filterPoints(pd, None, radius, ctx=ctx)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    info = mr.api.filter_info(ctx)
    print(f"filterPoints 1e6: {ref['n_edges']} pairs, {ref['iters']} iterations, {len(ref['keep'])} survivors; oracle {t_cpu:.2f} s, "
          f"GPU {t_gpu * 1e3:.1f} ms ({info['rounds']} thinning rounds)")
    assert info["n_edges"] == ref["n_edges"] and info["iters"] == ref["iters"]
    assert np.array_equal(keep.cpu().numpy(), ref["keep"])
    assert np.array_equal(op.cpu().numpy(), p[ref["keep"]])
    assert ref["n_edges"] > 10 * n and 0.5 * n < len(ref["keep"]) < n


def test_filter_rows_of_the_path():
    """The cloud the path produces (two main frames' rows, device resident) through mr_filter_rows."""
    import torch
    from mesh_reconstruction_b200 import synth
    W, H = 320, 240
    sc = synth.make_scene(W, H, 4, seed=2, step=0.12, mesh_err=0.03, mesh_res=10)
    frames = sc.frames()
    r = mr.Render(W, H, ctx=mr.api.Context(W, H))
    r.loadMesh(sc.vertices, sc.faces)
    rows = np.concatenate([mr.process_main_frame(r, frames[i], sc.cameras[i], [frames[i + 1]], [sc.cameras[i + 1]]).copy() for i in (0, 1, 2)])
    rows = rows[np.isfinite(rows).all(1)]
    d = rows[:, :3] / rows[:, 3:4]
    radius = float(((d.max(0) - d.min(0)).max() / 150) ** 2)
    ref = ofilter.filter_points(rows[:, :4], radius)
    got_rows, keep = mr.filter_rows(torch.from_numpy(rows).cuda(), radius, ctx=r.ctx)
    assert np.array_equal(keep.cpu().numpy(), ref["keep"])
    assert np.array_equal(got_rows.cpu().numpy(), rows[ref["keep"]])
    got_rows_h, keep_h = mr.filter_rows(rows, radius, ctx=r.ctx)          # host buffers through the same entry point
    assert np.array_equal(keep_h, ref["keep"]) and np.array_equal(got_rows_h, rows[ref["keep"]])
    assert 0 < len(ref["keep"]) < len(rows)


def test_too_many_pairs_is_refused_not_wrapped():
    """More than 2^31 neighbour pairs (a radius far too large for the cloud; the reference's own `int` table would overflow):
    the call must fail with MR_EINVAL -- the pair counts are scanned in 64 bits, so the total cannot wrap to a small or negative
    number and be used to size the tables."""
    rng = np.random.default_rng(5)
    n = 70000                                                   # n^2 / 2 = 2.45e9 pairs when every point sees every other
    p = np.concatenate([rng.random((n, 3)).astype(np.float32) * 0.01, np.ones((n, 1), np.float32)], 1)
    ctx = mr.api.Context(16, 16)
    with pytest.raises(mr.MeshReconError) as e:
        mr.filterPoints(p, np.zeros((n, 3), np.float32), 1.0, ctx=ctx)
    assert "2^31" in str(e.value)
    # the context stays usable
    q = p[:2000]
    _, _, keep = mr.filterPoints(q, np.zeros((len(q), 3), np.float32), 1e-6, ctx=ctx)
    assert len(keep) > 0
