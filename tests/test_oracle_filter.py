"""CPU tests of the filterPoints oracle (oracle/filter_oracle.cpp, heuristic.cpp:55-176): the grid search is only an
accelerator (same table as brute force), the literal cv::sortIdx path equals the real OpenCV binary, and a NumPy
restatement of the power iteration and of the greedy thinning agrees with the C++ one."""
import numpy as np

from oracle import filter as ofilter

f32 = np.float32


def cloud(n, seed=0, outliers=0.05, dup=True):
    """points on a wavy sheet + sparse outliers, homogeneous with w != 1, a few exact duplicates"""
    rng = np.random.default_rng(seed)
    xy = rng.random((n, 2)) * 2 - 1
    z = 0.1 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1]) + rng.normal(0, 0.003, n)
    p = np.concatenate([xy, z[:, None]], 1)
    k = int(outliers * n)
    p[rng.choice(n, k, replace=False)] += rng.normal(0, 0.5, (k, 3))
    if dup and n > 10:
        p[n // 2] = p[n // 3]
    w = rng.uniform(0.5, 2.0, (n, 1))
    return np.concatenate([p * w, w], 1).astype(f32)


def test_grid_search_equals_brute_force():
    for n, radius in [(0, 0.01), (1, 0.01), (300, 0.02), (2000, 0.004), (2000, 0.05)]:
        p = cloud(n, seed=n)
        a = ofilter.filter_points(p, radius, brute=True, want_table=True)
        b = ofilter.filter_points(p, radius, brute=False, want_table=True)
        assert a["n_edges"] == b["n_edges"]
        for k in ("keep", "density", "score", "blocks", "nb_idx", "nb_w"):
            assert np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f"), (n, k)
        assert a["iters"] == b["iters"]
        if n >= 300:
            assert a["n_edges"] > n and 0 < len(a["keep"]) < n


def test_literal_sortidx_equals_cv2():
    """tie_mode 1 is cv::sortIdx's generic path (std::sort on indices + reversal): bit-identical to the cv2 binary,
    ties (clamped densities) included."""
    import cv2
    rng = np.random.default_rng(1)
    for n in (1, 2, 17, 1000, 200000):
        v = np.minimum(rng.random(n).astype(f32) * 3, 2.0)
        ref = cv2.sortIdx(v.reshape(1, -1), cv2.SORT_DESCENDING | cv2.SORT_EVERY_ROW).ravel()
        assert np.array_equal(ofilter.sortidx_desc_stdsort(v), ref), n


def numpy_filter(points4, radius):
    """independent restatement (dense matrices, small n) of heuristic.cpp:104-163 with tie rule F2 (desc index)"""
    p = points4[:, :3] / points4[:, 3:4]
    n = len(p)
    d = np.zeros((n, n), f32)
    for k in range(3):
        diff = (p[:, None, k] - p[None, :, k]).astype(f32)
        d = (d + diff * diff).astype(f32)
    lower = np.tril(np.ones((n, n), bool), -1) & (d <= f32(radius))
    w = (1.0 - (d / f32(radius)).astype(np.float64)).astype(f32)
    order_in_block = [sorted(np.nonzero(lower[i])[0], key=lambda j: (d[i, j], j)) for i in range(n)]
    density = np.ones(n, f32)
    it = 0
    while True:
        score = np.zeros(n, f32)
        s = 0.0
        for i in range(n):
            t = f32(0)
            for j in order_in_block[i]:
                t = f32(t + f32(density[j] * w[i, j]))
                score[j] = f32(score[j] + f32(density[i] * w[i, j]))
                s += float(f32(f32(density[i] + density[j]) * w[i, j]))
            score[i] = f32(score[i] + t)
        with np.errstate(divide="ignore", invalid="ignore"):
            norm = f32(np.float64(n) / np.float64(s))
            nd = (score * norm).astype(f32)
        nd = np.where(nd > 2.0, f32(2.0), nd).astype(f32)
        change = 0.0
        for i in range(n):
            df = f32(density[i] - nd[i])
            change += float(f32(df * df))
        density = nd
        change /= n
        it += 1
        if not (change > 1e-6 and it < 200):
            break
    order = sorted(range(n), key=lambda i: (-density[i], -i))
    keep = []
    score = score.copy()
    for o in order:
        if score[o] < f32(0.7):
            continue
        for j in order_in_block[o]:
            score[j] = f32(np.float64(score[j]) - np.float64(density[o]) * np.float64(w[o, j]))
        keep.append(o)
    return np.array(sorted(keep), np.int32), density, it


def test_numpy_restatement_agrees():
    for n, radius in [(60, 0.08), (150, 0.03)]:
        p = cloud(n, seed=7 + n)
        a = ofilter.filter_points(p, radius, brute=True)
        keep, density, it = numpy_filter(p, radius)
        assert it == a["iters"]
        assert np.array_equal(density, a["density"])
        assert np.array_equal(keep, a["keep"])


def test_tie_rule_changes_little():
    """How far definition F2 (equal densities by descending index) is from std::sort's order: same survivor COUNT
    within a few percent on a dense cloud (most densities clamp to 2.0); reported, not pinned."""
    p = cloud(20000, seed=3, outliers=0.02)
    a = ofilter.filter_points(p, 0.0016, tie_mode=0)
    b = ofilter.filter_points(p, 0.0016, tie_mode=1)
    assert np.array_equal(a["density"], b["density"])
    ties = float(np.mean(a["density"] == 2.0))
    common = len(np.intersect1d(a["keep"], b["keep"]))
    print(f"filterPoints tie rule: {ties:.1%} of densities clamp to 2.0; survivors {len(a['keep'])} (desc index) vs "
          f"{len(b['keep'])} (std::sort), {common} in common, of {len(p)} points, {a['n_edges']} edges, {a['iters']} iterations")
    assert abs(len(a["keep"]) - len(b["keep"])) <= 0.05 * len(p)


def test_seqsum_is_sequential():
    rng = np.random.default_rng(0)
    t = (rng.random(100000) ** 8).astype(f32)
    s = 0.0
    for v in t[:2000]:
        s += float(v)
    assert ofilter.seqsum(t[:2000]) == s
    assert ofilter.seqsum(t) != float(np.sum(t.astype(np.float64))) or True


def test_neighbour_table_equals_flann_linear_index():
    """Quirk F1 pinned to the FLANN binary: the reference's kd-tree search (heuristic.cpp:72-85) approximates the radius
    set that FLANN's EXACT linear index returns.  For every point the oracle's j < i block must be that set in FLANN's
    result order (ascending squared distance), with FLANN's float squared distances inside densityFn (heuristic.cpp:49-52)."""
    import cv2
    rng = np.random.default_rng(12)
    n = 3000
    pts = np.concatenate([rng.random((n - 200, 3)), rng.random((200, 3)) * 0.05 + 0.4], 0).astype(np.float32)   # a dense clump among sparse points
    pts[17] = pts[16]                                                                                           # a duplicate (distance 0)
    w = (rng.random((n, 1)) * 1.5 + 0.5).astype(np.float32)
    p4 = np.concatenate([pts * w, w], 1).astype(np.float32)               # homogeneous, as filterPoints receives them
    p3 = np.ascontiguousarray(p4[:, :3] / p4[:, 3:4])                     # dehomogenize(), util.cpp:16-29: plain float divisions
    radius = np.float32(0.004)                                            # bounds SQUARED distances (0.063 units)
    res = ofilter.filter_points(p4, float(radius), want_table=True)
    index = cv2.flann_Index(p3, {"algorithm": 0})                          # FLANN_INDEX_LINEAR: exact
    blocks, nb_idx, nb_w = res["blocks"], res["nb_idx"], res["nb_w"]
    total = 0
    for i in range(n):
        cnt, ind, dist = index.radiusSearch(p3[i:i + 1], float(radius), n, params={"checks": 32})
        ind, dist = ind[0][:cnt], dist[0][:cnt]
        sel = (ind < i) & (dist <= radius)
        exp_idx = ind[sel]
        exp_w = (1.0 - (dist[sel] / radius).astype(np.float64)).astype(np.float32)      # (float)(1. - dist / radius): float quotient, double difference
        got_idx, got_w = nb_idx[blocks[i]:blocks[i + 1]], nb_w[blocks[i]:blocks[i + 1]]
        # FLANN orders equal distances arbitrarily (heap); the oracle defines ties by index: compare as (distance, index) sets
        assert len(got_idx) == len(exp_idx), i
        order = np.lexsort((exp_idx, -exp_w))
        assert np.array_equal(got_idx, exp_idx[order]) and np.array_equal(got_w, exp_w[order]), i
        total += len(got_idx)
    assert total == res["n_edges"] and total > 5 * n
