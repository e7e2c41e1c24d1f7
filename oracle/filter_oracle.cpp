// filter_oracle.cpp -- CPU restatement of Heuristic::filterPoints (heuristic.cpp:55-176).
// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the parity checker of mr_filter_points and the CPU baseline of its
// bench line; the product never links or loads it.
//
// Follows the reference statement by statement: dehomogenize (util.cpp:16-29), the neighbour table restricted to j < i
// (heuristic.cpp:70-98), the clamped power iteration with its float / double mix (heuristic.cpp:104-138) and the greedy
// thinning on the RAW scores of the last iteration (heuristic.cpp:140-163).  Two things the reference leaves to its
// libraries are DEFINED here:
//   F1  neighbour set.  The reference asks FLANN for radiusSearch() on 4 randomised kd-trees with 32 checks
//       (cv::flann::GenericIndex<L2_Simple>(KDTreeIndexParams()), SearchParams()): an approximate, seed dependent set
//       ("because FLANN is randomized", heuristic.cpp:88).  The oracle takes the EXACT set the search approximates:
//       every j with L2_Simple(p_i, p_j) <= radius, L2_Simple being FLANN's float accumulation ((dx*dx) + dy*dy) + dz*dz of
//       the SQUARED distance -- so `radius` is compared with squared distances and densityFn sees squared distances,
//       exactly as in the reference -- in FLANN's result order: ascending (distance, index)
//       (RadiusUniqueResultSet / sortAndCopy).
//   F2  order of equal densities.  cv::sortIdx(SORT_DESCENDING) is std::sort on indices with a value-only comparator
//       followed by a reversal: equal keys (every density clamped to 2.0!) come out in whatever order libstdc++'s
//       introsort leaves them.  tie_mode 0 (the definition the CUDA path implements): descending density, equal
//       densities by DESCENDING index (what a stable sort + the reversal give).  tie_mode 1: the literal std::sort +
//       reversal, pinned against cv2.sortIdx in tests/test_oracle_filter.py, to show what the definition changes.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <unordered_map>
#include <vector>

namespace {

struct Nb { int j; float d; };

inline float l2_simple(const float *a, const float *b)
{
    float result = 0.f;
    for (int k = 0; k < 3; k++) { float diff = a[k] - b[k]; result += diff * diff; }
    return result;
}

inline float density_fn(float dist, float radius) { return (float)(1. - dist / radius); }   // heuristic.cpp:49-52

struct LessThanIdx {      // OpenCV's comparator of sortIdx_
    const float *arr;
    bool operator()(int a, int b) const { return arr[a] < arr[b]; }
};

}  // namespace

extern "C" {

// points4: n x 4 homogeneous float32.  radius: alphaVals.back() / 4 (heuristic.cpp:63).  brute: 1 = O(n^2) search.
// Outputs (each may be NULL): keep[<= n] ascending surviving indices, density[n] / score[n] after the power iteration
// (score = RAW score of the last iteration, before thinning), *iters, *n_edges, blocks[n + 1], and the neighbour table
// nb_idx / nb_w[cap_edges] when cap_edges >= *n_edges.  Returns the number of survivors, or -1 on bad arguments.
int orc_filter_points(const float *points4, int n, float radius, int tie_mode, int brute, int32_t *keep, float *density_out,
                      float *score_out, int *iters_out, long long *n_edges_out, long long *blocks_out, int32_t *nb_idx, float *nb_w,
                      long long cap_edges)
{
    if (n < 0 || (n > 0 && !points4)) return -1;
    const int pointCount = n;
    std::vector<float> p3((size_t)n * 3);
    for (int i = 0; i < n; i++) {                       // dehomogenize
        const float *inp = points4 + 4 * (size_t)i;
        p3[3 * (size_t)i + 0] = inp[0] / inp[3];
        p3[3 * (size_t)i + 1] = inp[1] / inp[3];
        p3[3 * (size_t)i + 2] = inp[2] / inp[3];
    }
    // == neighbour table (j < i), blocks in ascending (distance, index) order ==
    std::vector<long long> blocks((size_t)n + 1, 0);
    std::vector<int> nbi;
    std::vector<float> nbw;
    {
        std::vector<Nb> found;
        auto flush = [&](int i) {
            std::sort(found.begin(), found.end(), [](const Nb &a, const Nb &b) { return a.d < b.d || (a.d == b.d && a.j < b.j); });
            blocks[i] = (long long)nbi.size();
            for (const Nb &f : found) { nbi.push_back(f.j); nbw.push_back(density_fn(f.d, radius)); }
            found.clear();
        };
        if (brute) {
            for (int i = 0; i < n; i++) {
                for (int j = 0; j < i; j++) {
                    float d = l2_simple(&p3[3 * (size_t)i], &p3[3 * (size_t)j]);
                    if (d <= radius) found.push_back({j, d});
                }
                flush(i);
            }
        } else {
            // uniform grid, cell edge a little above the Euclidean radius; only an accelerator: the accepted set is decided by
            // the same float comparison as above
            const double cell = std::sqrt((double)radius) * 1.0001 + 1e-30;
            auto key = [&](const float *p, int dx, int dy, int dz, bool &ok) {
                long long c[3];
                const int d[3] = {dx, dy, dz};
                ok = true;
                for (int k = 0; k < 3; k++) {
                    double v = std::floor((double)p[k] / cell);
                    if (!(std::fabs(v) < 1e6)) { ok = false; return 0ll; }     // NaN / inf / far away: no neighbours
                    c[k] = (long long)v + d[k] + (1 << 20);
                }
                return (c[0] << 42) | (c[1] << 21) | c[2];
            };
            std::unordered_map<long long, std::vector<int>> grid;
            grid.reserve((size_t)n);
            for (int i = 0; i < n; i++) {
                bool ok;
                for (int dx = -1; dx <= 1 && radius >= 0; dx++)
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dz = -1; dz <= 1; dz++) {
                            long long k = key(&p3[3 * (size_t)i], dx, dy, dz, ok);
                            if (!ok) continue;
                            auto it = grid.find(k);
                            if (it == grid.end()) continue;
                            for (int j : it->second) {          // only points inserted so far: j < i
                                float d = l2_simple(&p3[3 * (size_t)i], &p3[3 * (size_t)j]);
                                if (d <= radius) found.push_back({j, d});
                            }
                        }
                flush(i);
                long long k0 = key(&p3[3 * (size_t)i], 0, 0, 0, ok);
                if (ok) grid[k0].push_back(i);
            }
        }
        blocks[n] = (long long)nbi.size();
    }
    const long long E = (long long)nbi.size();
    if (n_edges_out) *n_edges_out = E;
    if (blocks_out) memcpy(blocks_out, blocks.data(), sizeof(long long) * ((size_t)n + 1));
    if (nb_idx && nb_w && cap_edges >= E) {
        memcpy(nb_idx, nbi.data(), sizeof(int) * (size_t)E);
        memcpy(nb_w, nbw.data(), sizeof(float) * (size_t)E);
    }
    // == local density: power iteration with clamping (heuristic.cpp:104-138) ==
    std::vector<float> density((size_t)n, 1.f), score((size_t)n, 0.f);
    double change;
    int densityIterationNo = 0;
    do {
        for (int i = 0; i < pointCount; i++) score[i] = 0.;
        double sum = 0.;
        for (int i = 0; i < pointCount; i++) {
            float densityTemp = 0.0;
            for (long long j = blocks[i]; j < blocks[i + 1]; j++) {
                densityTemp += density[nbi[j]] * nbw[j];
                score[nbi[j]] += density[i] * nbw[j];
                sum += (density[i] + density[nbi[j]]) * nbw[j];
            }
            score[i] += densityTemp;
        }
        float normalizer = pointCount / sum;
        change = 0.;
        for (int i = 0; i < pointCount; i++) {
            float normalizedDensity = score[i] * normalizer;
            if (normalizedDensity > 2.) normalizedDensity = 2.;
            float diff = density[i] - normalizedDensity;     // pow2(float)
            change += diff * diff;
            density[i] = normalizedDensity;
        }
        change /= pointCount;
        densityIterationNo += 1;
    } while (change > 1e-6 && densityIterationNo < 200);
    if (iters_out) *iters_out = densityIterationNo;
    if (density_out && n) memcpy(density_out, density.data(), sizeof(float) * (size_t)n);
    if (score_out && n) memcpy(score_out, score.data(), sizeof(float) * (size_t)n);
    // == greedy thinning along descending density (heuristic.cpp:140-163) ==
    const float densityLimit = .7f;
    std::vector<int> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    if (tie_mode == 1) {
        std::sort(order.begin(), order.end(), LessThanIdx{density.data()});       // cv::sortIdx, generic path
        std::reverse(order.begin(), order.end());                                    // (swaps j <-> len-1-j, same thing)
    } else {
        std::sort(order.begin(), order.end(), [&](int a, int b) { return density[a] > density[b] || (density[a] == density[b] && a > b); });
    }
    int writeIndex = 0;
    for (int i = 0; i < pointCount; i++) {
        int ord = order[i];
        if (score[ord] < densityLimit) continue;
        double localDensity = density[ord];
        for (long long j = blocks[ord]; j < blocks[ord + 1]; j++) score[nbi[j]] -= localDensity * nbw[j];
        if (i > writeIndex) order[writeIndex] = order[i];
        writeIndex++;
    }
    std::sort(order.begin(), order.begin() + writeIndex);
    if (keep) for (int i = 0; i < writeIndex; i++) keep[i] = order[i];
    return writeIndex;
}

// cv::sortIdx(row, SORT_DESCENDING) generic path, for pinning against the cv2 binary
void orc_sortidx_desc_stdsort(const float *v, int n, int32_t *out)
{
    std::vector<int> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), LessThanIdx{v});
    for (int j = 0; j < n / 2; j++) std::swap(order[j], order[n - 1 - j]);
    for (int i = 0; i < n; i++) out[i] = order[i];
}

// sequential double accumulation of float terms (the reference's `sum +=` / `change +=`), for the unit test of the
// CUDA emulation
double orc_seqsum(const float *t, long long n)
{
    double s = 0.;
    for (long long i = 0; i < n; i++) s += t[i];
    return s;
}

}  // extern "C"
