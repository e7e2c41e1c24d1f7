// jacobi3.cuh -- cv::eigen() on a symmetric 3x3 CV_32F matrix (OpenCV's JacobiImpl_<float>, as used by
// cv::PCA in util.cpp:299-301), specialised to n = 3 with every array index resolved at compile time so
// that the whole state lives in registers (the generic form indexes A, V, indR, indC dynamically and
// spills them to local memory).
//
// The generic algorithm only touches the upper triangle and, for n = 3, its pivot bookkeeping reduces to
// two indices: indR[0] in {1,2} and indC[2] in {0,1} (indR[1] == 2 and indC[1] == 0 always).  They are
// refreshed only for the rows/columns of the last rotation, i.e. one of them may be STALE -- that is part
// of the reference behaviour (it decides which off-diagonal is zeroed next) and is reproduced here.
// hypot() is OpenCV's own float formula (not libm's), see mr_hypot_f.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MR_HD __host__ __device__ __forceinline__
#else
#define MR_HD static inline
#endif

// OpenCV's own hypot (lapack.cpp, used by JacobiImpl_ instead of libm's): float arithmetic scaled by the larger
// operand.  Pinned bit for bit against cv2.eigen / cv2.PCACompute2 (tests/test_oracle_cv.py).
MR_HD float mr_hypot_f(float a, float b)
{
    a = fabsf(a);
    b = fabsf(b);
    if (a > b) { b = b / a; return a * sqrtf(1.f + b * b); }
    if (b > 0.f) { a = a / b; return b * sqrtf(1.f + a * a); }
    return 0.f;
}

// cov: c00, c01, c02, c11, c12, c22.  W: eigenvalues (descending), V: eigenvectors in rows.
MR_HD void mr_jacobi3(const float *cov6, float *Wout, float *Vout)
{
    const float eps = 1.1920929e-07f;
    float W0 = cov6[0], W1 = cov6[3], W2 = cov6[5];
    float A01 = cov6[1], A02 = cov6[2], A12 = cov6[4];
    float V00 = 1.f, V01 = 0.f, V02 = 0.f, V10 = 0.f, V11 = 1.f, V12 = 0.f, V20 = 0.f, V21 = 0.f, V22 = 1.f;
    int indR0 = (fabsf(A01) < fabsf(A02)) ? 2 : 1;
    int indC2 = (fabsf(A02) < fabsf(A12)) ? 1 : 0;
    for (int iters = 0; iters < 270; iters++) {
        // pivot search (strict '<' everywhere: first maximal candidate wins)
        int k = 0, l = indR0;
        float mv = fabsf(indR0 == 1 ? A01 : A02);
        float val = fabsf(A12);
        if (mv < val) { mv = val; k = 1; l = 2; }
        val = fabsf(A01);
        if (mv < val) { mv = val; k = 0; l = 1; }
        val = fabsf(indC2 == 0 ? A02 : A12);
        if (mv < val) { mv = val; k = indC2; l = 2; }
        const int pair = (k == 0) ? (l == 1 ? 0 : 1) : 2;      // 0:(0,1) 1:(0,2) 2:(1,2)
        const float p = pair == 0 ? A01 : (pair == 1 ? A02 : A12);
        if (fabsf(p) <= eps) break;
        const float Wk = (pair == 2) ? W1 : W0, Wl = (pair == 0) ? W1 : W2;
        float y = (float)((Wl - Wk) * 0.5);
        float t = fabsf(y) + mr_hypot_f(p, y);
        float s = mr_hypot_f(p, t);
        float c = t / s;
        s = p / s;
        t = (p / t) * p;
        if (y < 0) { s = -s; t = -t; }
        float a0, b0;
#define MR_ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
        if (pair == 0) {
            A01 = 0.f; W0 -= t; W1 += t;
            MR_ROT(A02, A12);
            MR_ROT(V00, V10); MR_ROT(V01, V11); MR_ROT(V02, V12);
            indR0 = (fabsf(A01) < fabsf(A02)) ? 2 : 1;          // idx 0: indR refreshed; idx 1: nothing dynamic
        } else if (pair == 1) {
            A02 = 0.f; W0 -= t; W2 += t;
            MR_ROT(A01, A12);
            MR_ROT(V00, V20); MR_ROT(V01, V21); MR_ROT(V02, V22);
            indR0 = (fabsf(A01) < fabsf(A02)) ? 2 : 1;
            indC2 = (fabsf(A02) < fabsf(A12)) ? 1 : 0;
        } else {
            A12 = 0.f; W1 -= t; W2 += t;
            MR_ROT(A01, A02);
            MR_ROT(V10, V20); MR_ROT(V11, V21); MR_ROT(V12, V22);
            indC2 = (fabsf(A02) < fabsf(A12)) ? 1 : 0;          // idx 2: indC refreshed; indR[0] stays stale
        }
#undef MR_ROT
    }
    // selection sort, descending (ties keep order), swapping eigenvector rows along
#define MR_SWAP_ROWS(Wa, Wb, a0_, a1_, a2_, b0_, b1_, b2_) { float tt; tt = Wa; Wa = Wb; Wb = tt; tt = a0_; a0_ = b0_; b0_ = tt; tt = a1_; a1_ = b1_; b1_ = tt; tt = a2_; a2_ = b2_; b2_ = tt; }
    {
        // k = 0: m = argmax over {0,1,2} with strict '<'
        int m = 0;
        if (W0 < W1) m = 1;
        if ((m == 0 ? W0 : W1) < W2) m = 2;
        if (m == 1) MR_SWAP_ROWS(W1, W0, V10, V11, V12, V00, V01, V02)
        else if (m == 2) MR_SWAP_ROWS(W2, W0, V20, V21, V22, V00, V01, V02)
        // k = 1
        if (W1 < W2) MR_SWAP_ROWS(W2, W1, V20, V21, V22, V10, V11, V12)
    }
#undef MR_SWAP_ROWS
    Wout[0] = W0; Wout[1] = W1; Wout[2] = W2;
    Vout[0] = V00; Vout[1] = V01; Vout[2] = V02; Vout[3] = V10; Vout[4] = V11; Vout[5] = V12; Vout[6] = V20; Vout[7] = V21; Vout[8] = V22;
}
