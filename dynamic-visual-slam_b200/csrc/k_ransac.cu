// k_ransac.cu — geometric validation, the data-parallel half: scoring fundamental-matrix hypotheses against all correspondences.
// Reference: Frontend::syncCallback, frontend.cpp:1134-1154 and :625-645 —
//     cv::findFundamentalMat(prev_pts, curr_pts, inliers_mask, cv::FM_RANSAC, 2.0, 0.99)
// OpenCV's RANSAC (calib3d, RANSACPointSetRegistrator + FMEstimatorCallback::computeError) scores every minimal-sample model F over ALL
// point pairs with the symmetric epipolar distance, in double precision:
//     a = F0 x1 + F1 y1 + F2, b = F3 x1 + F4 y1 + F5, c = F6 x1 + F7 y1 + F8;  s2 = 1 / (a a + b b);  d2 = x2 a + y2 b + c
//     a = F0 x2 + F3 y2 + F6, b = F1 x2 + F4 y2 + F7, c = F2 x2 + F5 y2 + F8;  s1 = 1 / (a a + b b);  d1 = x1 a + y1 b + c
//     err = (float) max(d1 d1 s1, d2 d2 s2);      inlier  <=>  err <= (float)(threshold * threshold)
// and keeps the model with the most inliers.  That K x N evaluation (K up to 1000 iterations, N up to ~1000 matches) is what runs here:
// one CTA per hypothesis, every operation rounded on its own (no FMA contraction) so that the inlier mask of a given F equals OpenCV's
// bit for bit; a second launch picks the hypothesis with the most inliers (ties: lowest index, as the sequential loop keeps the first).
// Sampling and the 7/8-point solves stay with the caller ("RANSAC is explicitly out of scope for now").
#include "orbx_internal.h"

struct FmatParams {
    const float *p1, *p2; int n;
    const double *F; int nh;
    float t2;
    int32_t *counts; uint8_t *masks;          // masks: [nh][n]
};

__global__ void __launch_bounds__(256) k_fmat_score(FmatParams P)
{
    __shared__ int s_cnt;
    const int hi = blockIdx.x;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    double F[9];
#pragma unroll
    for (int i = 0; i < 9; i++) F[i] = P.F[(size_t)hi * 9 + i];
    int local = 0;
    for (int i = threadIdx.x; i < P.n; i += blockDim.x) {
        const double x1 = (double)P.p1[2 * i], y1 = (double)P.p1[2 * i + 1], x2 = (double)P.p2[2 * i], y2 = (double)P.p2[2 * i + 1];
        double a = __dadd_rn(__dadd_rn(__dmul_rn(F[0], x1), __dmul_rn(F[1], y1)), F[2]);
        double b = __dadd_rn(__dadd_rn(__dmul_rn(F[3], x1), __dmul_rn(F[4], y1)), F[5]);
        double c = __dadd_rn(__dadd_rn(__dmul_rn(F[6], x1), __dmul_rn(F[7], y1)), F[8]);
        const double s2 = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(x2, a), __dmul_rn(y2, b)), c);
        a = __dadd_rn(__dadd_rn(__dmul_rn(F[0], x2), __dmul_rn(F[3], y2)), F[6]);
        b = __dadd_rn(__dadd_rn(__dmul_rn(F[1], x2), __dmul_rn(F[4], y2)), F[7]);
        c = __dadd_rn(__dadd_rn(__dmul_rn(F[2], x2), __dmul_rn(F[5], y2)), F[8]);
        const double s1 = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
        const double d1 = __dadd_rn(__dadd_rn(__dmul_rn(x1, a), __dmul_rn(y1, b)), c);
        const float err = (float)fmax(__dmul_rn(__dmul_rn(d1, d1), s1), __dmul_rn(__dmul_rn(d2, d2), s2));
        const bool in = err <= P.t2;
        P.masks[(size_t)hi * P.n + i] = in ? 1 : 0;
        local += in;
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) P.counts[hi] = s_cnt;
}

__global__ void __launch_bounds__(256) k_fmat_best(const int32_t *counts, int nh, const uint8_t *masks, int n, int32_t *best, uint8_t *best_mask)
{
    __shared__ unsigned long long s_key;
    if (threadIdx.x == 0) s_key = 0ull;
    __syncthreads();
    unsigned long long k = 0ull;                                // (count, nh - 1 - index): the maximum is the largest count at the lowest index
    for (int i = threadIdx.x; i < nh; i += blockDim.x) {
        const unsigned long long c = ((unsigned long long)(unsigned)counts[i] << 32) | (unsigned)(nh - 1 - i);
        k = c > k ? c : k;
    }
    atomicMax(&s_key, k);
    __syncthreads();
    const int b = nh - 1 - (int)(s_key & 0xFFFFFFFFull);
    if (threadIdx.x == 0) *best = b;
    if (best_mask) for (int i = threadIdx.x; i < n; i += blockDim.x) best_mask[i] = masks[(size_t)b * n + i];
}

void launch_fmat_score(orbx_handle *h, const float *d_p1, const float *d_p2, int n, const double *d_F, int nh, float t2,
                       int32_t *d_counts, uint8_t *d_masks, int32_t *d_best, uint8_t *d_best_mask)
{
    FmatParams P;
    P.p1 = d_p1; P.p2 = d_p2; P.n = n; P.F = d_F; P.nh = nh; P.t2 = t2; P.counts = d_counts; P.masks = d_masks;
    { ProfScope ps(h, ORBX_K_OTHER); k_fmat_score<<<nh, 256, 0, h->stream>>>(P); }
    { ProfScope ps(h, ORBX_K_OTHER); k_fmat_best<<<1, 256, 0, h->stream>>>(d_counts, nh, d_masks, n, d_best, d_best_mask); }
}
