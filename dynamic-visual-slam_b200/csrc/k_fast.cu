// k_fast.cu — FAST-9/16 with per-cell 3x3 NMS and the two-threshold fallback.
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (reference ORBextractor.cpp:785-872):
// for every ~35-px cell (+6 px overlap) `cv::FAST(roi, kps, iniThFAST, true)` and, iff that returned
// nothing, `cv::FAST(roi, kps, minThFAST, true)` (:826-827, :845-846).  FAST arithmetic: SURVEY App. A.2
//   S = max over the 16 arcs of 9 contiguous ring pixels of min(I(p)-I(ring)) resp. min(I(ring)-I(p));
//   corner iff S > th; score = S - 1; strict 3x3 NMS INSIDE the cell's ROI (scores outside the
//   ROI's 3-px margin read as 0).
//
// One CTA per cell, all levels of all frames in one launch.  The cell's ROI is staged in shared
// memory (pitch 80 so ring offsets are compile-time constants), scores go to a second shared tile,
// survivors are appended to the (frame, level) candidate list with one global atomic per CTA.
// Candidate order in the list is arbitrary; the quadtree kernel is order-independent (it
// reconstructs the reference's insertion order from coordinates where ties need it).
#include "orbx_internal.h"

#define FT_PITCH 80            // >= max ROI width 75 (wCell <= 69 for any box >= 35 px)
#define FT_ROWS 76
#define FS_PITCH 72            // score tile: detection area + 1-px zero ring, <= 71 wide
#define FS_ROWS 72
#define FAST_THREADS 128
#define FAST_OUT_CAP 1296      // NMS survivors are never 8-adjacent: <= ceil(69/2)^2 = 1225

struct FastParams {
    const uint8_t *l0; size_t l0_step, l0_fstride;
    const uint8_t *pyr; size_t pyr_slab;
    uint32_t *cand; size_t cand_slab;
    int32_t *ncand;
    int ini_th, min_th;
    int32_t *status;
};

// ring offsets in a pitch-80 tile, order of SURVEY App. A.2
#define RO(dx, dy) ((dy) * FT_PITCH + (dx))
__device__ __forceinline__ int fast_score(const uint8_t *p)
{
    // S = max( I(p) - min_k max9_k(ring),  max_k min9_k(ring) - I(p) ) over the 16 arcs of 9 contiguous ring
    // pixels: "all darker by more than t" <=> I(p) - max9 > t ; "all brighter" <=> min9 - I(p) > t.
    // Formulated on the raw ring values (no negated operands) with a log-step sliding window.
    const int v = p[0];
    int r[16];
    r[0] = p[RO(0, 3)];   r[1] = p[RO(1, 3)];    r[2] = p[RO(2, 2)];    r[3] = p[RO(3, 1)];
    r[4] = p[RO(3, 0)];   r[5] = p[RO(3, -1)];   r[6] = p[RO(2, -2)];   r[7] = p[RO(1, -3)];
    r[8] = p[RO(0, -3)];  r[9] = p[RO(-1, -3)];  r[10] = p[RO(-2, -2)]; r[11] = p[RO(-3, -1)];
    r[12] = p[RO(-3, 0)]; r[13] = p[RO(-3, 1)];  r[14] = p[RO(-2, 2)];  r[15] = p[RO(-1, 3)];
    int mn2[16], mx2[16], mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { mn2[k] = min(r[k], r[(k + 1) & 15]); mx2[k] = max(r[k], r[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; k++) { mn4[k] = min(mn2[k], mn2[(k + 2) & 15]); mx4[k] = max(mx2[k], mx2[(k + 2) & 15]); }
    int hi_of_min = 0, lo_of_max = 255;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), r[(k + 8) & 15]);
        const int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), r[(k + 8) & 15]);
        hi_of_min = max(hi_of_min, mn9);
        lo_of_max = min(lo_of_max, mx9);
    }
    const int s_dark = v - lo_of_max, s_bright = hi_of_min - v;
    return s_dark > s_bright ? s_dark : s_bright;
}

__global__ void __launch_bounds__(FAST_THREADS) k_fast_cells(FastParams P, const FrameGeom *__restrict__ G)
{
    __shared__ __align__(16) uint8_t s_img[FT_ROWS * FT_PITCH];
    __shared__ uint8_t s_sc[FS_ROWS * FS_PITCH];
    __shared__ uint32_t s_out[FAST_OUT_CAP];
    __shared__ int s_nout, s_base;

    const int f = blockIdx.y;
    int level = 0;
    const int nl = G->nlevels;
    for (int l = 1; l < nl; l++) if ((int)blockIdx.x >= G->lv[l].cell_first) level = l;
    const LevelGeom &g = G->lv[level];
    const int cell = blockIdx.x - g.cell_first;
    const int ci = cell / g.ncols, cj = cell % g.ncols;
    // cell ROI in image coordinates — ORBextractor.cpp:805-822
    const int maxBX = g.w - ORBX_BORDER, maxBY = g.h - ORBX_BORDER;
    const int iniX = ORBX_BORDER + cj * g.wcell, iniY = ORBX_BORDER + ci * g.hcell;
    if (iniY >= maxBY - 3 || iniX >= maxBX - 6) return;
    const int maxX = min(iniX + g.wcell + 6, maxBX), maxY = min(iniY + g.hcell + 6, maxBY);
    const int rw = maxX - iniX, rh = maxY - iniY;
    const int dw = rw - 6, dh = rh - 6;            // detection area, ROI-relative origin (3,3)
    if (dw <= 0 || dh <= 0) return;

    const uint8_t *src; size_t step;
    if (level == 0) { src = P.l0 + (size_t)f * P.l0_fstride; step = P.l0_step; }
    else { src = P.pyr + (size_t)f * P.pyr_slab + g.off; step = (size_t)g.pitch; }
    src += (size_t)iniY * step + iniX;

    for (int i = threadIdx.x; i < rh * rw; i += FAST_THREADS) {
        const int r = i / rw, c = i - r * rw;
        s_img[r * FT_PITCH + c] = __ldg(src + (size_t)r * step + c);
    }
    // zero the score tile once; the 1-px ring around the detection area stays 0 for both passes
    for (int i = threadIdx.x; i < (dh + 2) * FS_PITCH; i += FAST_THREADS) s_sc[i] = 0;
    if (threadIdx.x == 0) s_nout = 0;
    __syncthreads();

    for (int pass = 0; pass < 2; pass++) {
        const int th = pass == 0 ? P.ini_th : P.min_th;
        // scores
        for (int i = threadIdx.x; i < dw * dh; i += FAST_THREADS) {
            const int r = i / dw, c = i - r * dw;
            const uint8_t *p = &s_img[(r + 3) * FT_PITCH + (c + 3)];
            const int v = p[0], hi = v + th, lo = v - th;
            int sc = 0;
            // exact necessary condition: each opposite ring pair must contain an arc member
            int a = p[RO(0, 3)], b = p[RO(0, -3)];
            bool br = (a > hi) | (b > hi), dk = (a < lo) | (b < lo);
            if (br | dk) {
                a = p[RO(3, 0)]; b = p[RO(-3, 0)];
                br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
                if (br | dk) {
                    a = p[RO(2, 2)]; b = p[RO(-2, -2)];
                    br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
                    a = p[RO(2, -2)]; b = p[RO(-2, 2)];
                    br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
                    if (br | dk) {
                        const int s = fast_score(p);
                        if (s > th) sc = s - 1;
                    }
                }
            }
            s_sc[(r + 1) * FS_PITCH + (c + 1)] = (uint8_t)sc;
        }
        __syncthreads();
        // strict 3x3 NMS inside the cell
        for (int i = threadIdx.x; i < dw * dh; i += FAST_THREADS) {
            const int r = i / dw, c = i - r * dw;
            const uint8_t *q = &s_sc[(r + 1) * FS_PITCH + (c + 1)];
            const int s = q[0];
            if (s == 0) continue;
            if (s > q[-1] && s > q[1] && s > q[-FS_PITCH - 1] && s > q[-FS_PITCH] && s > q[-FS_PITCH + 1] &&
                s > q[FS_PITCH - 1] && s > q[FS_PITCH] && s > q[FS_PITCH + 1]) {
                const int o = atomicAdd(&s_nout, 1);
                // box-relative coordinates: kp.pt + (j*wCell, i*hCell) — ORBextractor.cpp:865-866
                if (o < FAST_OUT_CAP) s_out[o] = orbx_pack(cj * g.wcell + c + 3, ci * g.hcell + r + 3, s);
            }
        }
        __syncthreads();
        if (s_nout > 0) break;            // uniform: fallback only when the first pass found nothing
    }
    const int n = min(s_nout, FAST_OUT_CAP);
    if (n == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(&P.ncand[f * nl + level], n);
    __syncthreads();
    const int base = s_base;
    uint32_t *dst = P.cand + (size_t)f * P.cand_slab + g.cand_off;
    for (int i = threadIdx.x; i < n; i += FAST_THREADS) {
        if (base + i < g.cand_cap) dst[base + i] = s_out[i];
    }
    if (threadIdx.x == 0 && (base + n > g.cand_cap || s_nout > FAST_OUT_CAP)) atomicOr(P.status, ORBX_DS_CAND_OVERFLOW);
}

void launch_fast(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride)
{
    FastParams P;
    P.l0 = l0; P.l0_step = l0_step; P.l0_fstride = l0_fstride;
    P.pyr = h->d_pyr; P.pyr_slab = h->pyr_slab;
    P.cand = h->d_cand; P.cand_slab = h->geo.cand_entries;
    P.ncand = h->d_ncand;
    P.ini_th = h->prm.ini_th_fast; P.min_th = h->prm.min_th_fast;
    P.status = h->d_status;
    dim3 grid(h->geo.total_cells, nframes);
    ProfScope ps(h, ORBX_K_FAST);
    k_fast_cells<<<grid, FAST_THREADS, 0, h->stream>>>(P, h->d_geo);
}
