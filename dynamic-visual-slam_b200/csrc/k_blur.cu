// k_blur.cu — 7x7 sigma-2 Gaussian smoothing of every pyramid level, integer 8.8 fixed point.
// Replaces `GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101)` on the cloned
// level (reference ORBextractor.cpp:1132-1133).  Arithmetic: SURVEY.md App. A.3 — taps
// [18,34,48,56,48,34,18]/256, row pass in u16, column pass in u32, one rounding (c + 32768) >> 16.
//
// HBM-bound stage, all levels of all frames in ONE launch, no shared memory: a thread owns 4 adjacent
// columns (one aligned 32-bit word) and sweeps down BL_H output rows.  Per input row it loads the
// three aligned words around its columns (coalesced 128-byte warp requests; neighbours hit L1),
// forms the seven-tap windows with funnel shifts and evaluates the row pass with two IDP.4A dot
// products per pixel; the column pass keeps a 7-deep register window per column (rows unrolled by 7
// so the window rotates statically) and the four results leave as one 32-bit store.
#include "orbx_internal.h"

struct BlurParams {
    const uint8_t *l0; size_t l0_step, l0_fstride;
    const uint8_t *pyr; size_t pyr_slab;
    uint8_t *blur; size_t blur_slab;
};

__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

// 4 pixels starting at column xs of a row (xs multiple of 4), reflect-101 outside [0, w)
__device__ __forceinline__ uint32_t load_word(const uint8_t *row, int xs, int w)
{
    if (xs >= 0 && xs + 3 < w) return __ldg(reinterpret_cast<const uint32_t *>(row + xs));
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) v |= (uint32_t)__ldg(row + reflect101(xs + b, w)) << (8 * b);
    return v;
}

// row pass for the 4 pixels of word C given its left / right neighbours: 2 x IDP.4A per pixel
__device__ __forceinline__ void hpass4(uint32_t L, uint32_t C, uint32_t R, int &h0, int &h1, int &h2, int &h3)
{
    const uint32_t KLO = 0x38302212u;    // taps 18,34,48,56 on bytes x-3..x
    const uint32_t KHI = 0x00122230u;    // taps 48,34,18 on bytes x+1..x+3
    h0 = __dp4a(__funnelshift_r(L, C, 8), KLO, __dp4a(__funnelshift_r(C, R, 8), KHI, 0u));
    h1 = __dp4a(__funnelshift_r(L, C, 16), KLO, __dp4a(__funnelshift_r(C, R, 16), KHI, 0u));
    h2 = __dp4a(__funnelshift_r(L, C, 24), KLO, __dp4a(__funnelshift_r(C, R, 24), KHI, 0u));
    h3 = __dp4a(C, KLO, __dp4a(R, KHI, 0u));
}

__global__ void __launch_bounds__(128) k_blur7(BlurParams P, const FrameGeom *__restrict__ G)
{
    const int f = blockIdx.y;
    int level = 0;
    const int nl = G->nlevels;
    for (int l = 1; l < nl; l++) if ((int)blockIdx.x >= G->lv[l].blur_first) level = l;
    const LevelGeom &g = G->lv[level];
    const int t = blockIdx.x - g.blur_first;
    const int tx = t % g.blur_tx, ty = t / g.blur_tx;
    const int w = g.w, hgt = g.h;
    const int x = tx * ORBX_BLUR_TW + (threadIdx.x & 31) * 4;
    const int ybase = ty * (4 * ORBX_BLUR_H) + (threadIdx.x >> 5) * ORBX_BLUR_H;
    if (x >= w || ybase >= hgt) return;
    const uint8_t *src; size_t step;
    if (level == 0) { src = P.l0 + (size_t)f * P.l0_fstride; step = P.l0_step; }
    else { src = P.pyr + (size_t)f * P.pyr_slab + g.off; step = (size_t)g.pitch; }
    uint8_t *dst = P.blur + (size_t)f * P.blur_slab + g.boff + x;

    int win[4][7];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const uint8_t *row = src + (size_t)reflect101(ybase - 3 + k, hgt) * step;
        hpass4(load_word(row, x - 4, w), load_word(row, x, w), load_word(row, x + 4, w), win[0][k], win[1][k], win[2][k], win[3][k]);
    }
    for (int r0 = 0; r0 < ORBX_BLUR_H; r0 += 7) {
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int y = ybase + r0 + k;
            if (y < hgt) {
                const uint8_t *row = src + (size_t)reflect101(y + 3, hgt) * step;
                hpass4(load_word(row, x - 4, w), load_word(row, x, w), load_word(row, x + 4, w),
                       win[0][(k + 6) % 7], win[1][(k + 6) % 7], win[2][(k + 6) % 7], win[3][(k + 6) % 7]);
                uint32_t packed = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t acc = 32768u + 18u * (uint32_t)(win[j][k % 7] + win[j][(k + 6) % 7]) +
                                         34u * (uint32_t)(win[j][(k + 1) % 7] + win[j][(k + 5) % 7]) +
                                         48u * (uint32_t)(win[j][(k + 2) % 7] + win[j][(k + 4) % 7]) + 56u * (uint32_t)win[j][(k + 3) % 7];
                    packed |= (acc >> 16) << (8 * j);
                }
                *reinterpret_cast<uint32_t *>(dst + (size_t)y * g.bpitch) = packed;   // pitch % 128 == 0: the word is in-bounds
            }
        }
    }
}

void launch_blur(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride)
{
    BlurParams P;
    P.l0 = l0; P.l0_step = l0_step; P.l0_fstride = l0_fstride;
    P.pyr = h->d_pyr; P.pyr_slab = h->pyr_slab;
    P.blur = h->d_blur; P.blur_slab = h->blur_slab;
    dim3 grid(h->geo.total_blur_tiles, nframes);
    ProfScope ps(h, ORBX_K_BLUR);
    k_blur7<<<grid, 128, 0, h->stream>>>(P, h->d_geo);
}
