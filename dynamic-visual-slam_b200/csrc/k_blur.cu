// k_blur.cu — 7x7 sigma-2 Gaussian smoothing of every pyramid level, integer 8.8 fixed point.
// Replaces `GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101)` on the cloned
// level (reference ORBextractor.cpp:1132-1133).  Arithmetic: SURVEY.md App. A.3 — taps
// [18,34,48,56,48,34,18]/256, row pass in u16, column pass in u32, one rounding (c + 32768) >> 16.
//
// HBM-bound stage, all levels of all frames in ONE launch, no shared memory: a thread owns 4 adjacent
// columns (one aligned 32-bit word) and sweeps down ORBX_BLUR_H output rows.  Per input row it loads the
// aligned words around its columns (coalesced 128-byte warp requests; neighbours hit L1; the loads of 7
// rows are issued back to back for memory-level parallelism), forms the seven-tap windows with funnel
// shifts and evaluates the row pass with two IDP.4A dot products per pixel; the column pass keeps a
// 7-deep register window per column (rows unrolled by 7 so the window rotates statically) and the four
// results leave as one 32-bit store.  Image borders cost nothing in the sweep: the reflect-101 mapping is
// folded into per-thread PRMT selectors and load offsets computed once.
#include "orbx_internal.h"

struct BlurParams {
    const uint8_t *l0; size_t l0_step, l0_fstride;
    const uint8_t *pyr; size_t pyr_slab;
    uint8_t *blur; size_t blur_slab;
};

// reflect-101 with a single fold (valid for -len < p < 2*len - 1; levels are at least 8 px)
__device__ __forceinline__ int reflect1(int p, int len) { p = p < 0 ? -p : p; return p >= len ? 2 * len - 2 - p : p; }

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

// Border handling without branches in the sweep: every thread loads four aligned words per row from
// thread-constant offsets (oL, oC, oRa, oRb) and rebuilds its 12-byte window [x-4, x+8) with three PRMTs
// whose selectors encode the reflect-101 mapping.  Interior threads get the identity mapping.
struct BlurTaps { int oL, oRa, oRb; uint32_t selL, selC, selR; };

__device__ BlurTaps blur_taps(int x, int w)
{
    BlurTaps t;
    // L' = positions x-4..x-1 from PRMT(word@oL, word@x)
    if (x == 0) { t.oL = 4; t.selL = 0x5670u; }             // 4,3,2,1 <- (word@4).b0, (word@0).b3,b2,b1
    else { t.oL = x - 4; t.selL = 0x3210u; }
    // C' = positions x..x+3 from PRMT(word@oL, word@x); reflected sources lie in word@x or word@(x-4)
    uint32_t sc = 0;
    for (int j = 0; j < 4; j++) {
        const int s = reflect1(x + j, w);
        const uint32_t idx = s >= x ? 4u + (uint32_t)(s - x) : (uint32_t)(s - (x - 4));
        sc |= idx << (4 * j);
    }
    t.selC = sc;
    // R' = positions x+4..x+7 from PRMT(word@oRa, word@oRb) with oRb = oRa + 4 (or equal when one word suffices)
    int smin = 1 << 30, smax = -1;
    for (int j = 0; j < 4; j++) { const int s = reflect1(x + 4 + j, w); smin = min(smin, s); smax = max(smax, s); }
    t.oRa = smin & ~3; t.oRb = smax & ~3;
    uint32_t sr = 0;
    for (int j = 0; j < 4; j++) {
        const int s = reflect1(x + 4 + j, w);
        const uint32_t idx = (s & ~3) == t.oRa ? (uint32_t)(s & 3) : 4u + (uint32_t)(s & 3);
        sr |= idx << (4 * j);
    }
    t.selR = sr;
    return t;
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
    const BlurTaps tp = blur_taps(x, w);

    int win[4][7];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const uint8_t *row = src + (size_t)reflect1(ybase - 3 + k, hgt) * step;
        const uint32_t wl = __ldg(reinterpret_cast<const uint32_t *>(row + tp.oL)), wc = __ldg(reinterpret_cast<const uint32_t *>(row + x));
        const uint32_t wa = __ldg(reinterpret_cast<const uint32_t *>(row + tp.oRa)), wb = __ldg(reinterpret_cast<const uint32_t *>(row + tp.oRb));
        hpass4(__byte_perm(wl, wc, tp.selL), __byte_perm(wl, wc, tp.selC), __byte_perm(wa, wb, tp.selR), win[0][k], win[1][k], win[2][k], win[3][k]);
    }
    for (int r0 = 0; r0 < ORBX_BLUR_H; r0 += 7) {
        // issue the loads of the next 7 input rows back to back (memory-level parallelism), then compute
        uint32_t wl[7], wc[7], wa[7], wb[7];
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const uint8_t *row = src + (size_t)reflect1(min(ybase + r0 + k + 3, hgt + 2), hgt) * step;
            wl[k] = __ldg(reinterpret_cast<const uint32_t *>(row + tp.oL)); wc[k] = __ldg(reinterpret_cast<const uint32_t *>(row + x));
            wa[k] = __ldg(reinterpret_cast<const uint32_t *>(row + tp.oRa)); wb[k] = __ldg(reinterpret_cast<const uint32_t *>(row + tp.oRb));
        }
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int y = ybase + r0 + k;
            hpass4(__byte_perm(wl[k], wc[k], tp.selL), __byte_perm(wl[k], wc[k], tp.selC), __byte_perm(wa[k], wb[k], tp.selR),
                   win[0][(k + 6) % 7], win[1][(k + 6) % 7], win[2][(k + 6) % 7], win[3][(k + 6) % 7]);
            uint32_t packed = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t acc = 32768u + 18u * (uint32_t)(win[j][k % 7] + win[j][(k + 6) % 7]) +
                                     34u * (uint32_t)(win[j][(k + 1) % 7] + win[j][(k + 5) % 7]) +
                                     48u * (uint32_t)(win[j][(k + 2) % 7] + win[j][(k + 4) % 7]) + 56u * (uint32_t)win[j][(k + 3) % 7];
                packed |= (acc >> 16) << (8 * j);
            }
            if (y < hgt) *reinterpret_cast<uint32_t *>(dst + (size_t)y * g.bpitch) = packed;   // pitch % 128 == 0: the word is in-bounds
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
