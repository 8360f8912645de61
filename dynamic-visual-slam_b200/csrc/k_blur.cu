// k_blur.cu — 7x7 sigma-2 Gaussian smoothing of every pyramid level, integer 8.8 fixed point.
// Replaces `GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101)` on the cloned
// level (reference ORBextractor.cpp:1132-1133).  Arithmetic: SURVEY.md App. A.3 — taps
// [18,34,48,56,48,34,18]/256, row pass in u16, column pass in u32, one rounding (c + 32768) >> 16.
//
// HBM-bound stage: all levels of all frames in ONE launch; a CTA owns a 128 x 32 output tile,
// stages the (128+8) x (32+6) input window in shared memory with 32-bit loads, runs the row pass
// into a u16 shared tile and the column pass straight to 32-bit global stores.
#include "orbx_internal.h"

#define BT_W 128
#define BT_H 32
#define BIN_W (BT_W + 8)     // 4-byte aligned window: [x0-4, x0+132)
#define BIN_H (BT_H + 6)

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

__global__ void __launch_bounds__(256) k_blur7(BlurParams P, const FrameGeom *__restrict__ G)
{
    __shared__ __align__(16) uint8_t s_in[BIN_H][BIN_W];
    __shared__ __align__(16) uint16_t s_row[BIN_H][BT_W];
    const int f = blockIdx.y;
    int level = 0;
    const int nl = G->nlevels;
    for (int l = 1; l < nl; l++) if ((int)blockIdx.x >= G->lv[l].blur_first) level = l;
    const LevelGeom &g = G->lv[level];
    const int t = blockIdx.x - g.blur_first;
    const int x0 = (t % g.blur_tx) * BT_W, y0 = (t / g.blur_tx) * BT_H;
    const int w = g.w, hgt = g.h;
    const uint8_t *src; size_t step;
    if (level == 0) { src = P.l0 + (size_t)f * P.l0_fstride; step = P.l0_step; }
    else { src = P.pyr + (size_t)f * P.pyr_slab + g.off; step = (size_t)g.pitch; }

    // stage input window (reflect-101 on the level itself)
    for (int i = threadIdx.x; i < BIN_H * (BIN_W / 4); i += 256) {
        const int r = i / (BIN_W / 4), cw = i % (BIN_W / 4);
        const int sy = reflect101(y0 - 3 + r, hgt);
        const int sx = x0 - 4 + cw * 4;
        uint32_t v;
        if (sx >= 0 && sx + 3 < w) v = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)sy * step + sx));
        else {
            v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) v |= (uint32_t)__ldg(src + (size_t)sy * step + reflect101(sx + b, w)) << (8 * b);
        }
        *reinterpret_cast<uint32_t *>(&s_in[r][cw * 4]) = v;
    }
    __syncthreads();
    // row pass: s_row[r][c] = sum K_i * in[r][c + i + 1]   (window column c+4 is pixel x0+c)
    for (int i = threadIdx.x; i < BIN_H * BT_W; i += 256) {
        const int r = i / BT_W, c = i % BT_W;
        const uint8_t *p = &s_in[r][c + 1];
        const int acc = 18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3];
        s_row[r][c] = (uint16_t)acc;
    }
    __syncthreads();
    // column pass, 4 pixels per thread-item
    uint8_t *dst = P.blur + (size_t)f * P.blur_slab + g.boff;
    for (int i = threadIdx.x; i < BT_H * (BT_W / 4); i += 256) {
        const int r = i / (BT_W / 4), c4 = (i % (BT_W / 4)) * 4;
        const int y = y0 + r, x = x0 + c4;
        if (y >= hgt || x >= w) continue;
        uint32_t packed = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int c = c4 + b;
            const uint32_t acc = 18u * (s_row[r][c] + s_row[r + 6][c]) + 34u * (s_row[r + 1][c] + s_row[r + 5][c]) +
                                 48u * (s_row[r + 2][c] + s_row[r + 4][c]) + 56u * s_row[r + 3][c];
            packed |= ((acc + 32768u) >> 16) << (8 * b);
        }
        *reinterpret_cast<uint32_t *>(dst + (size_t)y * g.bpitch + x) = packed;   // pitch % 128 == 0: in-bounds
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
    k_blur7<<<grid, 256, 0, h->stream>>>(P, h->d_geo);
}
