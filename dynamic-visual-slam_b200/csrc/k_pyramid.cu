// k_pyramid.cu — 8-level scale-1.2 image pyramid, fixed-point bilinear (cv::resize INTER_LINEAR, CV_8UC1).
// Replaces ORBextractor::ComputePyramid (reference ORBextractor.cpp:1169-1194; resize call :1182).
// Arithmetic: SURVEY.md App. A.1 — 11-bit coefficients, horizontal pass in int32, vertical pass
//   out = (((b0*(R0>>4))>>16) + ((b1*(R1>>4))>>16) + 2) >> 2.
// The coefficient tables are built on the host exactly as OpenCV builds them (orbx_api.cu) so the
// kernel is pure integer.  The 19-px REFLECT_101 border of the reference is never read downstream
//  and is not materialised.
//
// HBM-bound stage: one CTA produces a 128 x 8 output tile; each thread produces 4 horizontally
// adjacent pixels and stores them as one 32-bit word (rows are 128-byte pitched).
#include "orbx_internal.h"

struct ResizeParams {
    const uint8_t *src; size_t src_step, src_fstride;
    uint8_t *dst; size_t dst_step, dst_fstride;
    int sw, sh, dw, dh;
    const ResizeTab *xtab, *ytab;
};

__global__ void __launch_bounds__(256) k_resize_linear(ResizeParams P)
{
    const int f = blockIdx.z;
    const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x4 >= P.dw || y >= P.dh) return;
    const ResizeTab ty = P.ytab[y];
    int sy0 = ty.ofs, sy1 = ty.ofs + 1;
    sy0 = sy0 < 0 ? 0 : (sy0 >= P.sh ? P.sh - 1 : sy0);
    sy1 = sy1 < 0 ? 0 : (sy1 >= P.sh ? P.sh - 1 : sy1);
    const uint8_t *S0 = P.src + (size_t)f * P.src_fstride + (size_t)sy0 * P.src_step;
    const uint8_t *S1 = P.src + (size_t)f * P.src_fstride + (size_t)sy1 * P.src_step;
    const int b0 = ty.a0, b1 = ty.a1;
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = x4 + i;
        if (x < P.dw) {
            const ResizeTab tx = P.xtab[x];
            const int sx = tx.ofs, sx1 = sx + 1 < P.sw ? sx + 1 : P.sw - 1;
            const int r0 = (int)__ldg(S0 + sx) * tx.a0 + (int)__ldg(S0 + sx1) * tx.a1;
            const int r1 = (int)__ldg(S1 + sx) * tx.a0 + (int)__ldg(S1 + sx1) * tx.a1;
            int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            packed |= (uint32_t)v << (8 * i);
        }
    }
    uint8_t *D = P.dst + (size_t)f * P.dst_fstride + (size_t)y * P.dst_step;
    *reinterpret_cast<uint32_t *>(D + x4) = packed;     // pitch is a multiple of 128: the word is in-bounds
}

void launch_resize_level(orbx_handle *h, int level, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride)
{
    const LevelGeom &gs = h->geo.lv[level - 1], &gd = h->geo.lv[level];
    ResizeParams P;
    if (level == 1) { P.src = l0; P.src_step = l0_step; P.src_fstride = l0_fstride; }
    else { P.src = h->d_pyr + gs.off; P.src_step = (size_t)gs.pitch; P.src_fstride = h->pyr_slab; }
    P.dst = h->d_pyr + gd.off; P.dst_step = (size_t)gd.pitch; P.dst_fstride = h->pyr_slab;
    P.sw = gs.w; P.sh = gs.h; P.dw = gd.w; P.dh = gd.h;
    P.xtab = h->d_xtab + gd.xtab_off; P.ytab = h->d_ytab + gd.ytab_off;
    dim3 block(32, 8), grid((gd.w + 127) / 128, (gd.h + 7) / 8, nframes);
    ProfScope ps(h, ORBX_K_RESIZE);
    k_resize_linear<<<grid, block, 0, h->stream>>>(P);
}
