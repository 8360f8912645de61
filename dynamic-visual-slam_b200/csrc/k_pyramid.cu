// k_pyramid.cu — 8-level scale-1.2 image pyramid, fixed-point bilinear (cv::resize INTER_LINEAR, CV_8UC1).
// Replaces ORBextractor::ComputePyramid (reference ORBextractor.cpp:1169-1194; resize call :1182).
// Arithmetic: SURVEY.md App. A.1 — 11-bit coefficients, horizontal pass in int32, vertical pass
//   out = (((b0*(R0>>4))>>16) + ((b1*(R1>>4))>>16) + 2) >> 2.
// The coefficient tables are built on the host exactly as OpenCV builds them (orbx_api.cu) so the
// kernel is pure integer.  The 19-px REFLECT_101 border of the reference is never read downstream
//  and is not materialised.
//
// HBM-bound stage.  One CTA produces a 128 x 64 output tile: the source window the tile needs
// (about 1.2x larger per axis) is staged in shared memory with aligned 128-bit loads issued back to
// back (memory-level parallelism), then each thread sweeps 8 output rows for 4 adjacent columns with
// its four horizontal table entries in registers; the horizontal interpolation of a source row is
// reused by the next output row when they share it (5 rows out of 6 at scale 1.2).  Results leave as
// one 32-bit store per thread and row (rows are 128-byte pitched).
#include "orbx_internal.h"

#define RZ_TW 128
#define RZ_TH 64
#define RZ_ROWS_PER_WARP 8

struct ResizeParams {
    const uint8_t *src; size_t src_step, src_fstride;
    uint8_t *dst; size_t dst_step, dst_fstride;
    int sw, sh, dw, dh;
    const ResizeTab *xtab, *ytab;
    int smem_pitch, smem_rows;
};

__global__ void __launch_bounds__(256) k_resize_linear(ResizeParams P)
{
    extern __shared__ __align__(16) uint8_t s_src[];
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * RZ_TW, y0 = blockIdx.y * RZ_TH;
    const int x1 = min(x0 + RZ_TW, P.dw) - 1, y1 = min(y0 + RZ_TH, P.dh) - 1;
    // source window of this tile
    const int sxlo = P.xtab[x0].ofs, sxhi = min(P.xtab[x1].ofs + 1, P.sw - 1);
    const int sylo = max(0, min(P.ytab[y0].ofs, P.sh - 1)), syhi = max(0, min(P.ytab[y1].ofs + 1, P.sh - 1));
    const int abase = sxlo & ~15;
    const int vecs = (sxhi - abase + 16) >> 4;
    const int rows = syhi - sylo + 1;
    const uint8_t *S = P.src + (size_t)f * P.src_fstride + (size_t)sylo * P.src_step + abase;
    const int pitch = P.smem_pitch;
    for (int i = threadIdx.x; i < rows * 16; i += 256) {
        const int r = i >> 4;
        for (int vi = i & 15; vi < vecs; vi += 16)
            reinterpret_cast<uint4 *>(s_src + r * pitch)[vi] = __ldg(reinterpret_cast<const uint4 *>(S + (size_t)r * P.src_step) + vi);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x4 = x0 + lane * 4;
    if (x4 >= P.dw) return;
    // horizontal table entries of the 4 columns (offsets relative to the staged window)
    int o0[4], o1[4], a0[4], a1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const ResizeTab tx = P.xtab[min(x4 + i, P.dw - 1)];
        o0[i] = tx.ofs - abase;
        o1[i] = min(tx.ofs + 1, P.sw - 1) - abase;
        a0[i] = tx.a0; a1[i] = tx.a1;
    }
    int cached_row = -1000000, hc[4] = { 0, 0, 0, 0 };
    uint8_t *D = P.dst + (size_t)f * P.dst_fstride + x4;
    for (int k = 0; k < RZ_ROWS_PER_WARP; k++) {
        const int y = y0 + warp * RZ_ROWS_PER_WARP + k;
        if (y > y1) break;
        const ResizeTab ty = P.ytab[y];
        const int sy0 = max(0, min(ty.ofs, P.sh - 1)), sy1 = max(0, min(ty.ofs + 1, P.sh - 1));
        int h0[4], h1[4];
        if (sy0 == cached_row) {
#pragma unroll
            for (int i = 0; i < 4; i++) h0[i] = hc[i];
        } else {
            const uint8_t *r0 = s_src + (sy0 - sylo) * pitch;
#pragma unroll
            for (int i = 0; i < 4; i++) h0[i] = (int)r0[o0[i]] * a0[i] + (int)r0[o1[i]] * a1[i];
        }
        if (sy1 == sy0) {
#pragma unroll
            for (int i = 0; i < 4; i++) h1[i] = h0[i];
        } else {
            const uint8_t *r1 = s_src + (sy1 - sylo) * pitch;
#pragma unroll
            for (int i = 0; i < 4; i++) h1[i] = (int)r1[o0[i]] * a0[i] + (int)r1[o1[i]] * a1[i];
        }
        cached_row = sy1;
        const int b0 = ty.a0, b1 = ty.a1;
        uint32_t packed = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            hc[i] = h1[i];
            int v = (((b0 * (h0[i] >> 4)) >> 16) + ((b1 * (h1[i] >> 4)) >> 16) + 2) >> 2;
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            packed |= (uint32_t)v << (8 * i);
        }
        *reinterpret_cast<uint32_t *>(D + (size_t)y * P.dst_step) = packed;     // pitch is a multiple of 128: the word is in-bounds
    }
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
    P.smem_pitch = h->geo.rz_pitch; P.smem_rows = h->geo.rz_rows;
    const size_t smem = (size_t)P.smem_pitch * P.smem_rows;
    static size_t configured = 0;
    if (smem > configured) { cudaFuncSetAttribute(k_resize_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
    dim3 grid((gd.w + RZ_TW - 1) / RZ_TW, (gd.h + RZ_TH - 1) / RZ_TH, nframes);
    ProfScope ps(h, ORBX_K_RESIZE);
    k_resize_linear<<<grid, 256, smem, h->stream>>>(P);
}
