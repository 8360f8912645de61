// k_pyramid.cu — ORBextractor::ComputePyramid (reference ORBextractor.cpp:1169-1194).
// Level l = cv::resize(level l-1, INTER_LINEAR), chained level to level.  Arithmetic restated from
// OpenCV's 8-bit fixed-point bilinear path: 11-bit horizontal/vertical
// coefficients (host-built tables, orbx_api.cu), out = (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2;
// the kernel is pure integer.  The 19-px REFLECT_101 border of the reference is never read downstream
//  and is not materialised.
//
// HBM-bound stage, one launch per level.  A CTA produces an rz_tw x rz_th output tile (192 x 64 at scale 1.2): the
// source window it needs (about 1.2x larger per axis, 16-byte aligned start) arrives by ONE TMA load into shared
// memory, the tile's vertical table (source row offsets + coefficients, clamped) is staged beside it, and each
// thread owns 4 adjacent output columns — their horizontal table entries live in registers — and walks down a
// quarter of the tile's rows.  The horizontal interpolation of a source row is reused by the next output row
// when they share it (5 rows out of 6 at scale 1.2).  Results leave as one 32-bit coalesced store per thread and
// row (rows are 128-byte pitched).
// The whole chain runs as ONE cooperative launch (k_pyramid_all): persistent CTAs walk the tiles of level 1, a grid-wide barrier
// (plus a generic->async proxy fence, because the next level is read by TMA) separates the levels — six barriers instead of six
// kernel boundaries, which is what the single-frame latency pays for.  k_resize_linear (one level per launch) is the fallback.
#include "orbx_internal.h"
#include "orbx_tma.h"
#include <algorithm>
#include <cstring>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define RZ_GROUPS 48                 // 4-pixel column groups per tile row (rz_tw <= 192)
#define RZ_BANDS 4
#define RZ_THREADS (RZ_GROUPS * RZ_BANDS)
#define RZ_MAX_TH 64

struct ResizeParams {
    uint8_t *dst; size_t dst_step, dst_fstride;
    int sw, sh, dw, dh;
    const ResizeTab *xtab, *ytab;
    int tw, th;                      // output tile
    int src_level;
};

// vertical pass of four pixels: out = (((b0 * t0) >> 16) + ((b1 * t1) >> 16) + 2) >> 2, packed into one word.
// The high halves of the eight products are gathered two per register (PRMT), so the additions, the rounding constant and the final
// shift run on two pixels at once.  No clamp: a0 + a1 and b0 + b1 are 2048 +- 1, so t <= 32655 and every sum is at most 1021 + 2
// (no carry between the halves, result <= 255).
__device__ __forceinline__ uint32_t rz_vpack(int b0, int b1, const int (&t0)[4], const int (&t1)[4])
{
    uint32_t u[4], v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { u[i] = (uint32_t)(b0 * t0[i]); v[i] = (uint32_t)(b1 * t1[i]); }
    const uint32_t s01 = __byte_perm(u[0], u[1], 0x7632) + __byte_perm(v[0], v[1], 0x7632) + 0x00020002u;
    const uint32_t s23 = __byte_perm(u[2], u[3], 0x7632) + __byte_perm(v[2], v[3], 0x7632) + 0x00020002u;
    return __byte_perm(s01 >> 2, s23 >> 2, 0x6420);
}

// horizontal pass of four adjacent output columns on one source row: (S[o] * a0 + S[o + 1] * a1) >> 4 each.
// The four byte pairs lie within 8 bytes of the first one (scale <= 2): three aligned words, two funnel shifts that bring byte o[0] to
// the front, then per column one PRMT (its pair into the low half) and one IDP.2A against the packed coefficients — 3 LDS + 10 ALU
// instead of 8 byte loads + 8 IMAD.  Steeper scales take the byte-wise path.
struct RzCols { int o[4], a0[4], a1[4]; int wofs, sh; uint32_t sel[4], cf[4]; };
__device__ __forceinline__ void rz_cols_finish(RzCols &c)
{
    c.wofs = c.o[0] & ~3; c.sh = 8 * (c.o[0] & 3);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int d = min(c.o[i] - c.o[0], 6);
        c.sel[i] = (uint32_t)d | ((uint32_t)(d + 1) << 4);
        c.cf[i] = (uint32_t)(c.a0[i] & 0xFFFF) | ((uint32_t)c.a1[i] << 16);
    }
}
__device__ __forceinline__ void rz_hrow(const uint8_t *q, const RzCols &c, bool bytewise, int (&t)[4])
{
    if (bytewise) {
#pragma unroll
        for (int i = 0; i < 4; i++) t[i] = ((int)q[c.o[i]] * c.a0[i] + (int)q[c.o[i] + 1] * c.a1[i]) >> 4;
        return;
    }
    const uint32_t *w = reinterpret_cast<const uint32_t *>(q + c.wofs);
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    const uint32_t A = __funnelshift_r(w0, w1, c.sh), B = __funnelshift_r(w1, w2, c.sh);
#pragma unroll
    for (int i = 0; i < 4; i++) t[i] = (int)(__dp2a_lo(c.cf[i], __byte_perm(A, B, c.sel[i]), 0u) >> 4);
}

__global__ void __launch_bounds__(RZ_THREADS) k_resize_linear(const __grid_constant__ LevelMaps M, ResizeParams P)
{
    extern __shared__ __align__(128) uint8_t s_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int4 s_rt[RZ_MAX_TH];                                              // per output row: smem offsets of its two source rows, b0, b1
    uint8_t *s_img = s_raw + ((128u - (smem_u32(s_raw) & 127u)) & 127u);          // TMA destination: 128-byte aligned
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * P.tw, y0 = blockIdx.y * P.th;
    const int x1 = min(x0 + P.tw, P.dw) - 1, y1 = min(y0 + P.th, P.dh) - 1;
    // source window of this tile: columns from a 16-byte boundary, rows from the first source row
    const int abase = __ldg(&P.xtab[x0].ofs) & ~15;
    const int sylo = max(0, min(__ldg(&P.ytab[y0].ofs), P.sh - 1));
    // Programmatic dependent launch: the levels are a chain of launches of this kernel.  Every CTA releases the next level's launch at
    // once (its CTAs take the slots this level's last wave frees and run their table prologue early); only the thread that issues the
    // TMA load of the source window — the one access to the previous level — waits for the previous launch to be complete and visible.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&s_bar, (uint32_t)(ORBX_RZ_BOX_ROWS * ORBX_TMA_BOX_BYTES));
        asm volatile("griddepcontrol.wait;" ::: "memory");
        tma_load_3d(s_img, &M.m[P.src_level], abase >> 2, sylo, f, &s_bar);
    }
    for (int r = threadIdx.x; r <= y1 - y0; r += RZ_THREADS) {
        const ResizeTab ty = P.ytab[y0 + r];
        const int sy0 = max(0, min(ty.ofs, P.sh - 1)), sy1 = max(0, min(ty.ofs + 1, P.sh - 1));
        s_rt[r] = make_int4((sy0 - sylo) * ORBX_TMA_BOX_BYTES, (sy1 - sylo) * ORBX_TMA_BOX_BYTES, ty.a0, ty.a1);
    }
    const int grp = threadIdx.x % RZ_GROUPS, band = threadIdx.x / RZ_GROUPS;
    const int x4 = x0 + 4 * grp;
    // horizontal table entries of the 4 columns: byte offsets inside the staged window (the right neighbour of the last
    // source column has coefficient 0, App. A.1) and the two coefficients
    RzCols C;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const ResizeTab tx = P.xtab[min(x4 + i, P.dw - 1)];
        C.o[i] = tx.ofs - abase; C.a0[i] = tx.a0; C.a1[i] = tx.a1;
    }
    rz_cols_finish(C);
    const bool bytewise = P.sw > 2 * P.dw;
    __syncthreads();                                                                 // barrier initialised, row table staged
    mbar_wait(&s_bar, 0);
    if (x4 > x1) return;
    const int nrows = y1 - y0 + 1, RB = (nrows + RZ_BANDS - 1) / RZ_BANDS;
    const int rb = band * RB, re = min(nrows, rb + RB);
    uint8_t *D = P.dst + (size_t)f * P.dst_fstride + (size_t)(y0 + rb) * P.dst_step + x4;
    int prev_off = -1, tp[4] = { 0, 0, 0, 0 };                                       // (h >> 4) of the last source row interpolated
    for (int r = rb; r < re; r++, D += P.dst_step) {
        const int4 e = s_rt[r];
        int t0[4], t1[4];
        if (e.x == prev_off) {
#pragma unroll
            for (int i = 0; i < 4; i++) t0[i] = tp[i];
        } else {
            rz_hrow(s_img + e.x, C, bytewise, t0);
        }
        if (e.y == e.x) {
#pragma unroll
            for (int i = 0; i < 4; i++) t1[i] = t0[i];
        } else {
            rz_hrow(s_img + e.y, C, bytewise, t1);
        }
        prev_off = e.y;
#pragma unroll
        for (int i = 0; i < 4; i++) tp[i] = t1[i];
        *reinterpret_cast<uint32_t *>(D) = rz_vpack(e.z, e.w, t0, t1);              // pitch % 128 == 0: the word is in-bounds
    }
}

int launch_resize_level(orbx_handle *h, int level, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride);

// ---- batch variant: warp = 32 column groups x 16 consecutive output rows, every branch warp-uniform, no register shuffling ----
// Same arithmetic and the same helpers as k_resize_linear; what changes is who does what.  There a 192-thread CTA maps 48 column groups
// x 4 row bands onto 6 warps, so two warps straddle two bands and diverge on every "can the previous row's horizontal pass be reused?"
// branch, and the reuse itself costs register moves; the profile showed 86 warp instructions per (warp, output row) against ~45 of
// arithmetic.  Here a CTA is 4 warps over a 128 x 64 output tile: a warp owns 128 columns x 16 rows, its reuse decisions are the same for
// all lanes, and the row loop is unrolled by two with the roles of the two row registers swapped, so "this row's lower source row is the
// next row's upper one" (5 rows out of 6 at scale 1.2) needs no copy at all.  Row codes are prepared once per tile in shared memory:
// code 0 = upper source row is the previous output row's lower one; anything else = interpolate both rows (always correct).
#define RZ2_TW 128
#define RZ2_TH 64
#define RZ2_WARPS 4
#define RZ2_ROWS (RZ2_TH / RZ2_WARPS)
template <bool bytewise> __global__ void __launch_bounds__(RZ2_WARPS * 32) k_resize_linear2(const __grid_constant__ LevelMaps M, ResizeParams P)
{
    extern __shared__ __align__(128) uint8_t s_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int4 s_rt[RZ2_TH];                                                 // per output row: smem offsets of its two source rows | code << 24, b0, b1
    uint8_t *s_img = s_raw + ((128u - (smem_u32(s_raw) & 127u)) & 127u);
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * RZ2_TW, y0 = blockIdx.y * RZ2_TH;
    const int x1 = min(x0 + RZ2_TW, P.dw) - 1, y1 = min(y0 + RZ2_TH, P.dh) - 1;
    const int abase = __ldg(&P.xtab[x0].ofs) & ~15;
    const int sylo = max(0, min(__ldg(&P.ytab[y0].ofs), P.sh - 1));
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&s_bar, (uint32_t)(ORBX_RZ_BOX_ROWS * ORBX_RZ2_BOX_BYTES));
        asm volatile("griddepcontrol.wait;" ::: "memory");
        tma_load_3d(s_img, &M.m[P.src_level], abase >> 2, sylo, f, &s_bar);
    }
    for (int r = threadIdx.x; r <= y1 - y0; r += RZ2_WARPS * 32) {
        const ResizeTab ty = P.ytab[y0 + r];
        const int sy0 = max(0, min(ty.ofs, P.sh - 1)), sy1 = max(0, min(ty.ofs + 1, P.sh - 1));
        int code = 1;
        if (r % RZ2_ROWS != 0) {                                                    // not the first row of a warp's band: compare with the previous output row
            const ResizeTab tp = P.ytab[y0 + r - 1];
            const int py1 = max(0, min(tp.ofs + 1, P.sh - 1));
            if (sy0 == py1 && sy1 != sy0) code = 0;
        }
        s_rt[r] = make_int4((sy0 - sylo) * ORBX_RZ2_BOX_BYTES, ((sy1 - sylo) * ORBX_RZ2_BOX_BYTES) | (code << 24), ty.a0, ty.a1);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int x4 = x0 + 4 * lane;
    RzCols C;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const ResizeTab tx = P.xtab[min(x4 + i, P.dw - 1)];
        C.o[i] = tx.ofs - abase; C.a0[i] = tx.a0; C.a1[i] = tx.a1;
    }
    rz_cols_finish(C);
    __syncthreads();
    mbar_wait(&s_bar, 0);
    if (x4 > x1) return;
    const int nrows = y1 - y0 + 1;
    const int rb = wid * RZ2_ROWS, re = min(nrows, rb + RZ2_ROWS);
    if (rb >= re) return;
    uint8_t *D = P.dst + (size_t)f * P.dst_fstride + (size_t)(y0 + rb) * P.dst_step + x4;
    int X[4], Y[4];                                                                  // horizontal passes of two source rows; which one is "upper" alternates
    int r = rb;
    for (; r + 1 < re; r += 2) {
        {   // even step: the previous row's lower source row (if any) sits in X; this row leaves its lower one in Y
            const int4 e = s_rt[r];
            if (e.y >> 24) rz_hrow(s_img + e.x, C, bytewise, X);
            rz_hrow(s_img + (e.y & 0xFFFFFF), C, bytewise, Y);
            *reinterpret_cast<uint32_t *>(D) = rz_vpack(e.z, e.w, X, Y);
            D += P.dst_step;
        }
        {   // odd step: roles swapped
            const int4 e = s_rt[r + 1];
            if (e.y >> 24) rz_hrow(s_img + e.x, C, bytewise, Y);
            rz_hrow(s_img + (e.y & 0xFFFFFF), C, bytewise, X);
            *reinterpret_cast<uint32_t *>(D) = rz_vpack(e.z, e.w, Y, X);
            D += P.dst_step;
        }
    }
    if (r < re) {
        const int4 e = s_rt[r];
        if (e.y >> 24) rz_hrow(s_img + e.x, C, bytewise, X);
        rz_hrow(s_img + (e.y & 0xFFFFFF), C, bytewise, Y);
        *reinterpret_cast<uint32_t *>(D) = rz_vpack(e.z, e.w, X, Y);
    }
}

// ---- all levels in one cooperative launch ----
struct PyrParams {
    uint8_t *pyr; size_t pyr_slab;
    const ResizeTab *xtab, *ytab;
    int tw, th, nframes;
};

__global__ void __launch_bounds__(RZ_THREADS) k_pyramid_all(const __grid_constant__ LevelMaps M, PyrParams P, const FrameGeom *__restrict__ G)
{
    extern __shared__ __align__(128) uint8_t s_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int4 s_rt[RZ_MAX_TH];
    uint8_t *s_img = s_raw + ((128u - (smem_u32(s_raw) & 127u)) & 127u);
    cg::grid_group grid = cg::this_grid();
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int grp = threadIdx.x % RZ_GROUPS, band = threadIdx.x / RZ_GROUPS;
    uint32_t loads = 0;                                                              // TMA loads this CTA has waited for (mbarrier phase)
    const int nl = G->nlevels;
    for (int l = 1; l < nl; l++) {
        const LevelGeom &gs = G->lv[l - 1], &gd = G->lv[l];
        const int sh = gs.h, dw = gd.w, dh = gd.h;
        const ResizeTab *xtab = P.xtab + gd.xtab_off, *ytab = P.ytab + gd.ytab_off;
        const int ntx = (dw + P.tw - 1) / P.tw, nty = (dh + P.th - 1) / P.th;
        const int ntiles = ntx * nty * P.nframes;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int f = t / (ntx * nty), r2 = t - f * (ntx * nty), by = r2 / ntx, bx = r2 - by * ntx;
            const int x0 = bx * P.tw, y0 = by * P.th;
            const int x1 = min(x0 + P.tw, dw) - 1, y1 = min(y0 + P.th, dh) - 1;
            const int abase = __ldg(&xtab[x0].ofs) & ~15;
            const int sylo = max(0, min(__ldg(&ytab[y0].ofs), sh - 1));
            if (threadIdx.x == 0) {
                mbar_expect_tx(&s_bar, (uint32_t)(ORBX_RZ_BOX_ROWS * ORBX_TMA_BOX_BYTES));
                tma_load_3d(s_img, &M.m[l - 1], abase >> 2, sylo, f, &s_bar);
            }
            for (int r = threadIdx.x; r <= y1 - y0; r += RZ_THREADS) {
                const ResizeTab ty = ytab[y0 + r];
                const int sy0 = max(0, min(ty.ofs, sh - 1)), sy1 = max(0, min(ty.ofs + 1, sh - 1));
                s_rt[r] = make_int4((sy0 - sylo) * ORBX_TMA_BOX_BYTES, (sy1 - sylo) * ORBX_TMA_BOX_BYTES, ty.a0, ty.a1);
            }
            const int x4 = x0 + 4 * grp;
            RzCols C;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const ResizeTab tx = xtab[min(x4 + i, dw - 1)];
                C.o[i] = tx.ofs - abase; C.a0[i] = tx.a0; C.a1[i] = tx.a1;
            }
            rz_cols_finish(C);
            const bool bytewise = gs.w > 2 * dw;
            __syncthreads();                                                         // row table staged
            mbar_wait(&s_bar, loads & 1u);
            loads++;
            if (x4 <= x1) {
                const int nrows = y1 - y0 + 1, RB = (nrows + RZ_BANDS - 1) / RZ_BANDS;
                const int rb = band * RB, re = min(nrows, rb + RB);
                uint8_t *D = P.pyr + (size_t)f * P.pyr_slab + gd.off + (size_t)(y0 + rb) * gd.pitch + x4;
                int prev_off = -1, tp[4] = { 0, 0, 0, 0 };
                for (int r = rb; r < re; r++, D += gd.pitch) {
                    const int4 e = s_rt[r];
                    int t0[4], t1[4];
                    if (e.x == prev_off) {
#pragma unroll
                        for (int i = 0; i < 4; i++) t0[i] = tp[i];
                    } else {
                        rz_hrow(s_img + e.x, C, bytewise, t0);
                    }
                    if (e.y == e.x) {
#pragma unroll
                        for (int i = 0; i < 4; i++) t1[i] = t0[i];
                    } else {
                        rz_hrow(s_img + e.y, C, bytewise, t1);
                    }
                    prev_off = e.y;
#pragma unroll
                    for (int i = 0; i < 4; i++) tp[i] = t1[i];
                    *reinterpret_cast<uint32_t *>(D) = rz_vpack(e.z, e.w, t0, t1);
                }
            }
            __syncthreads();                                                         // the window and the row table are reused by the next tile
        }
        if (l + 1 < nl) {
            // level l is complete everywhere before anyone loads it: generic-proxy stores -> async-proxy (TMA) loads
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
            grid.sync();
        }
    }
}

// ComputePyramid for the batch: one cooperative launch, or one launch per level where cooperative launches are unavailable
int launch_pyramid(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride)
{
    const FrameGeom &G = h->geo;
    if (G.nlevels < 2) return 0;
    if (orbx_ensure_tmaps(h, nframes, l0, l0_step, l0_fstride) != 0) return -1;
    const size_t smem = 128 + (size_t)ORBX_RZ_BOX_ROWS * ORBX_TMA_BOX_BYTES + 16;      // + the third word rz_hrow may read past the last row
    if (h->pyr_grid_cap == 0) {
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
        h->pyr_grid_cap = -1;
        if (coop) {
            int occ = 0;
            if (orbx_optin_smem(h, (const void *)k_pyramid_all, smem) && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pyramid_all, RZ_THREADS, smem) == cudaSuccess && occ > 0) h->pyr_grid_cap = occ * h->sm_count;
        }
    }
    // measured on B200: the fused launch saves ~35 us of CPU enqueue time and a few us of GPU time per frame at batch 1, but its six
    // grid-wide barriers cost more than six kernel boundaries once the batch fills the machine (0.37 vs 0.31 ms at 128 frames)
    if (h->pyr_grid_cap > 0 && nframes <= 4) {
        LevelMaps M;
        memcpy(M.m, h->tmap_rz, sizeof(M.m));
        PyrParams P;
        P.pyr = h->d_pyr; P.pyr_slab = h->pyr_slab; P.xtab = h->d_xtab; P.ytab = h->d_ytab;
        P.tw = G.rz_tw; P.th = G.rz_th; P.nframes = nframes;
        // latency path: halve the tile height while level 1 would leave SMs without a tile (60 tiles of 192 x 64 for one 1280 x 720 frame)
        while (P.th > 16 && ((G.lv[1].w + P.tw - 1) / P.tw) * ((G.lv[1].h + P.th - 1) / P.th) * nframes < h->sm_count) P.th = (P.th / 2 + 7) & ~7;
        const int tiles1 = ((G.lv[1].w + P.tw - 1) / P.tw) * ((G.lv[1].h + P.th - 1) / P.th) * nframes;
        const int grid = std::max(1, std::min(tiles1, h->pyr_grid_cap));
        const FrameGeom *dg = h->d_geo;
        void *args[] = { (void *)&M, (void *)&P, (void *)&dg };
        ProfScope ps(h, ORBX_K_RESIZE);
        if (cudaLaunchCooperativeKernel((const void *)k_pyramid_all, dim3(grid), dim3(RZ_THREADS), args, smem, h->stream) == cudaSuccess) return 0;
        cudaGetLastError();
        h->pyr_grid_cap = -1;                                      // fall back for good
    }
    for (int l = 1; l < G.nlevels; l++) if (launch_resize_level(h, l, nframes, l0, l0_step, l0_fstride) != 0) return -1;
    return 0;
}

int launch_resize_level(orbx_handle *h, int level, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride)
{
    const LevelGeom &gs = h->geo.lv[level - 1], &gd = h->geo.lv[level];
    if (orbx_ensure_tmaps(h, nframes, l0, l0_step, l0_fstride) != 0) return -1;
    LevelMaps M;
    memcpy(M.m, h->tmap_rz, sizeof(M.m));
    ResizeParams P;
    P.dst = h->d_pyr + gd.off; P.dst_step = (size_t)gd.pitch; P.dst_fstride = h->pyr_slab;
    P.sw = gs.w; P.sh = gs.h; P.dw = gd.w; P.dh = gd.h;
    P.xtab = h->d_xtab + gd.xtab_off; P.ytab = h->d_ytab + gd.ytab_off;
    P.tw = h->geo.rz_tw; P.th = h->geo.rz_th; P.src_level = level - 1;
    const size_t smem = 128 + (size_t)ORBX_RZ_BOX_ROWS * ORBX_TMA_BOX_BYTES + 16;      // + the third word rz_hrow may read past the last row
    if (!orbx_optin_smem(h, (const void *)k_resize_linear, smem)) return -1;
    dim3 grid((gd.w + P.tw - 1) / P.tw, (gd.h + P.th - 1) / P.th, nframes);
    ProfScope ps(h, ORBX_K_RESIZE);
    // levels >= 2 directly follow the launch that writes their source: allow them to start while it drains (see the kernel)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(RZ_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (level >= 2 && !h->prof_on && h->opt_pdl) ? 1 : 0;
#ifndef ORBX_RESIZE_V1
    if (h->geo.rz2_ok) {                                                       // the 128 x 64 tile's source window fits the 192-byte box at every level
        memcpy(M.m, h->tmap_rz2, sizeof(M.m));
        const size_t smem2 = 128 + (size_t)ORBX_RZ_BOX_ROWS * ORBX_RZ2_BOX_BYTES + 16;
        auto kern = P.sw > 2 * P.dw ? k_resize_linear2<true> : k_resize_linear2<false>;      // scales above 2 take the byte-wise horizontal pass
        if (!orbx_optin_smem(h, (const void *)kern, smem2)) return -1;
        cfg.gridDim = dim3((gd.w + RZ2_TW - 1) / RZ2_TW, (gd.h + RZ2_TH - 1) / RZ2_TH, nframes); cfg.blockDim = dim3(RZ2_WARPS * 32); cfg.dynamicSmemBytes = smem2;
        if (cudaLaunchKernelEx(&cfg, kern, M, P) != cudaSuccess) { cudaGetLastError(); kern<<<cfg.gridDim, RZ2_WARPS * 32, smem2, h->stream>>>(M, P); }
        return 0;
    }
#endif
    if (cudaLaunchKernelEx(&cfg, k_resize_linear, M, P) != cudaSuccess) { cudaGetLastError(); k_resize_linear<<<grid, RZ_THREADS, smem, h->stream>>>(M, P); }
    return 0;
}
