// k_blur.cu — 7x7 sigma-2 Gaussian smoothing of every pyramid level, integer 8.8 fixed point.
// Replaces `GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101)` on the cloned
// level (reference ORBextractor.cpp:1132-1133).  Arithmetic: SURVEY.md App. A.3 — taps
// [18,34,48,56,48,34,18]/256, row pass in u16, column pass in u32, one rounding (c + 32768) >> 16.
// Not on the default path any more: the descriptor kernel evaluates the same filter at the pixels it reads (k_describe.cu).  This
// kernel materialises whole blurred levels for orbx_get_blurred_level and for ORBX_OPT_FUSED_BLUR = 0.
//
// HBM-bound stage, all levels of all frames in ONE launch.  A CTA produces a 256 x hCell output tile: its input
// window (3-px halo, 288 bytes x (hCell + 6) rows) arrives by ONE TMA load (the per-level tensor maps shared with
// k_fast.cu; the unit zero-fills outside the image), border tiles then mirror the missing halo in shared memory
// (reflect-101), and every thread sweeps 4 adjacent columns (one aligned 32-bit word) down half the tile: three
// LDS.32 per input row, the seven-tap row pass as two IDP.4A dot products per pixel on funnel-shifted windows,
// the column pass on a 7-deep register window per column (rows unrolled by 7 so the window rotates statically),
// one PRMT tree to pack the four results and one 32-bit coalesced store.  No per-row address or border arithmetic.
#include "orbx_internal.h"
#include "orbx_tma.h"
#include <cstring>

#define BL_THREADS 128
#define BL_GROUPS 64                 // 4-pixel column groups per tile row: 256 output columns
#define BL_TW (4 * BL_GROUPS)
#define BL_PADROWS 8                 // rows behind the tile: the unrolled sweep may overrun a band by < 7 rows

struct BlurParams {
    uint8_t *blur; size_t blur_slab;
    const uint32_t *tiles;           // level:4 | tile column:12 | tile row:16
    int tile_rows;                   // max (hCell + 6) over the levels
};

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

__global__ void __launch_bounds__(BL_THREADS) k_blur7(const __grid_constant__ LevelMaps M, BlurParams P, const FrameGeom *__restrict__ G)
{
    extern __shared__ __align__(128) uint8_t s_raw[];
    __shared__ __align__(8) uint64_t s_bar;
    uint8_t *s_img = s_raw + ((128u - (smem_u32(s_raw) & 127u)) & 127u);          // TMA destination: 128-byte aligned
    const int f = blockIdx.y;
    const uint32_t td = __ldg(P.tiles + blockIdx.x);
    const int level = (int)(td & 15u), tx = (int)((td >> 4) & 0xFFFu), ty = (int)(td >> 16);
    const LevelGeom &g = G->lv[level];
    const int w = g.w, hgt = g.h, TH = g.hcell;                                     // tile = BL_TW x hCell output pixels
    const int x0 = tx * BL_TW, y0 = ty * TH;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&s_bar, (uint32_t)((TH + 6) * ORBX_TMA_BOX_BYTES));
        // box column 0 = image column x0 - 16, box row 0 = image row y0 - 3
        tma_load_3d(s_img, &M.m[level], (x0 >> 2) - 4, y0 - 3, f, &s_bar);
    }
    __syncthreads();                                                                 // barrier initialised before anyone polls it
    mbar_wait(&s_bar, 0);
    // ---- reflect-101 halo of border tiles (image pixel (x, y) lives at s_img[(y - y0 + 3) * 288 + x - x0 + 16]) ----
    const bool edge_l = x0 == 0, edge_r = x0 + BL_TW + 3 > w, edge_t = y0 < 3, edge_b = y0 + TH + 3 > hgt;
    if (edge_l || edge_r) {
        for (int r = threadIdx.x; r < TH + 6; r += BL_THREADS) {
            const int y = y0 - 3 + r;
            if (y < 0 || y >= hgt) continue;
            uint8_t *row = s_img + r * ORBX_TMA_BOX_BYTES + 16 - x0;                 // row[x] = pixel x
            if (edge_l) { row[-1] = row[1]; row[-2] = row[2]; row[-3] = row[3]; }
            if (edge_r) { row[w] = row[w - 2]; row[w + 1] = row[w - 3]; row[w + 2] = row[w - 4]; }
        }
        __syncthreads();
    }
    if (edge_t || edge_b) {
        // rows above / below the image mirror the (column-patched) rows inside it: y -> -y resp. 2(h-1) - y
        uint32_t *sw = reinterpret_cast<uint32_t *>(s_img);
        for (int i = threadIdx.x; i < (TH + 6) * ORBX_TMA_BOX_WORDS; i += BL_THREADS) {
            const int r = i / ORBX_TMA_BOX_WORDS, wd = i - r * ORBX_TMA_BOX_WORDS;
            const int y = y0 - 3 + r;
            if (y >= 0 && y < hgt) continue;
            const int ys = y < 0 ? -y : 2 * hgt - 2 - y, rs = ys - y0 + 3;
            if (ys >= 0 && ys < hgt && rs >= 0 && rs < TH + 6) sw[r * ORBX_TMA_BOX_WORDS + wd] = sw[rs * ORBX_TMA_BOX_WORDS + wd];
        }
        __syncthreads();
    }
    // ---- sweep: thread = 4 columns x half the tile rows ----
    const int grp = threadIdx.x & (BL_GROUPS - 1), band = threadIdx.x / BL_GROUPS;
    const int x = x0 + 4 * grp;
    if (x >= w) return;
    const int RB = (TH + 1) >> 1;
    const int rb = band * RB, re = min(min(TH, rb + RB), hgt - y0);                  // tile-relative output rows [rb, re)
    if (rb >= re) return;
    const uint32_t *colw = reinterpret_cast<const uint32_t *>(s_img) + 4 + grp + rb * ORBX_TMA_BOX_WORDS;   // word of input row (rb - 3)
    uint8_t *dst = P.blur + (size_t)f * P.blur_slab + g.boff + (size_t)(y0 + rb) * g.bpitch + x;

    int win[4][7];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const uint32_t *q = colw + k * ORBX_TMA_BOX_WORDS;
        hpass4(q[-1], q[0], q[1], win[0][k], win[1][k], win[2][k], win[3][k]);
    }
    const int nrows = re - rb;
    for (int r0 = 0; r0 < nrows; r0 += 7) {
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const uint32_t *q = colw + (r0 + k + 6) * ORBX_TMA_BOX_WORDS;
            hpass4(q[-1], q[0], q[1], win[0][(k + 6) % 7], win[1][(k + 6) % 7], win[2][(k + 6) % 7], win[3][(k + 6) % 7]);
            uint32_t acc[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                acc[j] = 32768u + 18u * (uint32_t)(win[j][k % 7] + win[j][(k + 6) % 7]) +
                         34u * (uint32_t)(win[j][(k + 1) % 7] + win[j][(k + 5) % 7]) +
                         48u * (uint32_t)(win[j][(k + 2) % 7] + win[j][(k + 4) % 7]) + 56u * (uint32_t)win[j][(k + 3) % 7];
            // byte 2 of each accumulator = (acc >> 16) & 255 (acc < 2^24)
            const uint32_t packed = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
            if (r0 + k < nrows) *reinterpret_cast<uint32_t *>(dst + (size_t)(r0 + k) * g.bpitch) = packed;   // pitch % 128 == 0: the word is in-bounds
        }
    }
}

// tiles [tile_first, tile_first + ntiles) of the level-ordered tile table (ntiles < 0: to the end)
int launch_blur(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride, cudaStream_t st, int tile_first, int ntiles)
{
    if (ntiles < 0) ntiles = h->geo.total_blur_tiles - tile_first;
    if (ntiles <= 0) return 0;
    if (orbx_ensure_tmaps(h, nframes, l0, l0_step, l0_fstride) != 0) return -1;
    LevelMaps M;
    memcpy(M.m, h->tmap, sizeof(M.m));
    BlurParams P;
    P.blur = h->d_blur; P.blur_slab = h->blur_slab;
    P.tiles = h->d_blur_tiles + tile_first;
    P.tile_rows = h->geo.max_hcell + 6;
    const size_t smem = 128 + (size_t)(P.tile_rows + BL_PADROWS) * ORBX_TMA_BOX_BYTES;
    if (!orbx_optin_smem(h, (const void *)k_blur7, smem)) return -1;
    dim3 grid(ntiles, nframes);
    ProfScope ps(h, ORBX_K_BLUR, st);
    k_blur7<<<grid, BL_THREADS, smem, st>>>(M, P, h->d_geo);
    return 0;
}
