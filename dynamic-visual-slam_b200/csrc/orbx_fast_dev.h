// orbx_fast_dev.h — device pieces shared by the two FAST formulations (k_fast.cu: warp per cell; k_fast_dense.cu: dense tiles):
// the packed 4-pixel pre-test sweep, the exact 16-arc score on u16x2 lanes and the deferred corner-list publish.
#pragma once
#include "orbx_internal.h"

#define RO(dx, dy) ((dy) * TP + (dx))

// S on the raw ring values: S = max( I(p) - min_k max9_k(ring), max_k min9_k(ring) - I(p) ).
// Packed lanes: lo16 = r, hi16 = 255 - r  =>  a lane-wise min yields (min r, 255 - max r).
// min9_k = min3( min3(r_k..r_k+2), min3(r_k+3..r_k+5), min3(r_k+6..r_k+8) ): 40 three-input min/max in all.
__device__ __forceinline__ uint32_t vmin3u2(uint32_t a, uint32_t b, uint32_t c) { return __vminu2(__vminu2(a, b), c); }
__device__ __forceinline__ uint32_t vmax3u2(uint32_t a, uint32_t b, uint32_t c) { return __vmaxu2(__vmaxu2(a, b), c); }
template <int TP> __device__ __forceinline__ int fast_score_packed(const uint8_t *p)
{
    const int v = p[0];
    uint32_t r[16];
#define PK(x) ((uint32_t)(x) * 0xFFFF0001u + 0x00FF0000u)
    r[0] = PK(p[RO(0, 3)]);   r[1] = PK(p[RO(1, 3)]);    r[2] = PK(p[RO(2, 2)]);    r[3] = PK(p[RO(3, 1)]);
    r[4] = PK(p[RO(3, 0)]);   r[5] = PK(p[RO(3, -1)]);   r[6] = PK(p[RO(2, -2)]);   r[7] = PK(p[RO(1, -3)]);
    r[8] = PK(p[RO(0, -3)]);  r[9] = PK(p[RO(-1, -3)]);  r[10] = PK(p[RO(-2, -2)]); r[11] = PK(p[RO(-3, -1)]);
    r[12] = PK(p[RO(-3, 0)]); r[13] = PK(p[RO(-3, 1)]);  r[14] = PK(p[RO(-2, 2)]);  r[15] = PK(p[RO(-1, 3)]);
#undef PK
    uint32_t m3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m3[k] = vmin3u2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
    uint32_t m9[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m9[k] = vmin3u2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
    uint32_t b4[4];
#pragma unroll
    for (int k = 0; k < 4; k++) b4[k] = __vmaxu2(vmax3u2(m9[4 * k], m9[4 * k + 1], m9[4 * k + 2]), m9[4 * k + 3]);
    const uint32_t best = __vmaxu2(__vmaxu2(b4[0], b4[1]), __vmaxu2(b4[2], b4[3]));   // per lane: max_k min9_k
    const int hi_of_min = (int)(best & 0xFFFFu);                  // max_k min9_k(r)
    const int lo_of_max = 255 - (int)(best >> 16);                // min_k max9_k(r)
    const int s_dark = v - lo_of_max, s_bright = hi_of_min - v;
    return s_dark > s_bright ? s_dark : s_bright;
}

// one tile row entering the sweep window: the thread's own word C plus the four shifted views of it
struct FastRow { uint32_t C, P2, M2, P3, M3; };
__device__ __forceinline__ FastRow fast_row(const uint32_t *q)
{
    const uint32_t L = q[-1], C = q[0], R = q[1];
    FastRow w;
    w.C = C;
    w.P2 = __byte_perm(C, R, 0x5432);      // columns x+2 .. x+5
    w.M2 = __byte_perm(L, C, 0x5432);      // columns x-2 .. x+1
    w.P3 = __byte_perm(C, R, 0x6543);      // columns x+3 .. x+6
    w.M3 = __byte_perm(L, C, 0x4321);      // columns x-3 .. x
    return w;
}

// Pre-test of R <= 8 detection rows x 4 pixels (8 flag bits per pixel column).  q = the item's word in tile row r0 (= ring row dy = -3 of the first
// detection row).  Result: bit (7-k) of byte j set iff pixel (row r0 + k, byte j) may be a corner at threshold T.
#ifndef FAST_RMAX
#define FAST_RMAX 16                 // tallest sweep unit: a 37-row cell then takes 3 units per word column (30 units = ONE warp iteration at 94 %
                                     // lane use) instead of 5 units of 8 rows (50 units = two iterations at 78 %)
#endif
#ifndef FAST_PAIRS
#define FAST_PAIRS 2                 // opposite ring pairs tested in the sweep: 4 = (0,8) (4,12) (2,10) (6,14); 2 = (0,8) (4,12) only (measured: 0.474 -> 0.450 ms per 128 frames)
#endif
#if FAST_PAIRS == 4
template <int TP> __device__ __noinline__ uint2 fast_sweep7(const uint32_t *q, uint32_t HM, uint32_t KK, int R)
{
    FastRow w[7];
#pragma unroll
    for (int k = 0; k < 6; k++) w[k] = fast_row(q + k * (TP / 4));
    uint32_t fl0 = 0u, fl1 = 0u;                                                 // rows 0..7 and rows 8..15 of the unit
#pragma unroll
    for (int k = 0; k < FAST_RMAX; k++) {
        if (k >= R) break;                                                       // units of R <= FAST_RMAX rows (uniform); the window indices are mod 7
        w[(k + 6) % 7] = fast_row(q + (k + 6) * (TP / 4));                       // ring row dy = +3 of detection row k
        const uint32_t C0 = w[(k + 3) % 7].C;
        const uint32_t p08 = __vabsdiffu4(w[(k + 6) % 7].C, C0) | __vabsdiffu4(w[k % 7].C, C0);
        const uint32_t p4c = __vabsdiffu4(w[(k + 3) % 7].P3, C0) | __vabsdiffu4(w[(k + 3) % 7].M3, C0);
        const uint32_t p2a = __vabsdiffu4(w[(k + 5) % 7].P2, C0) | __vabsdiffu4(w[(k + 1) % 7].M2, C0);
        const uint32_t p6e = __vabsdiffu4(w[(k + 1) % 7].P2, C0) | __vabsdiffu4(w[(k + 5) % 7].M2, C0);
        const uint32_t t0 = p08 & HM, t1 = p4c & HM, t2 = p2a & HM, t3 = p6e & HM;
        uint32_t acc = t0 | (t0 + KK);
        acc &= t1 | (t1 + KK);
        acc &= t2 | (t2 + KK);
        acc &= t3 | (t3 + KK);
        if (k < 8) fl0 |= (acc >> k) & (0x80808080u >> k);
        else fl1 |= (acc >> (k - 8)) & (0x80808080u >> (k - 8));
    }
    return make_uint2(fl0, fl1);
}
#else
// Two-pair sweep: the vertical pair (0,8) and the horizontal pair (4,12) only.  On textured frames they alone reject 97.5 % of the pixels
// (all four: 97.8 %), for half the arithmetic: a row needs its own word plus the two 3-byte-shifted views when it is the centre row,
// and the vertical difference |I(y+3) - I(y)| of a row serves the centres y and y+3 (it is kept three rows).
// RR > 0: the unit height as a compile-time constant — the row loop is then straight-line code (no exit test per row), so the loads and the
// SWAR chains of different rows interleave; RR = 0: any height up to FAST_RMAX, one exit test per row.
template <int TP, int RR> __device__ __forceinline__ uint2 fast_sweep7_body(const uint32_t *q, uint32_t HM, uint32_t KK, int R)
{
    uint32_t c[7], dv[3];                                                        // c: rows k..k+6 (mod 7); dv[j % 3] = |C(j+3) - C(j)|
#pragma unroll
    for (int k = 0; k < 6; k++) c[k] = q[k * (TP / 4)];
#pragma unroll
    for (int j = 0; j < 3; j++) dv[j] = __vabsdiffu4(c[j + 3], c[j]);
    uint32_t fl0 = 0u, fl1 = 0u;
#pragma unroll
    for (int k = 0; k < (RR > 0 ? RR : FAST_RMAX); k++) {
        if (RR == 0 && k >= R) break;
        const uint32_t *row = q + (k + 3) * (TP / 4);                            // the centre row of detection row k
        const uint32_t L = row[-1], Rw = row[1];
        c[(k + 6) % 7] = q[(k + 6) * (TP / 4)];
        const uint32_t C0 = c[(k + 3) % 7];
        const uint32_t up = dv[k % 3];                                           // |C(k+3) - C(k)|: ring pixel 8 (dy = -3) of centre k+3
        const uint32_t dn = __vabsdiffu4(c[(k + 6) % 7], C0);                    // ring pixel 0 (dy = +3)
        dv[k % 3] = dn;                                                          // = |C(j+3) - C(j)| for j = k+3, needed again at k+3
        const uint32_t t0 = (up | dn) & HM;
        const uint32_t t1 = (__vabsdiffu4(__byte_perm(C0, Rw, 0x6543), C0) | __vabsdiffu4(__byte_perm(L, C0, 0x4321), C0)) & HM;
        const uint32_t acc = (t0 | (t0 + KK)) & (t1 | (t1 + KK));
        if (k < 8) fl0 |= (acc >> k) & (0x80808080u >> k);
        else fl1 |= (acc >> (k - 8)) & (0x80808080u >> (k - 8));
    }
    return make_uint2(fl0, fl1);
}
#ifndef FAST_FIXED_R
#define FAST_FIXED_R 1               // 1: straight-line instances for the unit heights of ~36-row cells (12, 13); 0: the generic loop only
#endif
template <int TP, int RR> __device__ __noinline__ uint2 fast_sweep7_fixed(const uint32_t *q, uint32_t HM, uint32_t KK) { return fast_sweep7_body<TP, RR>(q, HM, KK, RR); }
template <int TP> __device__ __noinline__ uint2 fast_sweep7_any(const uint32_t *q, uint32_t HM, uint32_t KK, int R) { return fast_sweep7_body<TP, 0>(q, HM, KK, R); }
template <int TP> __device__ __forceinline__ uint2 fast_sweep7(const uint32_t *q, uint32_t HM, uint32_t KK, int R)
{
#if FAST_FIXED_R
    if (R == 12) return fast_sweep7_fixed<TP, 12>(q, HM, KK);                   // R is uniform over the warp (a property of the cell)
    if (R == 13) return fast_sweep7_fixed<TP, 13>(q, HM, KK);
#endif
    return fast_sweep7_any<TP>(q, HM, KK, R);
}
#endif

// loose pre-test threshold T = 2^sh - 1 <= th:  |d| > T  <=>  (|d| & HM) != 0;  t + KK sets bit 7 of every byte with t >= 2^sh
__device__ __forceinline__ void fast_masks(int th, uint32_t &HM, uint32_t &KK)
{
    const int sh = min(7, 31 - __clz(th + 1));
    HM = ((0xFFu << sh) & 0xFFu) * 0x01010101u;
    KK = (0x80u - (1u << sh)) * 0x01010101u;
}


// the warp's result list -> the (frame, level) corner list at the slot range a counter atomic reserved
#ifndef FS_RES
#define FS_RES 64
#endif
__device__ __forceinline__ void fast_write_out(const uint32_t *res, int n, int base, uint32_t *gdst, int cap, int32_t *status, int lane)
{
    for (int i = lane; i < n; i += 32) {
        if (base + i < cap) gdst[base + i] = res[i];
        else atomicOr(status, ORBX_DS_CAND_OVERFLOW);
    }
}
// synchronous variant (a cell with more results than the list holds)
static __device__ __noinline__ void fast_publish(const uint32_t *res, int n, int32_t *gcnt, uint32_t *gdst, int cap, int32_t *status, int lane)
{
    __syncwarp();
    int base = 0;
    if (lane == 0) base = atomicAdd(gcnt, n);
    fast_write_out(res, n, __shfl_sync(0xffffffffu, base, 0), gdst, cap, status, lane);
    __syncwarp();
}

