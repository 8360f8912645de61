// k_fast.cu — FAST-9/16 with per-cell 3x3 NMS and the two-threshold fallback.
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (reference ORBextractor.cpp:785-872):
// for every ~35-px cell (+6 px overlap) `cv::FAST(roi, kps, iniThFAST, true)` and, iff that returned
// nothing, `cv::FAST(roi, kps, minThFAST, true)` (:826-827, :845-846).  FAST arithmetic: SURVEY App. A.2
//   S = max over the 16 arcs of 9 contiguous ring pixels of min(I(p)-I(ring)) resp. min(I(ring)-I(p));
//   corner iff S > th; score = S - 1; strict 3x3 NMS INSIDE the cell's ROI (scores outside the
//   cell's detection area read as 0).
//
// Design.  A work item is a horizontal run of cells of one cell row of one level of one frame (a "strip",
// described by a host-built table).  The kernel is PERSISTENT: one CTA per resident slot loops over the items of
// the whole batch, and the strip's ROI arrives in shared memory by TMA (cp.async.bulk.tensor, one descriptor per
// pyramid level, frames as the third tensor dimension) into a double buffer: the tile of item i+1 is in flight
// while item i is processed, so no thread ever issues a staging load or waits for one.
//   1. packed sweep, block-wide: a sweep unit is one aligned 32-bit word of the tile (4 adjacent pixels) x 7
//      rows, walked with a 7-row register window (3 LDS.32 + 4 PRMT per row of 4 pixels).  Per row it
//      evaluates a polarity-agnostic pre-test on the four opposite ring pairs (0,8) (4,12) (2,10) (6,14):
//      |I(ring) - I(p)| for 4 pixels is ONE VABSDIFF4.U8; "some member of the pair differs by more than T"
//      with T = 2^k - 1 <= th is an OR, a mask and one add per pair (SWAR, no per-byte compares).  Every
//      9-arc contains a member of each opposite pair, so the test is an exact NECESSARY condition for
//      S > th; it is loose by design (T <= th, sign ignored) and passes ~5 % of the pixels.  Survivors go
//      straight into one shared queue (warp-aggregated slot allocation);
//   2. exact score of the queued survivors, dense over the block: the 16-arc min/max network runs on packed
//      u16x2 lanes (VIMNMX3.U16x2): low half = ring value, high half = 255 - ring value, so one instruction
//      serves the darker and the brighter polarity;
//   3. strict 3x3 NMS over the queue entries; neighbours in another cell are never read (the reference runs
//      FAST per cell ROI, so they count as 0);
//   4. a cell with no keypoint at iniThFAST is swept again at minThFAST (the reference's retry) by ONE warp —
//      only that cell, not the strip.
// Survivors are appended to the (frame, level) candidate list with one global atomic per item;
// list order is arbitrary (the quadtree kernel is order-independent).
//
// Toolchain note: an earlier formulation on signed differences (d = I(p) - I(ring), score via
// max(mn9, -mx9)) produced wrong results on sm_100a with nvcc 12.9 (the negation feeding a fused
// 3-input VIMNMX3 was lost); the raw-value formulation below has no negated min/max operands.
#include "orbx_internal.h"
#include "orbx_tma.h"
#include <algorithm>
#include <cstring>

#define FS_THREADS 192
#define FS_WARPS (FS_THREADS / 32)
#define FS_TP 288                // tile pitch = TMA box width: 72 u32 elements >= 16 + 15 + ORBX_FAST_MAX_W + 6 + 4
#define FS_TPW (FS_TP / 4)
#define FS_PADROWS 6             // rows behind the tile: a 7-row sweep unit may start on the last detection row
#define FS_SP 272                // score-map pitch: >= detection width + 2, multiple of 16
#define FS_QCAP 4096             // survivor queue (u16 tile offsets); FS_RWARPS x FS_WQ in the retry phase
#define FS_RWARPS 4               // warps that run cell retries concurrently
#define FS_WQ (FS_QCAP / FS_RWARPS)  // per-warp queue of the retry phase, >= 32 lanes x 28 flags
#define FS_OUT_CAP 512          // staged outputs; beyond it survivors are written straight to the global list
#define FS_MAX_CELLS 8
#ifndef FS_NBUF
#define FS_NBUF 2                // tile buffers: 2 = the next item's tile is in flight while this one is processed
#endif

struct FastParams {
    uint32_t *cand; size_t cand_slab;
    int32_t *ncand;
    const uint32_t *strips;      // level:4 | cells:4 | cell row:12 | first cell column:12
    int nstrips, nitems;         // items = nstrips x frames
    int ini_th, min_th;
    int32_t *status;
    int tile_rows;               // max (hCell + 6) over the levels
};

#define RO(dx, dy) ((dy) * FS_TP + (dx))

// S on the raw ring values: S = max( I(p) - min_k max9_k(ring), max_k min9_k(ring) - I(p) ).
// Packed lanes: lo16 = r, hi16 = 255 - r  =>  a lane-wise min yields (min r, 255 - max r).
// min9_k = min3( min3(r_k..r_k+2), min3(r_k+3..r_k+5), min3(r_k+6..r_k+8) ): 40 three-input min/max in all.
__device__ __forceinline__ uint32_t vmin3u2(uint32_t a, uint32_t b, uint32_t c) { return __vminu2(__vminu2(a, b), c); }
__device__ __forceinline__ uint32_t vmax3u2(uint32_t a, uint32_t b, uint32_t c) { return __vmaxu2(__vmaxu2(a, b), c); }
__device__ __forceinline__ int fast_score_packed(const uint8_t *p)
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

// Pre-test of 7 detection rows x 4 pixels.  q = the item's word in tile row r0 (= ring row dy = -3 of the first
// detection row).  Result: bit (7-k) of byte j set iff pixel (row r0 + k, byte j) may be a corner at threshold T.
__device__ __noinline__ uint32_t fast_sweep7(const uint32_t *q, uint32_t HM, uint32_t KK)
{
    FastRow w[7];
#pragma unroll
    for (int k = 0; k < 6; k++) w[k] = fast_row(q + k * FS_TPW);
    uint32_t fl = 0u;
#pragma unroll
    for (int k = 0; k < 7; k++) {
        w[(k + 6) % 7] = fast_row(q + (k + 6) * FS_TPW);                       // ring row dy = +3 of detection row k
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
        fl |= (acc >> k) & (0x80808080u >> k);
    }
    return fl;
}

// loose pre-test threshold T = 2^sh - 1 <= th:  |d| > T  <=>  (|d| & HM) != 0;  t + KK sets bit 7 of every byte with t >= 2^sh
__device__ __forceinline__ void fast_masks(int th, uint32_t &HM, uint32_t &KK)
{
    const int sh = min(7, 31 - __clz(th + 1));
    HM = ((0xFFu << sh) & 0xFFu) * 0x01010101u;
    KK = (0x80u - (1u << sh)) * 0x01010101u;
}

// geometry of one work item, derived from its strip descriptor
struct FastItem { int f, level, ci, cj0, ncell, wcell, hcell, iniX, iniY, ax, rw, rh, dw, dh; };
__device__ __forceinline__ FastItem fast_item(const FastParams &P, const FrameGeom *__restrict__ G, int item)
{
    FastItem t;
    t.f = item / P.nstrips;
    const uint32_t sd = __ldg(P.strips + (item - t.f * P.nstrips));
    t.level = (int)(sd & 15u); t.ncell = (int)((sd >> 4) & 15u); t.ci = (int)((sd >> 8) & 0xFFFu); t.cj0 = (int)(sd >> 20);
    const LevelGeom &g = G->lv[t.level];
    t.wcell = g.wcell; t.hcell = g.hcell;
    // strip ROI in image coordinates — ORBextractor.cpp:805-822 (cells cj0 .. cj0+ncell-1 of cell row ci)
    t.iniX = ORBX_BORDER + t.cj0 * t.wcell; t.iniY = ORBX_BORDER + t.ci * t.hcell;
    const int maxX = min(t.iniX + t.ncell * t.wcell + 6, g.w - ORBX_BORDER), maxY = min(t.iniY + t.hcell + 6, g.h - ORBX_BORDER);
    t.rw = maxX - t.iniX; t.rh = maxY - t.iniY;
    t.dw = t.rw - 6; t.dh = t.rh - 6;          // detection area of the strip, ROI-relative origin (3,3)
    t.ax = 16 + (t.iniX & 15);                 // tile byte of ROI column 0: TMA boxes start on 16-byte columns; one 16-byte unit of left margin
    return t;
}

__global__ void __launch_bounds__(FS_THREADS) k_fast_cells(const __grid_constant__ LevelMaps M, FastParams P, const FrameGeom *__restrict__ G)
{
    extern __shared__ __align__(128) uint8_t s_dyn_raw[];
    uint8_t *s_dyn = s_dyn_raw + ((128u - (smem_u32(s_dyn_raw) & 127u)) & 127u);      // TMA destinations are 128-byte aligned
    __shared__ uint16_t s_q[FS_QCAP];
    __shared__ uint32_t s_out[FS_OUT_CAP];
    __shared__ __align__(8) uint64_t s_full[2];
    __shared__ int s_qn, s_nout;
    __shared__ int s_ccnt[FS_MAX_CELLS];
    const int tile_bytes = ((P.tile_rows + FS_PADROWS) * FS_TP + 127) & ~127;
    uint8_t *s_sc = s_dyn + FS_NBUF * tile_bytes;                          // (tile_rows - 4) x FS_SP score map with a zero ring
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nl = G->nlevels;

    if (threadIdx.x == 0) {
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    // prologue: the first item's tile
    if (FS_NBUF == 2 && threadIdx.x == 0 && (int)blockIdx.x < P.nitems) {
        const FastItem t = fast_item(P, G, blockIdx.x);
        mbar_expect_tx(&s_full[0], (uint32_t)((t.hcell + 6) * FS_TP));
        tma_load_3d(s_dyn, &M.m[t.level], ((t.iniX & ~15) >> 2) - 4, t.iniY, t.f, &s_full[0]);
    }

    int it_n = 0;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x, it_n++) {
        const int buf = FS_NBUF == 2 ? (it_n & 1) : 1;
        // next item's tile into the other buffer (its previous readers passed the barrier that ended the last iteration)
        if (threadIdx.x == 0 && (FS_NBUF == 1 || item + (int)gridDim.x < P.nitems)) {
            const FastItem t = fast_item(P, G, FS_NBUF == 2 ? item + gridDim.x : item);
            mbar_expect_tx(&s_full[buf ^ 1], (uint32_t)((t.hcell + 6) * FS_TP));
            tma_load_3d(s_dyn + (FS_NBUF == 2 ? (buf ^ 1) : 0) * tile_bytes, &M.m[t.level], ((t.iniX & ~15) >> 2) - 4, t.iniY, t.f, &s_full[buf ^ 1]);
        }
        const FastItem T = fast_item(P, G, item);
        const LevelGeom &g = G->lv[T.level];
        const int ax = T.ax, dw = T.dw, dh = T.dh, wcell = T.wcell;
        const uint8_t *s_img = s_dyn + (FS_NBUF == 2 ? buf : 0) * tile_bytes;
        // zero the score map (1-px ring included) while the tile lands
        for (int i = threadIdx.x; i < ((dh + 2) * FS_SP) / 16; i += FS_THREADS) reinterpret_cast<uint4 *>(s_sc)[i] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x == 0) { s_nout = 0; s_qn = 0; }
        if (threadIdx.x < FS_MAX_CELLS) s_ccnt[threadIdx.x] = 0;
        mbar_wait(&s_full[FS_NBUF == 2 ? buf : 0], (uint32_t)(FS_NBUF == 2 ? ((it_n >> 1) & 1) : (it_n & 1)));
        __syncthreads();

        const int w0 = (ax + 3) >> 2;                                  // tile word holding detection column 0
        const int nGs = ((ax + 3 + dw - 1) >> 2) - w0 + 1;             // words holding detection columns (<= 64)
        const int nseg = (dh + 6) / 7;
        const uint32_t *words = reinterpret_cast<const uint32_t *>(s_img);
        uint32_t *gdst = P.cand + (size_t)T.f * P.cand_slab + g.cand_off;
        int32_t *gcnt = &P.ncand[T.f * nl + T.level];
        const float invc = 1.0f / (float)wcell;
        // pass 0: every cell at iniThFAST.  pass 1: the cells that produced nothing, at minThFAST (the reference's retry, :843-846).
        // Both passes are block-wide, so the retry of one or two cells is spread over all warps instead of stalling them.
        int nredo = 0;
        uint32_t redo_cells = 0u, redo_mask = 0u;                       // pass 1: up to 8 cell indices, 4 bits each / bit per cell (uniform)
        for (int pass = 0; pass < 2; pass++) {
            const int th = pass == 0 ? P.ini_th : P.min_th;
            uint32_t HM, KK;
            fast_masks(th, HM, KK);
            // ---- packed sweep: pass 0 = all words of the strip, pass 1 = the words of the retried cells ----
            int units = nGs * nseg;
            if (pass == 1) {
                units = 0;
                for (int r = 0; r < nredo; r++) {
                    const int c = (int)((redo_cells >> (4 * r)) & 15u), c_lo = c * wcell, c_hi = min(c_lo + wcell, dw);
                    units += ((((ax + 3 + c_hi - 1) >> 2) - ((ax + 3 + c_lo) >> 2)) + 1) * nseg;
                }
            }
            for (int u0 = 0; u0 < units; u0 += FS_THREADS) {
                const int u = u0 + threadIdx.x;
                uint32_t word = 0u;
                int base_off = 0;
                if (u < units) {
                    int uu = u, gfirst = 0, nG = nGs, c_lo = 0, c_hi = dw;
                    if (pass == 1) {
                        for (int r = 0; r < nredo; r++) {
                            const int c = (int)((redo_cells >> (4 * r)) & 15u);
                            c_lo = c * wcell; c_hi = min(c_lo + wcell, dw);
                            gfirst = ((ax + 3 + c_lo) >> 2) - w0; nG = ((ax + 3 + c_hi - 1) >> 2) - w0 - gfirst + 1;
                            if (uu < nG * nseg) break;
                            uu -= nG * nseg;
                        }
                    }
                    const int seg = __float2int_rd(((float)uu + 0.5f) / (float)nG), gidx = gfirst + uu - seg * nG;
                    const uint32_t raw = fast_sweep7(words + (7 * seg) * FS_TPW + w0 + gidx, HM, KK);
                    const int cb = 4 * (w0 + gidx) - (ax + 3);                // detection column of byte 0
                    uint32_t cm = 0u;
#pragma unroll
                    for (int j = 0; j < 4; j++) if (cb + j >= c_lo && cb + j < c_hi) cm |= 0xFEu << (8 * j);
                    const int nv = min(7, dh - 7 * seg);
                    word = raw & cm & (((0xFF00u >> nv) & 0xFFu) * 0x01010101u);
                    base_off = (7 * seg + 3) * FS_TP + 4 * (w0 + gidx);
                }
                const int cnt = __popc(word);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                int wbase = 0;
                if (lane == 31 && incl > 0) wbase = atomicAdd(&s_qn, incl);
                wbase = __shfl_sync(0xffffffffu, wbase, 31);
                int slot = wbase + incl - cnt;
                while (word) {
                    const int bit = __ffs((int)word) - 1;
                    word &= word - 1;
                    if (slot < FS_QCAP) s_q[slot] = (uint16_t)(base_off + (7 - (bit & 7)) * FS_TP + (bit >> 3));
                    slot++;
                }
            }
            __syncthreads();
            const bool ovf = s_qn > FS_QCAP;                            // uniform
            const int qn = ovf ? 0 : s_qn;
            // ---- exact score (ROI coords = detection coords + 3): queue entries, or every pixel if the queue overflowed ----
            if (!ovf) {
                for (int i = threadIdx.x; i < qn; i += FS_THREADS) {
                    const int off = s_q[i];
                    const int s = fast_score_packed(s_img + off);
                    const int tr = off / FS_TP, tc = off - tr * FS_TP - ax;
                    s_sc[(tr - 2) * FS_SP + (tc - 2)] = (uint8_t)(s > th ? s - 1 : 0);
                }
            } else {
                for (int i = threadIdx.x; i < dw * dh; i += FS_THREADS) {
                    const int r = i / dw, c = i - r * dw;
                    const int s = fast_score_packed(s_img + (r + 3) * FS_TP + ax + c + 3);
                    s_sc[(r + 1) * FS_SP + (c + 1)] = (uint8_t)(s > th ? s - 1 : 0);
                }
            }
            __syncthreads();
            // ---- strict 3x3 NMS inside each cell ----
            const int nitems = ovf ? dw * dh : qn;
            for (int i = threadIdx.x; i < nitems; i += FS_THREADS) {
                int r, c;
                if (ovf) { r = i / dw; c = i - r * dw; }
                else { const int off = s_q[i]; const int tr = off / FS_TP; r = tr - 3; c = off - tr * FS_TP - ax - 3; }
                const uint8_t *q = &s_sc[(r + 1) * FS_SP + (c + 1)];
                const int s = q[0];
                if (s == 0) continue;
                const int cell = min(__float2int_rd(((float)c + 0.5f) * invc), T.ncell - 1);
                if (pass == 1 && !((redo_mask >> cell) & 1u)) continue;   // (overflow scan only: cells that already have keypoints)
                const int c_lo = cell * wcell, c_hi = min(c_lo + wcell, dw);
                bool ok = s > q[-FS_SP] && s > q[FS_SP];
                if (c > c_lo) ok = ok && s > q[-1] && s > q[-FS_SP - 1] && s > q[FS_SP - 1];
                if (c < c_hi - 1) ok = ok && s > q[1] && s > q[-FS_SP + 1] && s > q[FS_SP + 1];
                if (ok) {
                    if (pass == 0) atomicAdd(&s_ccnt[cell], 1);
                    // box-relative coordinates: kp.pt + (j*wCell, i*hCell) — ORBextractor.cpp:865-866
                    const uint32_t val = orbx_pack(T.cj0 * wcell + c + 3, T.ci * T.hcell + r + 3, s);
                    const int o = atomicAdd(&s_nout, 1);
                    if (o < FS_OUT_CAP) s_out[o] = val;
                    else {                                                    // staging full: straight to the global list
                        const int go = atomicAdd(gcnt, 1);
                        if (go < g.cand_cap) gdst[go] = val; else atomicOr(P.status, ORBX_DS_CAND_OVERFLOW);
                    }
                }
            }
            __syncthreads();
            if (pass == 1) break;
            for (int c = 0; c < T.ncell; c++) if (s_ccnt[c] == 0) { redo_cells |= (uint32_t)c << (4 * nredo); redo_mask |= 1u << c; nredo++; }
            if (nredo == 0) break;
            if (threadIdx.x == 0) s_qn = 0;
            __syncthreads();
        }
        // ---- flush the item's candidates: one global atomic, warp 0 ----
        if (warp == 0) {
            const int n = min(s_nout, FS_OUT_CAP);
            if (n > 0) {
                int base = 0;
                if (lane == 0) base = atomicAdd(gcnt, n);
                base = __shfl_sync(0xffffffffu, base, 0);
                for (int i = lane; i < n; i += 32) if (base + i < g.cand_cap) gdst[base + i] = s_out[i];
                if (lane == 0 && base + n > g.cand_cap) atomicOr(P.status, ORBX_DS_CAND_OVERFLOW);
            }
        }
        __syncthreads();                                                // every reader of tile `buf` and of the queues is done
    }
}

// ---- host: TMA descriptors (driver entry point fetched through the runtime: no libcuda link dependency) ----
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled get_encode()
{
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)p;
    }
    return fn;
}

// level as a 3-D tensor of u32 elements: (row pitch / 4) x rows x frames; box = FS_TPW x (hCell + 6) x 1, zero fill outside
static bool encode_level(CUtensorMap *m, const uint8_t *base, size_t pitch, int rows, size_t fstride, int frames, int box_rows)
{
    static_assert(FS_TPW == ORBX_TMA_BOX_WORDS, "FAST tile pitch = TMA box width");
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc || ((uintptr_t)base & 15) || (pitch & 15) || pitch == 0) return false;
    if (frames <= 1 || fstride < pitch) { frames = 1; fstride = pitch * (size_t)rows; }
    fstride = (fstride + 15) & ~(size_t)15;
    const cuuint64_t dims[3] = { (cuuint64_t)(pitch / 4), (cuuint64_t)rows, (cuuint64_t)frames };
    const cuuint64_t strides[2] = { (cuuint64_t)pitch, (cuuint64_t)fstride };
    const cuuint32_t box[3] = { ORBX_TMA_BOX_WORDS, (cuuint32_t)box_rows, 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// tensor maps of the current geometry: levels >= 1 are fixed per geometry, level 0 follows the caller's frames
int orbx_ensure_tmaps(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride)
{
    const FrameGeom &G = h->geo;
    if (!h->tmap_valid) {
        for (int l = 1; l < G.nlevels; l++)
            if (!encode_level(&h->tmap[l], h->d_pyr + G.lv[l].off, (size_t)G.lv[l].pitch, G.lv[l].h, h->pyr_slab, h->prm.max_batch, G.lv[l].hcell + 6) ||
                !encode_level(&h->tmap_rz[l], h->d_pyr + G.lv[l].off, (size_t)G.lv[l].pitch, G.lv[l].h, h->pyr_slab, h->prm.max_batch, ORBX_RZ_BOX_ROWS)) return -1;
        h->tmap_valid = true; h->tmap_l0 = nullptr;
    }
    if (h->tmap_l0 != l0 || h->tmap_l0_step != l0_step || h->tmap_l0_fstride != l0_fstride || h->tmap_l0_frames < nframes) {
        if (!encode_level(&h->tmap[0], l0, l0_step, G.lv[0].h, l0_fstride, nframes, G.lv[0].hcell + 6) ||
            !encode_level(&h->tmap_rz[0], l0, l0_step, G.lv[0].h, l0_fstride, nframes, ORBX_RZ_BOX_ROWS)) return -1;
        h->tmap_l0 = l0; h->tmap_l0_step = l0_step; h->tmap_l0_fstride = l0_fstride; h->tmap_l0_frames = nframes;
    }
    return 0;
}

int launch_fast(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride)
{
    const FrameGeom &G = h->geo;
    if (G.total_strips <= 0) return 0;
    if (orbx_ensure_tmaps(h, nframes, l0, l0_step, l0_fstride) != 0) return -1;
    LevelMaps M;
    memcpy(M.m, h->tmap, sizeof(M.m));
    FastParams P;
    P.cand = h->d_cand; P.cand_slab = G.cand_entries;
    P.ncand = h->d_ncand;
    P.strips = h->d_strips; P.nstrips = G.total_strips; P.nitems = G.total_strips * nframes;
    P.ini_th = h->prm.ini_th_fast; P.min_th = h->prm.min_th_fast;
    P.status = h->d_status;
    P.tile_rows = G.max_hcell + 6;
    const int tile_bytes = ((P.tile_rows + FS_PADROWS) * FS_TP + 127) & ~127;
    const size_t smem = 128 + FS_NBUF * (size_t)tile_bytes + (size_t)(P.tile_rows - 4) * FS_SP;
    static size_t configured = 0;
    if (smem > configured || h->fast_grid_cap <= 0) {
        cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = std::max(configured, smem);
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fast_cells, FS_THREADS, smem);
        h->fast_grid_cap = std::max(1, occ) * h->sm_count;
    }
    const int grid = std::min(P.nitems, h->fast_grid_cap);
    ProfScope ps(h, ORBX_K_FAST);
    k_fast_cells<<<grid, FS_THREADS, smem, h->stream>>>(M, P, h->d_geo);
    return 0;
}
