// k_fast.cu — FAST-9/16 with per-cell 3x3 NMS and the two-threshold fallback.
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (reference ORBextractor.cpp:785-872):
// for every ~35-px cell (+6 px overlap) `cv::FAST(roi, kps, iniThFAST, true)` and, iff that returned
// nothing, `cv::FAST(roi, kps, minThFAST, true)` (:826-827, :845-846).  FAST arithmetic: SURVEY App. A.2
//   S = max over the 16 arcs of 9 contiguous ring pixels of min(I(p)-I(ring)) resp. min(I(ring)-I(p));
//   corner iff S > th; score = S - 1; strict 3x3 NMS INSIDE the cell's ROI (scores outside the
//   cell's detection area read as 0).
//
// Design.  One CTA owns a horizontal run of cells of one cell row (a "strip", described by a host-built
// table); all levels of all frames go in one launch.  Cells are independent in the reference (NMS and the
// threshold retry are per cell ROI), so after the shared staging each WARP owns whole cells and the only block
// barriers are the ones around the staging and the block-wide sweep.
//   1. the strip's ROI is staged in shared memory with aligned 128-bit loads;
//   2. packed sweep, block-wide: a work item is one aligned 32-bit word of the tile (4 adjacent pixels) x 7
//      rows, walked with a 7-row register window (3 LDS.32 + 4 PRMT per row of 4 pixels).  Per row it
//      evaluates a polarity-agnostic pre-test on the four opposite ring pairs (0,8) (4,12) (2,10) (6,14):
//      |I(ring) - I(p)| for 4 pixels is ONE VABSDIFF4.U8; "some member of the pair differs by more than T"
//      with T = 2^k - 1 <= th is an OR, a mask and one add per pair (SWAR, no per-byte compares).  Every
//      9-arc contains a member of each opposite pair, so the test is an exact NECESSARY condition for
//      S > th; it is loose by design (T <= th, sign ignored) and passes ~5 % of the pixels.  The 28 flags of an
//      item are one word of a shared bitmap;
//   3. per cell (one warp): the cell's flag words are compacted into the warp's queue (shuffle prefix sums) and
//      scored exactly: the 16-arc min/max network runs on packed u16x2 lanes (VIMNMX3.U16x2): low half = ring
//      value, high half = 255 - ring value, so one instruction serves the darker and the brighter polarity;
//   4. strict 3x3 NMS over the queue entries; neighbours in another cell are never read (they count as 0);
//   5. a cell with no keypoint at iniThFAST is swept again at minThFAST by its own warp (the reference's retry)
//      — only that cell, not the strip.
// Survivors are appended to the (frame, level) candidate list with one global atomic per CTA;
// list order is arbitrary (the quadtree kernel is order-independent).
//
// Toolchain note: an earlier formulation on signed differences (d = I(p) - I(ring), score via
// max(mn9, -mx9)) produced wrong results on sm_100a with nvcc 12.9 (the negation feeding a fused
// 3-input VIMNMX3 was lost); the raw-value formulation below has no negated min/max operands.
#include "orbx_internal.h"

#define FS_THREADS 192
#define FS_WARPS (FS_THREADS / 32)
#define FS_TP 288                // tile pitch: >= 15 (alignment) + ORBX_FAST_MAX_W + 6 + 8, multiple of 16
#define FS_TPW (FS_TP / 4)
#define FS_PAD 16                // bytes in front of the tile: the word left of tile column 0 is addressable
#define FS_PADROWS 16            // rows behind the tile: a 7-row item may start on the last detection row
#define FS_SP 272                // score-map pitch: >= detection width + 2, multiple of 16
#define FS_WQ 1024               // per-warp survivor queue (u16 tile offsets) >= 32 lanes x 28 flags
#define FS_OUT_CAP 1024          // staged outputs; beyond it survivors are written straight to the global list
#define FS_MAXSEG 10             // 7-row segments per strip: hCell <= 69
#define FS_GROUPS 64             // aligned 4-pixel words per strip row (ORBX_FAST_MAX_W / 4 + alignment)

struct FastParams {
    const uint8_t *l0; size_t l0_step, l0_fstride;
    const uint8_t *pyr; size_t pyr_slab;
    uint32_t *cand; size_t cand_slab;
    int32_t *ncand;
    const uint32_t *strips;      // level:4 | cells:4 | cell row:12 | first cell column:12
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

__global__ void __launch_bounds__(FS_THREADS) k_fast_cells(FastParams P, const FrameGeom *__restrict__ G)
{
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ uint16_t s_wq[FS_WARPS][FS_WQ];
    __shared__ uint32_t s_flag[FS_MAXSEG * FS_GROUPS];
    __shared__ uint32_t s_out[FS_OUT_CAP];
    __shared__ int s_nout, s_base;
    uint8_t *s_img = s_dyn + FS_PAD;                                               // (tile_rows + FS_PADROWS) x FS_TP
    uint8_t *s_sc = s_dyn + FS_PAD + (P.tile_rows + FS_PADROWS) * FS_TP;           // (tile_rows - 4) x FS_SP score map with a zero ring

    const int f = blockIdx.y;
    const uint32_t sd = __ldg(P.strips + blockIdx.x);
    const int level = (int)(sd & 15u), ci = (int)((sd >> 8) & 0xFFFu), cj0 = (int)(sd >> 20);
    int ncell = (int)((sd >> 4) & 15u);
    const int nl = G->nlevels;
    const LevelGeom &g = G->lv[level];
    const int wcell = g.wcell;
    // strip ROI in image coordinates — ORBextractor.cpp:805-822 (cells cj0 .. cj0+ncell-1 of cell row ci)
    const int maxBX = g.w - ORBX_BORDER, maxBY = g.h - ORBX_BORDER;
    const int iniX = ORBX_BORDER + cj0 * wcell, iniY = ORBX_BORDER + ci * g.hcell;
    if (iniY >= maxBY - 3 || iniX >= maxBX - 6) return;
    // cells whose iniX >= maxBorderX-6 are skipped by the reference (:815-816)
    while (ncell > 0 && ORBX_BORDER + (cj0 + ncell - 1) * wcell >= maxBX - 6) ncell--;
    if (ncell <= 0) return;
    const int maxX = min(iniX + ncell * wcell + 6, maxBX), maxY = min(iniY + g.hcell + 6, maxBY);
    const int rw = maxX - iniX, rh = maxY - iniY;
    const int dw = rw - 6, dh = rh - 6;            // detection area of the strip, ROI-relative origin (3,3)
    if (dw <= 0 || dh <= 0) return;

    const uint8_t *src; size_t step;
    if (level == 0) { src = P.l0 + (size_t)f * P.l0_fstride; step = P.l0_step; }
    else { src = P.pyr + (size_t)f * P.pyr_slab + g.off; step = (size_t)g.pitch; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // stage the ROI with aligned 128-bit loads (lane = 16-byte column, warp = row); tile column `ax` is image column iniX
    const int ax = iniX & 15;
    const int vecs = (ax + rw + 15) >> 4;          // <= 18
    if (lane < vecs) {
        const uint8_t *gp = src + (size_t)(iniY + warp) * step + (iniX - ax) + lane * 16;
        uint8_t *sp = s_img + warp * FS_TP + lane * 16;
        for (int r = warp; r < rh; r += FS_WARPS, gp += (size_t)FS_WARPS * step, sp += FS_WARPS * FS_TP)
            *reinterpret_cast<uint4 *>(sp) = __ldg(reinterpret_cast<const uint4 *>(gp));
    }
    // zero the score map (1-px ring included)
    for (int i = threadIdx.x; i < ((dh + 2) * FS_SP) / 16; i += FS_THREADS) reinterpret_cast<uint4 *>(s_sc)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) s_nout = 0;
    __syncthreads();

    // ---- block-wide packed sweep at iniThFAST: one flag word per (7-row segment, tile word) ----
    const int w0 = (ax + 3) >> 2;                                  // tile word holding detection column 0
    const int nGs = ((ax + 3 + dw - 1) >> 2) - w0 + 1;             // words holding detection columns (<= FS_GROUPS)
    const int nseg = (dh + 6) / 7;
    const uint32_t *words = reinterpret_cast<const uint32_t *>(s_img);
    uint32_t HM, KK;
    fast_masks(P.ini_th, HM, KK);
    {
        const float inv = 1.0f / (float)nGs;
        for (int it = threadIdx.x; it < nGs * nseg; it += FS_THREADS) {
            const int seg = __float2int_rd(((float)it + 0.5f) * inv), gidx = it - seg * nGs;
            s_flag[seg * FS_GROUPS + gidx] = fast_sweep7(words + (7 * seg) * FS_TPW + w0 + gidx, HM, KK);
        }
    }
    __syncthreads();

    uint32_t *gdst = P.cand + (size_t)f * P.cand_slab + g.cand_off;
    int32_t *gcnt = &P.ncand[f * nl + level];
    uint16_t *wq = s_wq[warp];

    // ---- one warp per cell: compaction, exact score, NMS, retry ----
    for (int cell = warp; cell < ncell; cell += FS_WARPS) {
        const int c_lo = cell * wcell, c_hi = min(c_lo + wcell, dw);          // detection columns of the cell
        const int ga = ((ax + 3 + c_lo) >> 2) - w0, nG = ((ax + 3 + c_hi - 1) >> 2) - w0 - ga + 1;
        const int items = nG * nseg;
        const float inv = 1.0f / (float)nG;
        for (int pass = 0; pass < 2; pass++) {
            const int th = pass == 0 ? P.ini_th : P.min_th;
            int qn = 0;
            bool ovf = false;
            for (int it0 = 0; it0 < items; it0 += 32) {
                const int it = it0 + lane;
                uint32_t word = 0u;
                int base_off = 0;
                if (it < items) {
                    const int seg = __float2int_rd(((float)it + 0.5f) * inv), gidx = ga + it - seg * nG;
                    const uint32_t raw = pass == 0 ? s_flag[seg * FS_GROUPS + gidx]
                                                   : fast_sweep7(words + (7 * seg) * FS_TPW + w0 + gidx, HM, KK);
                    // columns of this word inside the cell, rows of this segment inside the strip
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
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (qn + total > FS_WQ) {                                     // queue full: score what is queued, NMS will scan the cell
                    __syncwarp();
                    for (int i = lane; i < qn; i += 32) {
                        const int off = wq[i];
                        const int s = fast_score_packed(s_img + off);
                        const int tr = off / FS_TP, tc = off - tr * FS_TP - ax;
                        s_sc[(tr - 2) * FS_SP + (tc - 2)] = (uint8_t)(s > th ? s - 1 : 0);
                    }
                    __syncwarp();
                    qn = 0; ovf = true;
                }
                int slot = qn + incl - cnt;
                while (word) {
                    const int bit = __ffs((int)word) - 1;
                    word &= word - 1;
                    wq[slot++] = (uint16_t)(base_off + (7 - (bit & 7)) * FS_TP + (bit >> 3));
                }
                qn += total;
            }
            __syncwarp();
            // exact scoring of the queued survivors (ROI coords = detection coords + 3)
            for (int i = lane; i < qn; i += 32) {
                const int off = wq[i];
                const int s = fast_score_packed(s_img + off);
                const int tr = off / FS_TP, tc = off - tr * FS_TP - ax;
                s_sc[(tr - 2) * FS_SP + (tc - 2)] = (uint8_t)(s > th ? s - 1 : 0);
            }
            __syncwarp();
            // strict 3x3 NMS inside the cell: queue entries, or every pixel of the cell if the queue overflowed
            const int cw = c_hi - c_lo;
            const int nitems = ovf ? cw * dh : qn;
            const float invw = 1.0f / (float)cw;
            int found = 0;
            for (int i0 = 0; i0 < nitems; i0 += 32) {
                const int i = i0 + lane;
                bool ok = false;
                int r = 0, c = 0, s = 0;
                if (i < nitems) {
                    if (ovf) { r = __float2int_rd(((float)i + 0.5f) * invw); c = c_lo + i - r * cw; }
                    else { const int off = wq[i]; const int tr = off / FS_TP; r = tr - 3; c = off - tr * FS_TP - ax - 3; }
                    const uint8_t *q = &s_sc[(r + 1) * FS_SP + (c + 1)];
                    s = q[0];
                    if (s != 0) {
                        ok = s > q[-FS_SP] && s > q[FS_SP];
                        if (c > c_lo) ok = ok && s > q[-1] && s > q[-FS_SP - 1] && s > q[FS_SP - 1];
                        if (c < c_hi - 1) ok = ok && s > q[1] && s > q[-FS_SP + 1] && s > q[FS_SP + 1];
                    }
                }
                found += __popc(__ballot_sync(0xffffffffu, ok));
                if (ok) {
                    // box-relative coordinates: kp.pt + (j*wCell, i*hCell) — ORBextractor.cpp:865-866
                    const uint32_t val = orbx_pack(cj0 * wcell + c + 3, ci * g.hcell + r + 3, s);
                    const int o = atomicAdd(&s_nout, 1);
                    if (o < FS_OUT_CAP) s_out[o] = val;
                    else {                                                    // staging full: straight to the global list
                        const int go = atomicAdd(gcnt, 1);
                        if (go < g.cand_cap) gdst[go] = val; else atomicOr(P.status, ORBX_DS_CAND_OVERFLOW);
                    }
                }
            }
            // the reference retries a cell at minThFAST only if iniThFAST produced nothing (:843-846)
            if (found > 0 || pass == 1) break;
            fast_masks(P.min_th, HM, KK);
            __syncwarp();
        }
        fast_masks(P.ini_th, HM, KK);
    }
    __syncthreads();
    const int n = min(s_nout, FS_OUT_CAP);
    if (n == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(gcnt, n);
    __syncthreads();
    const int base = s_base;
    for (int i = threadIdx.x; i < n; i += FS_THREADS) {
        if (base + i < g.cand_cap) gdst[base + i] = s_out[i];
    }
    if (threadIdx.x == 0 && base + n > g.cand_cap) atomicOr(P.status, ORBX_DS_CAND_OVERFLOW);
}

void launch_fast(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride)
{
    FastParams P;
    P.l0 = l0; P.l0_step = l0_step; P.l0_fstride = l0_fstride;
    P.pyr = h->d_pyr; P.pyr_slab = h->pyr_slab;
    P.cand = h->d_cand; P.cand_slab = h->geo.cand_entries;
    P.ncand = h->d_ncand;
    P.strips = h->d_strips;
    P.ini_th = h->prm.ini_th_fast; P.min_th = h->prm.min_th_fast;
    P.status = h->d_status;
    P.tile_rows = h->geo.max_hcell + 6;
    const size_t smem = FS_PAD + (size_t)(P.tile_rows + FS_PADROWS) * FS_TP + (size_t)(P.tile_rows - 4) * FS_SP;
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    dim3 grid(h->geo.total_strips, nframes);
    ProfScope ps(h, ORBX_K_FAST);
    k_fast_cells<<<grid, FS_THREADS, smem, h->stream>>>(P, h->d_geo);
}
