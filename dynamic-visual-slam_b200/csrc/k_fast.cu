// k_fast.cu — FAST-9/16 with per-cell 3x3 NMS and the two-threshold fallback.
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (reference ORBextractor.cpp:785-872):
// for every ~35-px cell (+6 px overlap) `cv::FAST(roi, kps, iniThFAST, true)` and, iff that returned
// nothing, `cv::FAST(roi, kps, minThFAST, true)` (:826-827, :845-846).  FAST arithmetic: SURVEY App. A.2
//   S = max over the 16 arcs of 9 contiguous ring pixels of min(I(p)-I(ring)) resp. min(I(ring)-I(p));
//   corner iff S > th; score = S - 1; strict 3x3 NMS INSIDE the cell's ROI (scores outside the
//   cell's detection area read as 0).
//
// Design.  One CTA owns a horizontal run of cells of one cell row (a "strip"); all levels of all
// frames go in one launch.
//   1. the strip's ROI is staged in shared memory with aligned 128-bit loads;
//   2. packed column sweep: a thread owns one aligned 32-bit word of the tile (4 adjacent pixels) and walks
//      down a band of rows with a 7-row register window (3 LDS.32 + 4 PRMT per row of 4 pixels).  Per row it
//      evaluates a polarity-agnostic pre-test on the four opposite ring pairs (0,8) (4,12) (2,10) (6,14):
//      |I(ring) - I(p)| for 4 pixels is ONE VABSDIFF4.U8; "some member of the pair differs by more than T"
//      with T = 2^k - 1 <= th is an OR, a mask and one add per pair (SWAR, no per-byte compares).  Every
//      9-arc contains a member of each opposite pair, so the test is an exact NECESSARY condition for
//      S > th; it is loose by design (T <= th, sign ignored) and passes ~5 % of the pixels;
//   3. survivor flags are kept bit-packed per thread (7 rows x 4 pixels per register), compacted once into a
//      shared queue and scored exactly: the 16-arc min/max network runs on packed u16x2 lanes
//      (VIMNMX3.U16x2): low half = ring value, high half = 255 - ring value, so one instruction serves
//      the darker and the brighter polarity;
//   4. NMS touches queue entries only; neighbours in another cell of the strip are masked to 0;
//   5. cells with no survivor at iniThFAST are swept again at minThFAST (the reference's retry).
// Survivors are appended to the (frame, level) candidate list with one global atomic per CTA;
// list order is arbitrary (the quadtree kernel is order-independent).
//
// Toolchain note: an earlier formulation on signed differences (d = I(p) - I(ring), score via
// max(mn9, -mx9)) produced wrong results on sm_100a with nvcc 12.9 (the negation feeding a fused
// 3-input VIMNMX3 was lost); the raw-value formulation below has no negated min/max operands.
#include "orbx_internal.h"

#define FS_THREADS 256
#define FS_GROUPS 64             // 4-pixel column groups per row band (FS_THREADS = FS_GROUPS x FS_BANDS)
#define FS_BANDS 4
#define FS_TP 288                // tile pitch: >= 15 (alignment) + ORBX_FAST_MAX_W + 6 + 8, multiple of 16
#define FS_TPW (FS_TP / 4)
#define FS_PAD 16                // bytes in front of the tile: the word left of tile column 0 is addressable
#define FS_PADROWS 16            // rows behind the tile: the unrolled sweep may overrun a band by < 16 rows
#define FS_SP 272                // score-map pitch: >= detection width + 2, multiple of 16
#define FS_QCAP 6144             // survivor queue entries (u16 tile offsets); beyond it survivors are scored inline
#define FS_OUT_CAP 1024          // staged outputs; beyond it survivors are written straight to the global list
#define FS_MAX_CELLS 8
#define FS_MAX_GROUPS7 3         // 7-row flag registers per thread: bands of up to 21 rows (hCell <= 69 => <= 18)

struct FastParams {
    const uint8_t *l0; size_t l0_step, l0_fstride;
    const uint8_t *pyr; size_t pyr_slab;
    uint32_t *cand; size_t cand_slab;
    int32_t *ncand;
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
__device__ __forceinline__ FastRow fast_row(const uint32_t *colw, int trow)
{
    const uint32_t *q = colw + trow * FS_TPW;
    const uint32_t L = q[-1], C = q[0], R = q[1];
    FastRow w;
    w.C = C;
    w.P2 = __byte_perm(C, R, 0x5432);      // columns x+2 .. x+5
    w.M2 = __byte_perm(L, C, 0x5432);      // columns x-2 .. x+1
    w.P3 = __byte_perm(C, R, 0x6543);      // columns x+3 .. x+6
    w.M3 = __byte_perm(L, C, 0x4321);      // columns x-3 .. x
    return w;
}

__global__ void __launch_bounds__(FS_THREADS) k_fast_cells(FastParams P, const FrameGeom *__restrict__ G)
{
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ uint16_t s_q[FS_QCAP];
    __shared__ uint32_t s_out[FS_OUT_CAP];
    __shared__ int s_qn, s_nout, s_base, s_redo, s_ovf;
    __shared__ int s_wsum[FS_THREADS / 32];
    __shared__ int s_ccnt[FS_MAX_CELLS];
    __shared__ uint8_t s_col2cell[ORBX_FAST_MAX_W];
    uint8_t *s_img = s_dyn + FS_PAD;                                               // (tile_rows + FS_PADROWS) x FS_TP
    uint8_t *s_sc = s_dyn + FS_PAD + (P.tile_rows + FS_PADROWS) * FS_TP;           // (tile_rows - 4) x FS_SP score map with a zero ring

    const int f = blockIdx.y;
    int level = 0;
    const int nl = G->nlevels;
    for (int l = 1; l < nl; l++) if ((int)blockIdx.x >= G->lv[l].strip_first) level = l;
    const LevelGeom &g = G->lv[level];
    const int sidx = blockIdx.x - g.strip_first;
    const int ci = sidx / g.strips_per_row, sj = sidx - ci * g.strips_per_row;
    const int cj0 = sj * g.cells_per_strip;
    // strip ROI in image coordinates — ORBextractor.cpp:805-822 (cells cj0 .. cj0+ncell-1 of cell row ci)
    const int maxBX = g.w - ORBX_BORDER, maxBY = g.h - ORBX_BORDER;
    const int iniX = ORBX_BORDER + cj0 * g.wcell, iniY = ORBX_BORDER + ci * g.hcell;
    if (iniY >= maxBY - 3 || iniX >= maxBX - 6) return;
    int ncell = min(g.cells_per_strip, g.ncols - cj0);
    // cells whose iniX >= maxBorderX-6 are skipped by the reference (:815-816)
    while (ncell > 0 && ORBX_BORDER + (cj0 + ncell - 1) * g.wcell >= maxBX - 6) ncell--;
    if (ncell <= 0) return;
    const int maxX = min(iniX + ncell * g.wcell + 6, maxBX), maxY = min(iniY + g.hcell + 6, maxBY);
    const int rw = maxX - iniX, rh = maxY - iniY;
    const int dw = rw - 6, dh = rh - 6;            // detection area of the strip, ROI-relative origin (3,3)
    if (dw <= 0 || dh <= 0) return;

    const uint8_t *src; size_t step;
    if (level == 0) { src = P.l0 + (size_t)f * P.l0_fstride; step = P.l0_step; }
    else { src = P.pyr + (size_t)f * P.pyr_slab + g.off; step = (size_t)g.pitch; }
    // stage the ROI with aligned 128-bit loads; tile column `ax` is image column iniX
    const int ax = iniX & 15;
    const int vecs = (ax + rw + 15) >> 4;
    src += (size_t)iniY * step + (iniX - ax);
    for (int i = threadIdx.x; i < rh * 32; i += FS_THREADS) {
        const int r = i >> 5, vi = i & 31;
        if (vi < vecs) reinterpret_cast<uint4 *>(s_img + r * FS_TP)[vi] = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)r * step) + vi);
    }
    for (int c = threadIdx.x; c < dw; c += FS_THREADS) s_col2cell[c] = (uint8_t)min(c / g.wcell, ncell - 1);
    if (threadIdx.x < FS_MAX_CELLS) s_ccnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_nout = 0; s_redo = 0; }
    const int lane = threadIdx.x & 31;
    uint32_t *gdst = P.cand + (size_t)f * P.cand_slab + g.cand_off;
    int32_t *gcnt = &P.ncand[f * nl + level];

    // sweep geometry: thread = tile word wi (4 pixels) x row band
    const int grp = threadIdx.x & (FS_GROUPS - 1), band = threadIdx.x / FS_GROUPS;
    const int wi = ((ax + 3) >> 2) + grp;                         // first word holding a detection column + group
    const int cbase = 4 * wi - (ax + 3);                          // detection column of byte 0 of the word (-3 .. )
    const uint32_t *colw = reinterpret_cast<const uint32_t *>(s_img) + wi;
    const int RB = (dh + FS_BANDS - 1) / FS_BANDS;                // rows per band (<= 18 for hCell <= 69)
    const int r_begin = band * RB, r_end = min(dh, r_begin + RB);

    for (int pass = 0; pass < 2; pass++) {
        const int th = pass == 0 ? P.ini_th : P.min_th;
        // loose pre-test threshold T = 2^sh - 1 <= th: |d| > T  <=>  (|d| & HM) != 0
        const int sh = min(7, 31 - __clz(th + 1));
        const uint32_t HM = ((0xFFu << sh) & 0xFFu) * 0x01010101u;
        const uint32_t KK = (0x80u - (1u << sh)) * 0x01010101u;   // t + KK sets bit 7 of every byte with t >= 2^sh
        // zero the score map (1-px ring included) and the queue
        for (int i = threadIdx.x; i < ((dh + 2) * FS_SP) / 16; i += FS_THREADS) reinterpret_cast<uint4 *>(s_sc)[i] = make_uint4(0, 0, 0, 0);
        if (threadIdx.x == 0) { s_qn = 0; s_ovf = 0; }
        __syncthreads();
        // columns this thread may report: inside the detection area and (retry pass) in a cell that is still empty
        uint32_t vm = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = cbase + j;
            if (c >= 0 && c < dw && (pass == 0 || s_ccnt[s_col2cell[c]] == 0)) vm |= 0xFEu << (8 * j);
        }
        // ---- packed sweep: flags[gi] bit (7-k) of byte j = pixel (row r_begin + 7*gi + k, column cbase + j) survives ----
        uint32_t flags[FS_MAX_GROUPS7] = { 0u, 0u, 0u };
        if (vm != 0u && r_begin < r_end) {
            FastRow w[7];
#pragma unroll
            for (int k = 0; k < 6; k++) w[k] = fast_row(colw, r_begin + k);
#pragma unroll
            for (int gi = 0; gi < FS_MAX_GROUPS7; gi++) {
                const int r0 = r_begin + 7 * gi;
                if (r0 < r_end) {
                    uint32_t fl = 0u;
#pragma unroll
                    for (int k = 0; k < 7; k++) {
                        w[(k + 6) % 7] = fast_row(colw, r0 + k + 6);           // ring row dy = +3 of detection row r0 + k
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
                    const int nv = min(7, r_end - r0);                         // rows of this group inside the band
                    flags[gi] = fl & vm & (((0xFF00u >> nv) & 0xFFu) * 0x01010101u);
                }
            }
        }
        // block-wide compaction of the flags into the queue
        const int cnt = __popc(flags[0]) + __popc(flags[1]) + __popc(flags[2]);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) s_wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int wv = 0; wv < FS_THREADS / 32; wv++) { const int t = s_wsum[wv]; s_wsum[wv] = run; run += t; }
            s_qn = run;
        }
        __syncthreads();
        int slot = s_wsum[threadIdx.x >> 5] + incl - cnt;
#pragma unroll
        for (int gi = 0; gi < FS_MAX_GROUPS7; gi++) {
            uint32_t m = flags[gi];
            while (m) {
                const int bit = __ffs((int)m) - 1;
                m &= m - 1;
                const int r = r_begin + 7 * gi + 7 - (bit & 7), tcol = 4 * wi + (bit >> 3);   // tile byte column
                if (slot < FS_QCAP) s_q[slot] = (uint16_t)((r + 3) * FS_TP + tcol);
                else {                                                // queue full: score inline, NMS will scan the map
                    const int s = fast_score_packed(s_img + (r + 3) * FS_TP + tcol);
                    s_sc[(r + 1) * FS_SP + (tcol - ax - 3 + 1)] = (uint8_t)(s > th ? s - 1 : 0);
                    s_ovf = 1;
                }
                slot++;
            }
        }
        __syncthreads();
        const int qn = min(s_qn, FS_QCAP);
        // ---- exact scoring of the queued survivors ----
        for (int i = threadIdx.x; i < qn; i += FS_THREADS) {
            const int off = s_q[i];
            const int s = fast_score_packed(s_img + off);
            const int tr = off / FS_TP, tc = off - tr * FS_TP - ax;    // ROI coords = detection coords + 3
            s_sc[(tr - 2) * FS_SP + (tc - 2)] = (uint8_t)(s > th ? s - 1 : 0);
        }
        __syncthreads();
        // ---- strict 3x3 NMS inside each cell: queue entries, or the whole map if the queue overflowed ----
        const bool scan_all = s_ovf != 0;
        const int nitems = scan_all ? dw * dh : qn;
        for (int i = threadIdx.x; i < nitems; i += FS_THREADS) {
            int r, c;
            if (scan_all) { r = i / dw; c = i - r * dw; }
            else { const int off = s_q[i]; const int tr = off / FS_TP; r = tr - 3; c = off - tr * FS_TP - ax - 3; }
            const uint8_t *q = &s_sc[(r + 1) * FS_SP + (c + 1)];
            const int s = q[0];
            if (s == 0) continue;
            const int cell = s_col2cell[c];
            const int cl = c - cell * g.wcell;                          // column inside the cell's detection area
            const bool hasL = cl > 0, hasR = (cl < g.wcell - 1) && (c < dw - 1);
            bool ok = s > q[-FS_SP] && s > q[FS_SP];
            if (hasL) ok = ok && s > q[-1] && s > q[-FS_SP - 1] && s > q[FS_SP - 1];
            if (hasR) ok = ok && s > q[1] && s > q[-FS_SP + 1] && s > q[FS_SP + 1];
            if (ok) {
                atomicAdd(&s_ccnt[cell], 1);
                // box-relative coordinates: kp.pt + (j*wCell, i*hCell) — ORBextractor.cpp:865-866
                const uint32_t val = orbx_pack(cj0 * g.wcell + c + 3, ci * g.hcell + r + 3, s);
                const int o = atomicAdd(&s_nout, 1);
                if (o < FS_OUT_CAP) s_out[o] = val;
                else {                                                    // staging full: straight to the global list
                    const int go = atomicAdd(gcnt, 1);
                    if (go < g.cand_cap) gdst[go] = val; else atomicOr(P.status, ORBX_DS_CAND_OVERFLOW);
                }
            }
        }
        __syncthreads();
        // fallback pass only for cells that produced nothing (uniform decision)
        if (pass == 0) {
            if (threadIdx.x < ncell && s_ccnt[threadIdx.x] == 0) s_redo = 1;
            __syncthreads();
            if (!s_redo) break;
        }
    }
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
