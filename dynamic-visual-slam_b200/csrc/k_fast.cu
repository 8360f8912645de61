// k_fast.cu — FAST-9/16 with per-cell 3x3 NMS and the two-threshold fallback.
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree (reference ORBextractor.cpp:785-872):
// for every ~35-px cell (+6 px overlap) `cv::FAST(roi, kps, iniThFAST, true)` and, iff that returned
// nothing, `cv::FAST(roi, kps, minThFAST, true)` (:826-827, :845-846).  FAST arithmetic: SURVEY App. A.2
//   S = max over the 16 arcs of 9 contiguous ring pixels of min(I(p)-I(ring)) resp. min(I(ring)-I(p));
//   corner iff S > th; score = S - 1; strict 3x3 NMS INSIDE the cell's ROI (scores outside the
//   cell's detection area read as 0).
//
// Design.  The reference's unit of work is the CELL (FAST, NMS and the threshold retry all run on one cell ROI), and here a
// cell is owned by ONE WARP from start to finish: there is no block barrier anywhere in the kernel.
//   * persistent one-warp CTAs loop over the (frame, level, cell) items of the whole batch (host-built cell table);
//   * the cell's ROI (+3 px ring margin, 16-byte aligned start) arrives in the warp's shared memory by TMA
//     (cp.async.bulk.tensor.3d, one descriptor per pyramid level, frames as the third tensor dimension); the next
//     cell's window is prefetched into L2 (cp.async.bulk.prefetch.tensor) while the current one is processed;
//   1. packed sweep: a sweep unit is one aligned 32-bit word of the tile (4 adjacent pixels) x 7 rows, walked with a
//      7-row register window (3 LDS.32 + 4 PRMT per row of 4 pixels); lanes take the cell's units round-robin.  Per row it
//      evaluates a polarity-agnostic pre-test on the four opposite ring pairs (0,8) (4,12) (2,10) (6,14):
//      |I(ring) - I(p)| for 4 pixels is ONE VABSDIFF4.U8; "some member of the pair differs by more than T"
//      with T = 2^k - 1 <= th is an OR, a mask and one add per pair (SWAR, no per-byte compares).  Every
//      9-arc contains a member of each opposite pair, so the test is an exact NECESSARY condition for
//      S > th; it is loose by design (T <= th, sign ignored) and passes ~5 % of the pixels.  Survivors are
//      compacted into the warp's queue with shuffle prefix sums;
//   2. exact score of the queued survivors, dense over the lanes: the 16-arc min/max network runs on packed
//      u16x2 lanes (VIMNMX3.U16x2): low half = ring value, high half = 255 - ring value, so one instruction
//      serves the darker and the brighter polarity;
//   3. strict 3x3 NMS over the queue entries (scores outside the cell count as 0, as in the reference's per-ROI FAST);
//   4. a cell with no keypoint at iniThFAST is swept again at minThFAST (the reference's retry) by the same warp.
// Keypoints are appended to the (frame, level) candidate list with one warp-aggregated global atomic per NMS step;
// list order is arbitrary (the quadtree kernel is order-independent).
//
// Toolchain note: an earlier formulation on signed differences (d = I(p) - I(ring), score via
// max(mn9, -mx9)) produced wrong results on sm_100a with nvcc 12.9 (the negation feeding a fused
// 3-input VIMNMX3 was lost); the raw-value formulation below has no negated min/max operands.
#include "orbx_internal.h"
#include "orbx_tma.h"
#include <algorithm>
#include <cstring>

// tile pitch TP = TMA box width, a template parameter: 80 bytes when every cell ROI (15 alignment + wCell + 6 + 4) fits, else 96
// (wCell <= 69).  A 7-row sweep unit may start on the last detection row: the rows it reads past the tile fall into the score map
// that follows the tile in shared memory (their flags are masked).
#define FS_WQ 512                // survivor queue (u16 tile offsets); a sweep step with more survivors than this is scored in place

struct FastParams {
    uint32_t *cand; size_t cand_slab;
    int32_t *ncand;
    const uint32_t *cells;       // level:4 | cell row:14 | cell column:14
    int ncells, nitems;          // items = ncells x frames
    int ini_th, min_th;
    int32_t *status;
    int tile_rows;               // max (hCell + 6) over the levels
    int map_pitch;               // score-map pitch: >= max cell width + 2, multiple of 16
};

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
template <int TP> __device__ __noinline__ uint32_t fast_sweep7(const uint32_t *q, uint32_t HM, uint32_t KK, int R)
{
    FastRow w[7];
#pragma unroll
    for (int k = 0; k < 6; k++) w[k] = fast_row(q + k * (TP / 4));
    uint32_t fl = 0u;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k >= R) break;                                                       // units of R <= 8 rows (uniform); the window indices are mod 7
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

// geometry of one work item, derived from its cell descriptor
struct FastItem { int f, level, ci, cj, wcell, hcell, iniX, iniY, ax, dw, dh; };
__device__ __forceinline__ FastItem fast_item(const FastParams &P, const FrameGeom *__restrict__ G, int item)
{
    FastItem t;
    t.f = item / P.ncells;
    const uint32_t cd = __ldg(P.cells + (item - t.f * P.ncells));
    t.level = (int)(cd & 15u); t.ci = (int)((cd >> 4) & 0x3FFFu); t.cj = (int)(cd >> 18);
    const LevelGeom &g = G->lv[t.level];
    t.wcell = g.wcell; t.hcell = g.hcell;
    // cell ROI in image coordinates — ORBextractor.cpp:805-822
    t.iniX = ORBX_BORDER + t.cj * t.wcell; t.iniY = ORBX_BORDER + t.ci * t.hcell;
    const int maxX = min(t.iniX + t.wcell + 6, g.w - ORBX_BORDER), maxY = min(t.iniY + t.hcell + 6, g.h - ORBX_BORDER);
    t.dw = maxX - t.iniX - 6; t.dh = maxY - t.iniY - 6;      // detection area, ROI-relative origin (3,3)
    t.ax = t.iniX & 15;                                      // tile byte of ROI column 0 (TMA boxes start on 16-byte columns)
    return t;
}

template <int TP> __global__ void __launch_bounds__(32) k_fast_cells(const __grid_constant__ LevelMaps M, FastParams P, const FrameGeom *__restrict__ G)
{
    extern __shared__ __align__(128) uint8_t s_raw[];
    __shared__ __align__(8) uint64_t s_full;
    uint8_t *s_dyn = s_raw + ((128u - (smem_u32(s_raw) & 127u)) & 127u);
    // layout: [128-byte pad | tile | score map | queue]: the word left of tile column 0 (read, never used) falls into the pad
    uint8_t *s_img = s_dyn + 128;                                                        // tile_rows x TP, 128-byte aligned (TMA destination)
    const int tile_bytes = ((P.tile_rows * TP) + 127) & ~127;
    uint8_t *s_sc = s_img + tile_bytes;                                                  // (tile_rows - 4) x map_pitch, zero ring
    const int SP = P.map_pitch;
    const int map_bytes = (((P.tile_rows - 4) * SP) + 127) & ~127;
    uint16_t *wq = reinterpret_cast<uint16_t *>(s_sc + map_bytes);
    const int lane = threadIdx.x;
    const int nl = G->nlevels;

    if (lane == 0) {
        mbar_init(&s_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    int it_n = 0;
    for (int item = blockIdx.x; item < P.nitems; item += gridDim.x, it_n++) {
        const FastItem T = fast_item(P, G, item);
        const LevelGeom &g = G->lv[T.level];
        const int ax = T.ax, dw = T.dw, dh = T.dh;
        if (lane == 0) {
            // every lane finished reading the previous tile (the __syncwarp that ends the loop body)
            mbar_expect_tx(&s_full, (uint32_t)((T.hcell + 6) * TP));
            tma_load_3d(s_img, &M.m[T.level], (T.iniX & ~15) >> 2, T.iniY, T.f, &s_full);
            if (item + (int)gridDim.x < P.nitems) {                                     // next cell's window -> L2
                const FastItem N = fast_item(P, G, item + gridDim.x);
                tma_prefetch_3d(&M.m[N.level], (N.iniX & ~15) >> 2, N.iniY, N.f);
            }
        }
        // zero the score map (1-px ring included) while the tile lands
        for (int i = lane; i < ((dh + 2) * SP) / 16; i += 32) reinterpret_cast<uint4 *>(s_sc)[i] = make_uint4(0, 0, 0, 0);
        mbar_wait(&s_full, (uint32_t)(it_n & 1));
        __syncwarp();

        const int w0 = (ax + 3) >> 2;                                  // tile word holding detection column 0
        const int nG = ((ax + 3 + dw - 1) >> 2) - w0 + 1;              // words holding detection columns (<= 19)
        // rows per sweep unit: the R in 4..8 that minimises (warp iterations) x (cost of a unit = 6 window rows + R tested rows)
        int R = 8, nseg = (dh + 7) / 8;
        {
            int best = ((nG * nseg + 31) >> 5) * (42 + 30 * 8);
            for (int r = 7; r >= 4; r--) {
                const int ns = (dh + r - 1) / r, cost = ((nG * ns + 31) >> 5) * (42 + 30 * r);
                if (cost < best) { best = cost; R = r; nseg = ns; }
            }
        }
        const int units = nG * nseg;
        const float inv = 1.0f / (float)nG;
        const uint32_t *words = reinterpret_cast<const uint32_t *>(s_img);
        uint32_t *gdst = P.cand + (size_t)T.f * P.cand_slab + g.cand_off;
        int32_t *gcnt = &P.ncand[T.f * nl + T.level];
        for (int pass = 0; pass < 2; pass++) {
            const int th = pass == 0 ? P.ini_th : P.min_th;
            uint32_t HM, KK;
            fast_masks(th, HM, KK);
            int qn = 0;
            bool ovf = false;
            // ---- packed sweep + compaction into the warp's queue ----
            for (int u0 = 0; u0 < units; u0 += 32) {
                const int u = u0 + lane;
                uint32_t word = 0u;
                int base_off = 0;
                if (u < units) {
                    const int seg = __float2int_rd(((float)u + 0.5f) * inv), gidx = u - seg * nG;
                    const uint32_t raw = fast_sweep7<TP>(words + (R * seg) * (TP / 4) + w0 + gidx, HM, KK, R);
                    const int cb = 4 * (w0 + gidx) - (ax + 3);                // detection column of byte 0
                    uint32_t cm = 0u;
#pragma unroll
                    for (int j = 0; j < 4; j++) if (cb + j >= 0 && cb + j < dw) cm |= 0xFFu << (8 * j);
                    const int nv = min(R, dh - R * seg);
                    word = raw & cm & (((0xFF00u >> nv) & 0xFFu) * 0x01010101u);
                    base_off = (R * seg + 3) * TP + 4 * (w0 + gidx);
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
                        const int s = fast_score_packed<TP>(s_img + off);
                        const int tr = off / TP, tc = off - tr * TP - ax;
                        s_sc[(tr - 2) * SP + (tc - 2)] = (uint8_t)(s > th ? s - 1 : 0);
                    }
                    __syncwarp();
                    qn = 0; ovf = true;
                    if (total > FS_WQ) {                                      // this step alone does not fit: score its survivors in place
                        while (word) {
                            const int bit = __ffs((int)word) - 1;
                            word &= word - 1;
                            const int off = base_off + (7 - (bit & 7)) * TP + (bit >> 3);
                            const int s = fast_score_packed<TP>(s_img + off);
                            const int tr = off / TP, tc = off - tr * TP - ax;
                            s_sc[(tr - 2) * SP + (tc - 2)] = (uint8_t)(s > th ? s - 1 : 0);
                        }
                        continue;
                    }
                }
                int slot = qn + incl - cnt;
                while (word) {
                    const int bit = __ffs((int)word) - 1;
                    word &= word - 1;
                    wq[slot++] = (uint16_t)(base_off + (7 - (bit & 7)) * TP + (bit >> 3));
                }
                qn += total;
            }
            __syncwarp();
            // ---- exact score of the queued survivors (ROI coords = detection coords + 3) ----
            for (int i = lane; i < qn; i += 32) {
                const int off = wq[i];
                const int s = fast_score_packed<TP>(s_img + off);
                const int tr = off / TP, tc = off - tr * TP - ax;
                s_sc[(tr - 2) * SP + (tc - 2)] = (uint8_t)(s > th ? s - 1 : 0);
            }
            __syncwarp();
            // ---- strict 3x3 NMS inside the cell: queue entries, or every pixel of the cell if the queue overflowed ----
            const int nitems = ovf ? dw * dh : qn;
            const float invw = 1.0f / (float)dw;
            int found = 0;
            for (int i0 = 0; i0 < nitems; i0 += 32) {
                const int i = i0 + lane;
                bool ok = false;
                int r = 0, c = 0, s = 0;
                if (i < nitems) {
                    if (ovf) { r = __float2int_rd(((float)i + 0.5f) * invw); c = i - r * dw; }
                    else { const int off = wq[i]; const int tr = off / TP; r = tr - 3; c = off - tr * TP - ax - 3; }
                    const uint8_t *q = &s_sc[(r + 1) * SP + (c + 1)];
                    s = q[0];
                    // the zero ring around the map stands for "outside the cell's detection area"
                    if (s != 0) ok = s > q[-SP] && s > q[SP] && s > q[-1] && s > q[-SP - 1] && s > q[SP - 1] && s > q[1] && s > q[-SP + 1] && s > q[SP + 1];
                }
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (bal) {
                    const int nk = __popc(bal);
                    int base = 0;
                    if (lane == 0) base = atomicAdd(gcnt, nk);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (ok) {
                        // box-relative coordinates: kp.pt + (j*wCell, i*hCell) — ORBextractor.cpp:865-866
                        const int go = base + __popc(bal & ((1u << lane) - 1u));
                        if (go < g.cand_cap) gdst[go] = orbx_pack(T.cj * T.wcell + c + 3, T.ci * T.hcell + r + 3, s);
                        else atomicOr(P.status, ORBX_DS_CAND_OVERFLOW);
                    }
                    found += nk;
                }
            }
            // the reference retries a cell at minThFAST iff iniThFAST produced nothing (:843-846)
            if (found > 0 || pass == 1) break;
            __syncwarp();
        }
        __syncwarp();                                                   // every lane is done with the tile, the map and the queue
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
static bool encode_level(CUtensorMap *m, const uint8_t *base, size_t pitch, int rows, size_t fstride, int frames, int box_rows, int box_words = ORBX_TMA_BOX_WORDS)
{
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc || ((uintptr_t)base & 15) || (pitch & 15) || pitch == 0) return false;
    if (frames <= 1 || fstride < pitch) { frames = 1; fstride = pitch * (size_t)rows; }
    fstride = (fstride + 15) & ~(size_t)15;
    const cuuint64_t dims[3] = { (cuuint64_t)(pitch / 4), (cuuint64_t)rows, (cuuint64_t)frames };
    const cuuint64_t strides[2] = { (cuuint64_t)pitch, (cuuint64_t)fstride };
    const cuuint32_t box[3] = { (cuuint32_t)box_words, (cuuint32_t)box_rows, 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// FAST tile pitch (= TMA box width) for the geometry: 15 alignment bytes + the widest cell ROI (wCell + 6) + one spare word
static int fast_tile_pitch(const FrameGeom &G) { return G.max_wcell + 25 <= 80 ? 80 : 96; }

// tensor maps of the current geometry: levels >= 1 are fixed per geometry, level 0 follows the caller's frames
int orbx_ensure_tmaps(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride)
{
    const FrameGeom &G = h->geo;
    if (!h->tmap_valid) {
        for (int l = 1; l < G.nlevels; l++)
            if (!encode_level(&h->tmap[l], h->d_pyr + G.lv[l].off, (size_t)G.lv[l].pitch, G.lv[l].h, h->pyr_slab, h->prm.max_batch, G.lv[l].hcell + 6) ||
                !encode_level(&h->tmap_rz[l], h->d_pyr + G.lv[l].off, (size_t)G.lv[l].pitch, G.lv[l].h, h->pyr_slab, h->prm.max_batch, ORBX_RZ_BOX_ROWS) ||
                !encode_level(&h->tmap_cell[l], h->d_pyr + G.lv[l].off, (size_t)G.lv[l].pitch, G.lv[l].h, h->pyr_slab, h->prm.max_batch, G.lv[l].hcell + 6, fast_tile_pitch(G) / 4)) return -1;
        h->tmap_valid = true; h->tmap_l0 = nullptr;
    }
    if (h->tmap_l0 != l0 || h->tmap_l0_step != l0_step || h->tmap_l0_fstride != l0_fstride || h->tmap_l0_frames < nframes) {
        if (!encode_level(&h->tmap[0], l0, l0_step, G.lv[0].h, l0_fstride, nframes, G.lv[0].hcell + 6) ||
            !encode_level(&h->tmap_rz[0], l0, l0_step, G.lv[0].h, l0_fstride, nframes, ORBX_RZ_BOX_ROWS) ||
            !encode_level(&h->tmap_cell[0], l0, l0_step, G.lv[0].h, l0_fstride, nframes, G.lv[0].hcell + 6, fast_tile_pitch(G) / 4)) return -1;
        h->tmap_l0 = l0; h->tmap_l0_step = l0_step; h->tmap_l0_fstride = l0_fstride; h->tmap_l0_frames = nframes;
    }
    return 0;
}

int launch_fast(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride)
{
    const FrameGeom &G = h->geo;
    if (G.total_cells_valid <= 0) return 0;
    if (orbx_ensure_tmaps(h, nframes, l0, l0_step, l0_fstride) != 0) return -1;
    LevelMaps M;
    memcpy(M.m, h->tmap_cell, sizeof(M.m));
    FastParams P;
    P.cand = h->d_cand; P.cand_slab = G.cand_entries;
    P.ncand = h->d_ncand;
    P.cells = h->d_cells; P.ncells = G.total_cells_valid; P.nitems = G.total_cells_valid * nframes;
    P.ini_th = h->prm.ini_th_fast; P.min_th = h->prm.min_th_fast;
    P.status = h->d_status;
    P.tile_rows = G.max_hcell + 6;
    P.map_pitch = (G.max_wcell + 2 + 15) & ~15;
    const int TP = fast_tile_pitch(G);
    const size_t map_bytes = (size_t)((((P.tile_rows - 4) * P.map_pitch) + 127) & ~127);
    const size_t smem = 128 + 128 + (size_t)(((P.tile_rows * TP) + 127) & ~127) + map_bytes + FS_WQ * 2;
    auto kern = TP == 80 ? k_fast_cells<80> : k_fast_cells<96>;
    if (smem != h->fast_smem || TP != h->fast_tp) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, smem);
        h->fast_grid_cap = std::max(1, occ) * h->sm_count; h->fast_smem = smem; h->fast_tp = TP;
    }
    const int grid = std::min(P.nitems, h->fast_grid_cap);
    ProfScope ps(h, ORBX_K_FAST);
    kern<<<grid, 32, smem, h->stream>>>(M, P, h->d_geo);
    return 0;
}
