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
//   * persistent one-warp CTAs; after its first cell a warp DRAWS its next (frame, level, cell) items from a global counter — cell
//     cost varies by an order of magnitude, a static split left a third of the warps idle at the end;
//   * a cell is described by a host-built 32-byte record (orbx_build_fast_cells): tile coordinates, detection area, sweep-unit
//     shape, output list — two LDG.128 instead of ~250 instructions of address arithmetic per cell;
//   * the cell's ROI (+3 px ring margin, 16-byte aligned start) arrives in the warp's shared memory by TMA
//     (cp.async.bulk.tensor.3d, one descriptor per pyramid level, frames as the third tensor dimension); software pipeline: the
//     draw for the cell after next is in flight during the sweep, the next cell's window is prefetched into L2
//     (cp.async.bulk.prefetch.tensor) and loaded as soon as the tile is free, and a cell's corners go to the global list one
//     cell later, when the counter atomic that reserved their slots has long returned;
//   1. packed sweep: a sweep unit is one aligned 32-bit word of the tile (4 adjacent pixels) x R rows, R <= 16 picked per cell
//      shape on the host so that the cell is as few full warp iterations as possible (a 37-row cell: 30 units of 13 rows = ONE
//      iteration); a unit is walked with a 7-row register window (3 LDS.32 + 4 PRMT per row of 4 pixels).  Per row it
//      evaluates a polarity-agnostic pre-test on the four opposite ring pairs (0,8) (4,12) (2,10) (6,14):
//      |I(ring) - I(p)| for 4 pixels is ONE VABSDIFF4.U8; "some member of the pair differs by more than T"
//      with T = 2^k - 1 <= th is an OR, a mask and one add per pair (SWAR, no per-byte compares).  Every
//      9-arc contains a member of each opposite pair, so the test is an exact NECESSARY condition for
//      S > th; it is loose by design (T <= th, sign ignored) and passes ~5 % of the pixels.  A lane's survivors (two flag
//      words, rows 0-7 and 8-15) take slots of the warp's queue with one shared-memory atomic;
//   2. exact score of the queued survivors, dense over the lanes: the 16-arc min/max network runs on packed
//      u16x2 lanes (VIMNMX3.U16x2): low half = ring value, high half = 255 - ring value, so one instruction
//      serves the darker and the brighter polarity;
//   3. strict, branch-free 3x3 NMS over the queue entries against a zero-ringed score map (scores outside the cell count as 0,
//      as in the reference's per-ROI FAST);
//   4. a cell with no keypoint at iniThFAST is swept again at minThFAST (the reference's retry) by the same warp.
// List order is arbitrary (the quadtree kernel is order-independent).
//
// Toolchain note: an earlier formulation on signed differences (d = I(p) - I(ring), score via
// max(mn9, -mx9)) produced wrong results on sm_100a with nvcc 12.9 (the negation feeding a fused
// 3-input VIMNMX3 was lost); the raw-value formulation below has no negated min/max operands.
#include "orbx_internal.h"
#include "orbx_tma.h"
#include <algorithm>
#include <cstring>
#include <vector>

// tile pitch TP = TMA box width, a template parameter: 80 bytes when every cell ROI (15 alignment + wCell + 6 + 4) fits, else 96
// (wCell <= 69).  The last sweep unit of a word column may reach up to 14 rows past the tile: those reads fall into the score map
// that follows the tile in shared memory (their flags are masked).
#ifndef FS_WQ
#define FS_WQ 512                // survivor queue (u16 tile offsets); a sweep step with more survivors than this is scored in place
#endif

struct FastParams {
    uint32_t *cand; size_t cand_slab;
    int32_t *ncand;
    const uint4 *cells;          // two 16-byte words per cell (orbx_build_fast_cells)
    int ncells, nitems;          // items = ncells x frames
    float inv_ncells;
    int ini_th, min_th;
    int nlevels;
    int32_t *status;
    int32_t *work;               // dynamic work counter (zeroed with the corner counts)
    int tile_rows;               // max (hCell + 6) over the levels
    int map_pitch;               // score-map pitch: >= max cell width + 2, multiple of 16
    const int32_t *items;        // retry mode (k_fast_dense.cu): the (frame, cell) items to run and their device-side count; null = all of them
    const int32_t *nitems_dev;
};

#include "orbx_fast_dev.h"

// the (frame-independent) description of one cell, built on the host with the geometry (orbx_build_fast_cells): two 16-byte loads
// replace ~250 instructions of per-cell address arithmetic
//   c0.x = tile word column (TMA coordinate 0) | first tile row << 16      c0.y = level | ax << 8 | dw << 16 | dh << 24
//   c0.z = R | nseg << 8 | nG << 16 | w0 << 24                             c0.w = kx | ky << 16
//   c1.x = 1 / nG (float bits)   c1.y = offset of the level's corner list   c1.z = its capacity   c1.w = bytes the TMA box delivers
struct FastCellRegs { uint4 c0, c1; };
// item -> (frame, cell): float estimate + exact fix-up (no integer division on the per-cell path)
template <bool RETRY> __device__ __forceinline__ void fast_split(const FastParams &P, int idx, int nitems, int &f, int &c)
{
    int item = idx;
    if (RETRY) item = idx < nitems ? P.items[idx] : 0;          // written by the previous kernel of the stream: plain loads
    f = __float2int_rz(((float)item + 0.5f) * P.inv_ncells);
    c = item - f * P.ncells;
    if (c < 0) { f--; c += P.ncells; }
    else if (c >= P.ncells) { f++; c -= P.ncells; }
}

template <int TP> __device__ __forceinline__ void fast_score_to_map(const uint8_t *s_img, uint8_t *s_sc, int SP, int off, int ax, int th)
{
    const int s = fast_score_packed<TP>(s_img + off);
    const int tr = off / TP, tc = off - tr * TP - ax;                  // ROI coordinates = detection coordinates + 3
    s_sc[(tr - 2) * SP + (tc - 2)] = (uint8_t)(s > th ? s - 1 : 0);
}

template <int TP, bool RETRY> __global__ void __launch_bounds__(32) k_fast_cells(const __grid_constant__ LevelMaps M, FastParams P, const FrameGeom *__restrict__ G)
{
    extern __shared__ __align__(128) uint8_t s_raw[];
    __shared__ __align__(8) uint64_t s_full;
    __shared__ int s_q[2];                                // survivor queue: fill, first slot that did not fit
    __shared__ uint32_t s_res[2 * FS_RES];                // result lists of the current and the previous cell
    ORBX_PDL_ENTRY();
    uint8_t *s_dyn = s_raw + ((128u - (smem_u32(s_raw) & 127u)) & 127u);
    // layout: [128-byte pad | tile | score map | queue]: the word left of tile column 0 (read, never used) falls into the pad
    uint8_t *s_img = s_dyn + 128;                                                        // tile_rows x TP, 128-byte aligned (TMA destination)
    const int tile_bytes = ((P.tile_rows * TP) + 127) & ~127;
    uint8_t *s_sc = s_img + tile_bytes;                                                  // (tile_rows - 4) x map_pitch, zero ring
    const int SP = P.map_pitch;
    const int map_bytes = (((P.tile_rows - 4) * SP) + 127) & ~127;
    uint16_t *wq = reinterpret_cast<uint16_t *>(s_sc + map_bytes);
    const uint32_t *words = reinterpret_cast<const uint32_t *>(s_img);
    const int lane = threadIdx.x;

    if (lane == 0) {
        mbar_init(&s_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    // Work distribution: cell cost varies by an order of magnitude (empty cells are swept twice, textured ones score hundreds of
    // survivors), so after its first cell a warp draws the next ones from a global counter.  Software pipeline over the warp's
    // cells: the draw for the cell after next flies during the sweep, the next window is prefetched into L2 at the top and
    // loaded by TMA as soon as the tile is free, and a finished cell's results are written to the corner list one cell later,
    // when the counter atomic that reserved their slots has long returned.
    const int nwork = RETRY ? *P.nitems_dev : P.nitems;
    if ((int)blockIdx.x >= nwork) return;
    int it_n = 0, nres = 0, buf = 0;
    int pn = 0, pbase = 0, pcap = 0;                                  // the previous cell's list: length, reserved base (lane 0), list capacity
    uint32_t *pdst = nullptr;
    int f, ci, nf = 0, nc = 0;                                        // (frame, cell) of the current and of the next item
    fast_split<RETRY>(P, blockIdx.x, nwork, f, ci);
    FastCellRegs C;
    C.c0 = __ldg(P.cells + 2 * ci); C.c1 = __ldg(P.cells + 2 * ci + 1);
    int nxt = 0;
    if (lane == 0) {
        mbar_expect_tx(&s_full, C.c1.w);
        tma_load_3d(s_img, &M.m[C.c0.y & 15u], (int)(C.c0.x & 0xFFFFu), (int)(C.c0.x >> 16), f, &s_full);
        nxt = atomicAdd(P.work, 1) + (int)gridDim.x;
    }
    nxt = __shfl_sync(0xffffffffu, nxt, 0);
    fast_split<RETRY>(P, nxt, nwork, nf, nc);
    for (;; it_n++) {
        const bool has_next = nxt < nwork;
        int drawn = 0;
        uint32_t nx = 0u, ny = 0u;
        if (has_next && lane == 0) {
            drawn = atomicAdd(P.work, 1) + (int)gridDim.x;                              // consumed at the bottom of the loop
            const uint4 n0 = __ldg(P.cells + 2 * nc);                                   // consumed after the tile wait below (read again later: L1-resident)
            nx = n0.x; ny = n0.y;
        }
        const int level = (int)(C.c0.y & 15u), ax = (int)((C.c0.y >> 8) & 0xFFu), dw = (int)((C.c0.y >> 16) & 0xFFu), dh = (int)(C.c0.y >> 24);
        const int R = (int)(C.c0.z & 0xFFu), nseg = (int)((C.c0.z >> 8) & 0xFFu), nG = (int)((C.c0.z >> 16) & 0xFFu), w0 = (int)(C.c0.z >> 24);
        const float inv = __uint_as_float(C.c1.x);
        uint32_t *res = s_res + buf * FS_RES;
        uint32_t *gdst = P.cand + (size_t)f * P.cand_slab + C.c1.y;
        int32_t *gcnt = &P.ncand[f * P.nlevels + level];
        // zero the score map (1-px ring included) while the tile lands
        for (int i = lane; i < ((dh + 2) * SP) / 16; i += 32) reinterpret_cast<uint4 *>(s_sc)[i] = make_uint4(0, 0, 0, 0);
        mbar_wait(&s_full, (uint32_t)(it_n & 1));
        if (has_next && lane == 0) tma_prefetch_3d(&M.m[ny & 15u], (int)(nx & 0xFFFFu), (int)(nx >> 16), nf);   // next window -> L2
        __syncwarp();

        const int units = nG * nseg;
        for (int pass = RETRY ? 1 : 0; pass < 2; pass++) {
            const int th = pass == 0 ? P.ini_th : P.min_th;
            uint32_t HM, KK;
            fast_masks(th, HM, KK);
            if (lane == 0) { s_q[0] = 0; s_q[1] = 0x7fffffff; }
            __syncwarp();
            // ---- packed sweep; survivors take queue slots with one shared-memory atomic per lane ----
            for (int u0 = 0; u0 < units; u0 += 32) {
                const int u = u0 + lane;
                if (u < units) {
                    const int seg = __float2int_rd(((float)u + 0.5f) * inv), wcol = w0 + u - seg * nG, row0 = R * seg;
                    const uint2 raw = fast_sweep7<TP>(words + row0 * (TP / 4) + wcol, HM, KK, R);
                    const int cb = 4 * wcol - (ax + 3);                       // detection column of byte 0 (-3 .. dw-1)
                    const uint32_t cm = (0xFFFFFFFFu << (8 * max(0, -cb))) & (0xFFFFFFFFu >> (32 - 8 * min(4, dw - cb)));
                    const int nv = min(R, dh - row0);                         // valid rows of the unit: flag bit 7-k of a byte = row k (k < 8), row 8+k
                    uint32_t w0f = raw.x & cm & (((0xFF00u >> min(nv, 8)) & 0xFFu) * 0x01010101u);
                    uint32_t w1f = raw.y & cm & (((0xFF00u >> max(nv - 8, 0)) & 0xFFu) * 0x01010101u);
                    if (w0f | w1f) {
                        const int cnt = __popc(w0f) + __popc(w1f), base_off = (row0 + 3) * TP + 4 * wcol;
                        int slot = atomicAdd(&s_q[0], cnt);
                        if (slot + cnt <= FS_WQ) {
#pragma unroll
                            for (int hw = 0; hw < 2; hw++) {
                                uint32_t word = hw ? w1f : w0f;
                                const int bo = base_off + 8 * hw * TP;
                                while (word) {
                                    const int bit = __ffs((int)word) - 1;
                                    word &= word - 1;
                                    wq[slot++] = (uint16_t)(bo + (7 - (bit & 7)) * TP + (bit >> 3));
                                }
                            }
                        } else {                                              // queue full: score in place, NMS will scan the whole cell
                            atomicMin(&s_q[1], slot);
#pragma unroll
                            for (int hw = 0; hw < 2; hw++) {
                                uint32_t word = hw ? w1f : w0f;
                                const int bo = base_off + 8 * hw * TP;
                                while (word) {
                                    const int bit = __ffs((int)word) - 1;
                                    word &= word - 1;
                                    fast_score_to_map<TP>(s_img, s_sc, SP, bo + (7 - (bit & 7)) * TP + (bit >> 3), ax, th);
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            const bool ovf = s_q[1] <= FS_WQ;
            const int qn = ovf ? s_q[1] : s_q[0];
            // ---- exact score of the queued survivors ----
            for (int i = lane; i < qn; i += 32) fast_score_to_map<TP>(s_img, s_sc, SP, wq[i], ax, th);
            __syncwarp();
            // ---- strict 3x3 NMS inside the cell: queue entries, or every pixel of the cell if the queue overflowed ----
            const int nitems = ovf ? dw * dh : qn;
            const float invw = ovf ? 1.0f / (float)dw : 0.0f;
            int found = 0;
            for (int i0 = 0; i0 < nitems; i0 += 32) {
                const int i = i0 + lane;
                const bool act = i < nitems;
                int r = 0, c = 0;
                if (act) {
                    if (ovf) { r = __float2int_rd(((float)i + 0.5f) * invw); c = i - r * dw; }
                    else { const int off = wq[i]; const int tr = off / TP; r = tr - 3; c = off - tr * TP - ax - 3; }
                }
                // all nine loads in flight at once; the zero ring around the map stands for "outside the cell's detection area"
                const uint8_t *q = &s_sc[(r + 1) * SP + (c + 1)];
                const int s = q[0];
                const int n0 = q[-SP - 1], n1 = q[-SP], n2 = q[-SP + 1], n3 = q[-1], n4 = q[1], n5 = q[SP - 1], n6 = q[SP], n7 = q[SP + 1];
                const int m = max(max(max(n0, n1), max(n2, n3)), max(max(n4, n5), max(n6, n7)));
                const bool ok = act && s > m;                             // strict: s > m >= 0, so s != 0
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (bal) {
                    const int nk = __popc(bal);
                    if (nres + nk > FS_RES) { fast_publish(res, nres, gcnt, gdst, (int)C.c1.z, P.status, lane); nres = 0; }
                    // box-relative coordinates: kp.pt + (j*wCell, i*hCell) — ORBextractor.cpp:865-866
                    if (ok) res[nres + __popc(bal & ((1u << lane) - 1u))] = orbx_pack((int)(C.c0.w & 0xFFFFu) + c, (int)(C.c0.w >> 16) + r, s);
                    nres += nk;
                    found += nk;
                }
            }
            __syncwarp();
            // the reference retries a cell at minThFAST iff iniThFAST produced nothing (:843-846)
            if (found > 0 || pass == 1) break;
        }
        // every lane is done with the tile, the map and the queue
        if (has_next && lane == 0) {
            const uint4 n0 = __ldg(P.cells + 2 * nc), n1 = __ldg(P.cells + 2 * nc + 1);
            mbar_expect_tx(&s_full, n1.w);
            tma_load_3d(s_img, &M.m[n0.y & 15u], (int)(n0.x & 0xFFFFu), (int)(n0.x >> 16), nf, &s_full);
        }
        // write out the previous cell's list (its counter atomic was issued one cell ago), then open this cell's
        if (pn) fast_write_out(s_res + (buf ^ 1) * FS_RES, pn, __shfl_sync(0xffffffffu, pbase, 0), pdst, pcap, P.status, lane);
        pn = nres;
        if (nres) {
            if (lane == 0) pbase = atomicAdd(gcnt, nres);
            pdst = gdst; pcap = (int)C.c1.z; buf ^= 1; nres = 0;
        }
        if (!has_next) break;
        f = nf; ci = nc;
        C.c0 = __ldg(P.cells + 2 * ci); C.c1 = __ldg(P.cells + 2 * ci + 1);
        nxt = __shfl_sync(0xffffffffu, drawn, 0);
        fast_split<RETRY>(P, nxt, nwork, nf, nc);
    }
    if (pn) fast_write_out(s_res + (buf ^ 1) * FS_RES, pn, __shfl_sync(0xffffffffu, pbase, 0), pdst, pcap, P.status, lane);
}

// ---- host: the cell records ----
static int fast_tile_pitch(const FrameGeom &G);
void orbx_build_fast_cells(const FrameGeom &G, const std::vector<uint32_t> &ctab, std::vector<uint4> &out)
{
    const int TP = fast_tile_pitch(G);
    out.resize(2 * ctab.size());
    for (size_t i = 0; i < ctab.size(); i++) {
        const uint32_t cd = ctab[i];
        const int level = (int)(cd & 15u), ci = (int)((cd >> 4) & 0x3FFFu), cj = (int)(cd >> 18);
        const LevelGeom &g = G.lv[level];
        // cell ROI in image coordinates — ORBextractor.cpp:805-822
        const int iniX = ORBX_BORDER + cj * g.wcell, iniY = ORBX_BORDER + ci * g.hcell;
        const int maxX = std::min(iniX + g.wcell + 6, g.w - ORBX_BORDER), maxY = std::min(iniY + g.hcell + 6, g.h - ORBX_BORDER);
        const int dw = maxX - iniX - 6, dh = maxY - iniY - 6;      // detection area, ROI-relative origin (3,3)
        const int ax = iniX & 15;                                  // tile byte of ROI column 0 (TMA boxes start on 16-byte columns)
        const int w0 = (ax + 3) >> 2;                              // tile word holding detection column 0
        const int nG = ((ax + 3 + dw - 1) >> 2) - w0 + 1;          // words holding detection columns (<= 19)
        // rows per sweep unit: the R in 4..FAST_RMAX that minimises (warp iterations) x (cost of a unit = 6 window rows of 7 instructions
        // + R tested rows of 23) + the per-iteration bookkeeping (~80)
        int R = 0, nseg = 0, best = 0;
        for (int r = FAST_RMAX; r >= 4; r--) {
            const int ns = (dh + r - 1) / r, cost = ((nG * ns + 31) >> 5) * (42 + 23 * r + 80);
            if (R == 0 || cost < best) { best = cost; R = r; nseg = ns; }
        }
        const float inv = 1.0f / (float)nG;
        uint32_t invb; memcpy(&invb, &inv, 4);
        uint4 c0, c1;
        c0.x = (uint32_t)((iniX & ~15) >> 2) | ((uint32_t)iniY << 16);
        c0.y = (uint32_t)level | ((uint32_t)ax << 8) | ((uint32_t)dw << 16) | ((uint32_t)dh << 24);
        c0.z = (uint32_t)R | ((uint32_t)nseg << 8) | ((uint32_t)nG << 16) | ((uint32_t)w0 << 24);
        c0.w = (uint32_t)(cj * g.wcell + 3) | ((uint32_t)(ci * g.hcell + 3) << 16);
        c1.x = invb; c1.y = (uint32_t)g.cand_off; c1.z = (uint32_t)g.cand_cap; c1.w = (uint32_t)((g.hcell + 6) * TP);
        out[2 * i] = c0; out[2 * i + 1] = c1;
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
static bool encode_level(CUtensorMap *m, const uint8_t *base, size_t pitch, int rows, size_t fstride, int frames, int box_rows, int box_words = ORBX_TMA_BOX_WORDS);
bool orbx_encode_level(CUtensorMap *m, const uint8_t *base, size_t pitch, int rows, size_t fstride, int frames, int box_rows, int box_words)
{
    return encode_level(m, base, pitch, rows, fstride, frames, box_rows, box_words);
}
static bool encode_level(CUtensorMap *m, const uint8_t *base, size_t pitch, int rows, size_t fstride, int frames, int box_rows, int box_words)
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
                !encode_level(&h->tmap_rz2[l], h->d_pyr + G.lv[l].off, (size_t)G.lv[l].pitch, G.lv[l].h, h->pyr_slab, h->prm.max_batch, ORBX_RZ_BOX_ROWS, ORBX_RZ2_BOX_WORDS) ||
                !encode_level(&h->tmap_cell[l], h->d_pyr + G.lv[l].off, (size_t)G.lv[l].pitch, G.lv[l].h, h->pyr_slab, h->prm.max_batch, G.lv[l].hcell + 6, fast_tile_pitch(G) / 4)) return -1;
        h->tmap_valid = true; h->tmap_l0 = nullptr;
    }
    if (h->tmap_l0 != l0 || h->tmap_l0_step != l0_step || h->tmap_l0_fstride != l0_fstride || h->tmap_l0_frames < nframes) {
        if (!encode_level(&h->tmap[0], l0, l0_step, G.lv[0].h, l0_fstride, nframes, G.lv[0].hcell + 6) ||
            !encode_level(&h->tmap_rz[0], l0, l0_step, G.lv[0].h, l0_fstride, nframes, ORBX_RZ_BOX_ROWS) ||
            !encode_level(&h->tmap_rz2[0], l0, l0_step, G.lv[0].h, l0_fstride, nframes, ORBX_RZ_BOX_ROWS, ORBX_RZ2_BOX_WORDS) ||
            !encode_level(&h->tmap_cell[0], l0, l0_step, G.lv[0].h, l0_fstride, nframes, G.lv[0].hcell + 6, fast_tile_pitch(G) / 4)) return -1;
        h->tmap_l0 = l0; h->tmap_l0_step = l0_step; h->tmap_l0_fstride = l0_fstride; h->tmap_l0_frames = nframes;
    }
    return 0;
}

// the warp-per-cell kernel over every (frame, cell) item, or (retry mode, after the dense formulation of k_fast_dense.cu) over the listed
// items at minThFAST only
int launch_fast_cells(orbx_handle *h, int nframes, const int32_t *d_items, const int32_t *d_nitems)
{
    ProfScope ps(h, d_items ? ORBX_K_FAST_RETRY : ORBX_K_FAST);
    const FrameGeom &G = h->geo;
    LevelMaps M;
    memcpy(M.m, h->tmap_cell, sizeof(M.m));
    FastParams P;
    P.cand = h->d_cand; P.cand_slab = G.cand_entries;
    P.ncand = h->d_ncand;
    P.cells = reinterpret_cast<const uint4 *>(h->d_cells); P.nlevels = G.nlevels; P.ncells = G.total_cells_valid; P.nitems = G.total_cells_valid * nframes; P.inv_ncells = 1.0f / (float)G.total_cells_valid;
    P.ini_th = h->prm.ini_th_fast; P.min_th = h->prm.min_th_fast;
    P.status = h->d_status;
    P.work = h->d_ncand + (size_t)h->prm.max_batch * ORBX_MAX_LEVELS;
    P.items = d_items; P.nitems_dev = d_nitems;
    P.tile_rows = G.max_hcell + 6;
    P.map_pitch = (G.max_wcell + 2 + 15) & ~15;
    const int TP = fast_tile_pitch(G);
    const size_t map_bytes = (size_t)((((P.tile_rows - 4) * P.map_pitch) + 127) & ~127);
    const size_t smem = 128 + 128 + (size_t)(((P.tile_rows * TP) + 127) & ~127) + map_bytes + FS_WQ * 2;
    auto kern = d_items ? (TP == 80 ? k_fast_cells<80, true> : k_fast_cells<96, true>) : (TP == 80 ? k_fast_cells<80, false> : k_fast_cells<96, false>);
    if (!orbx_optin_smem(h, (const void *)kern, smem)) return -1;
    if (smem != h->fast_smem || TP != h->fast_tp) {
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, smem);
        h->fast_grid_cap = std::max(1, occ) * h->sm_count; h->fast_smem = smem; h->fast_tp = TP;
    }
    int grid = std::min(P.nitems, h->fast_grid_cap);
    if (!h->opt_serial && h->opt_fast_ctas > 0) grid = std::min(grid, h->opt_fast_ctas * h->sm_count);
    orbx_launch_pdl(h, kern, dim3(grid), dim3(32), smem, h->stream, M, P, (const FrameGeom *)h->d_geo);
    return 0;
}

int launch_fast(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride)
{
    const FrameGeom &G = h->geo;
    if (G.total_cells_valid <= 0) return 0;
    if (orbx_ensure_tmaps(h, nframes, l0, l0_step, l0_fstride) != 0) return -1;
    // dense formulation for batches (ORBX_OPT_FAST_DENSE): one pass over whole levels + per-corner NMS, then this kernel on the few
    // cells that need the minThFAST retry.  The half-batch lanes of ORBX_OPT_OVERLAP share one dense scratch: warp-per-cell there.
    const bool dense = h->opt_fast_dense == 2 || (h->opt_fast_dense == 1 && nframes >= ORBX_FAST_DENSE_MIN_FRAMES);
    if (dense && !h->in_overlap && h->dense_ok) return launch_fast_dense(h, nframes);
    return launch_fast_cells(h, nframes, nullptr, nullptr);
}
