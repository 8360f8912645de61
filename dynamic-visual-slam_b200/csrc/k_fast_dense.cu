// k_fast_dense.cu — the batch formulation of the cell loop of ORBextractor::ComputeKeyPointsOctTree (reference ORBextractor.cpp:785-872).
// Same results as k_fast.cu (warp per cell), organised around two facts about the reference's cells:
//   * the DETECTION areas of the cells (ROI minus cv::FAST's 3-px ring margin) tile the level without overlap — the +6 px of a cell ROI
//     (:818-822) is only the ring margin; the FAST score S of a pixel does not depend on the cell it is evaluated in, only the 3x3 NMS
//     does (scores outside the cell's detection area count as 0) and the minThFAST retry (:843-846) is a per-cell decision;
//   * so the expensive part — pre-test and exact score at iniThFAST — can run over whole levels in regular tiles, without the per-cell
//     set-up, the partially filled sweep iterations and the queue imbalance of a 35-px cell (k_fast.cu spends ~40 % of its instructions there).
// Three steps, all on the stream of the step:
//   1. k_fast_dense: persistent one-warp CTAs draw (frame, tile) items; a tile is 128 x 16 detection pixels = ONE full sweep iteration
//      (lane = one aligned word column, 16 rows), its 144 x 22 window arrives by TMA.  Survivors of the pre-test are queued in lane
//      order (warp prefix sum, deterministic), scored exactly, and written (a) as a dense u8 score map (S-1 if S > iniThFAST else 0 —
//      cv::FAST's score buffer) with coalesced 16-byte stores and (b) as a corner list per (frame, level).
//   2. k_fast_nms: one thread per corner: its cell from the coordinates, strict 3x3 maximum against the score map restricted to the
//      cell's detection area, keypoint appended to the (frame, level) candidate list, the cell marked as served.
//   3. k_fast_retry_list + k_fast_cells in retry mode: the cells that produced nothing at iniThFAST (7 % on the bench frames) run the
//      warp-per-cell kernel at minThFAST.
// List order differs from k_fast.cu (both are arbitrary; the quadtree kernel is order-independent); the SETS are identical (tests).
#include "orbx_internal.h"
#include "orbx_tma.h"
#include "orbx_fast_dev.h"
#include <algorithm>
#include <cstring>
#include <vector>

#define FD_TP 144                 // tile pitch = TMA box width: 4 (left word) + 128 + 4 (right word) -> multiple of 16
#define FD_ROWS 16                // detection rows per tile
#define FD_BOX_ROWS (FD_ROWS + 6)
#define FD_TILE_BYTES 3200        // 22 x 144 = 3168, rounded to 128
#ifndef FD_WQ
#define FD_WQ 384                 // survivor queue entries (u16 tile offsets); a tile with more survivors takes extra rounds
#endif
#ifndef FD_RES
#define FD_RES 160                // corners of a tile (two lists: deferred publish); more: flushed in between
#endif
#ifndef FD_MINB
#define FD_MINB 24
#endif
#define FN_THREADS 256
#define FN_BLOCKS_X 12

#ifndef FD_NMS_ITEMS
#define FD_NMS_ITEMS 128          // NMS work items per frame (each takes every 128th group of 128 corners of every level)
#endif
#ifndef FD_NMS_LAG
#define FD_NMS_LAG 8              // the NMS items of frame f follow the tile items of frame f + 8 (the resident warps span ~2.3 frames of items and publish a tile one
                                  // tile late): its tiles are done when they are drawn, its maps (10 frames x 3.3 MB in flight) still in L2
#endif

struct DenseLevelDev { int map_off, map_pitch, cl_off, cl_cap; };
struct DenseParams {
    uint32_t *clist; size_t clist_slab;          // corner lists, entries per frame
    int32_t *ncorner;                            // [frame][level]
    uint8_t *smap; size_t smap_slab;             // score maps, bytes per frame
    const uint4 *tiles; int ntiles, nframes, nitems;   // nitems: tile items and NMS items (dense_item)
    float inv_per;                               // 1 / (ntiles + FD_NMS_ITEMS)
    int th, nlevels;
    int32_t *status, *work;
    int32_t *done;                               // [frame]: tiles whose score map and corners are in global memory
    DenseLevelDev lv[ORBX_MAX_LEVELS];
};

// ---- strict 3x3 NMS inside the corner's cell (cv::FAST's NMS, which the reference runs per cell ROI) ----
struct NmsLevel {
    int map_off, map_pitch, cl_off, cl_cap;
    int x1, y1, wcell, hcell, ncv, nrv, cellv_first, cand_cap;      // x1, y1: first column / row past the last cell's detection area
    unsigned mw, mh;                                                  // ceil(2^32 / wcell), ceil(2^32 / hcell): exact quotients by __umulhi for coordinates < 2^16
    unsigned cand_off;
};
struct NmsParams {
    const uint32_t *clist; size_t clist_slab; const int32_t *ncorner;
    const uint8_t *smap; size_t smap_slab;
    uint32_t *cand; size_t cand_slab; int32_t *ncand;
    uint8_t *found; int ncells, nlevels;
    int32_t *status;
    NmsLevel lv[ORBX_MAX_LEVELS];
};

// The calling warp takes the groups base0, base0 + stride, ... of the (frame, level) corner list; a lane takes FOUR consecutive entries
// (one 16-byte load) and has their 32 neighbour loads in flight together: the work is a chain of dependent global round trips
// (entry -> neighbours -> slot atomic -> store), four corners per trip instead of one.  n = entries in the list.
__device__ __forceinline__ void nms_groups(const NmsParams &P, int f, int level, int n, int base0, int stride, int lane)
{
    const NmsLevel &L = P.lv[level];
    const uint4 *list4 = reinterpret_cast<const uint4 *>(P.clist + (size_t)f * P.clist_slab + L.cl_off);   // list starts are multiples of 64 entries
    const uint8_t *map = P.smap + (size_t)f * P.smap_slab + L.map_off;
    uint32_t *cdst = P.cand + (size_t)f * P.cand_slab + L.cand_off;
    int32_t *ccnt = &P.ncand[f * P.nlevels + level];
    uint8_t *found = P.found + (size_t)f * P.ncells + L.cellv_first;
    const int pitch = L.map_pitch, wcell = L.wcell, hcell = L.hcell;
    for (int base = base0; 4 * base < n; base += stride) {
        const int g = base + lane;                                                 // group of four entries
        uint4 c4 = make_uint4(0, 0, 0, 0);
        if (4 * g < n) c4 = list4[g];
        const uint32_t cs[4] = { c4.x, c4.y, c4.z, c4.w };
        int nbr[4][8], ctr[4], cell[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t c = cs[k];
            const bool act = 4 * g + k < n;
            const int xr = act ? orbx_px(c) - 3 : 0, yr = act ? orbx_py(c) - 3 : 0;                     // relative to the first detection column / row (level 19)
            const int cj = (int)__umulhi((unsigned)xr, L.mw), ci = (int)__umulhi((unsigned)yr, L.mh);
            // the cell's detection area — ORBextractor.cpp:805-822 minus cv::FAST's 3-px margin; the corner is inside it by construction
            const int cx0 = cj * wcell, cx1 = min(cx0 + wcell, L.x1), cy0 = ci * hcell, cy1 = min(cy0 + hcell, L.y1);
            const uint8_t *p = map + yr * pitch + (xr + (ORBX_BORDER + 3 - 4));
            const bool xl = act && xr > cx0, xh = act && xr + 1 < cx1, yl = act && yr > cy0, yh = act && yr + 1 < cy1;
            nbr[k][0] = yl ? p[-pitch] : 0;        nbr[k][1] = (yl && xl) ? p[-pitch - 1] : 0;  nbr[k][2] = (yl && xh) ? p[-pitch + 1] : 0;
            nbr[k][3] = xl ? p[-1] : 0;            nbr[k][4] = xh ? p[1] : 0;
            nbr[k][5] = yh ? p[pitch] : 0;         nbr[k][6] = (yh && xl) ? p[pitch - 1] : 0;   nbr[k][7] = (yh && xh) ? p[pitch + 1] : 0;
            ctr[k] = act ? orbx_ps(c) : 0;                                       // S - 1 >= 1 for a listed corner
            cell[k] = ci * L.ncv + cj;
        }
        unsigned keep = 0u;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int m = max(max(max(nbr[k][0], nbr[k][1]), max(nbr[k][2], nbr[k][3])), max(max(nbr[k][4], nbr[k][5]), max(nbr[k][6], nbr[k][7])));
            if (ctr[k] > m) { keep |= 1u << k; found[cell[k]] = 1; }             // strict maximum; the cell is served: no minThFAST retry — ORBextractor.cpp:843-846
        }
        const int cnt = __popc(keep);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total) {
            int pos = 0;
            if (lane == 31) pos = atomicAdd(ccnt, total);
            pos = __shfl_sync(0xffffffffu, pos, 31) + incl - cnt;
#pragma unroll
            for (int k = 0; k < 4; k++) if (keep & (1u << k)) {
                if (pos < L.cand_cap) cdst[pos] = cs[k];
                else atomicOr(P.status, ORBX_DS_CAND_OVERFLOW);
                pos++;
            }
        }
    }
}

// stand-alone NMS pass (FUSE = false): grid (blocks, frames, levels)
__global__ void __launch_bounds__(FN_THREADS) k_fast_nms(NmsParams P)
{
    ORBX_PDL_ENTRY();
    const int f = blockIdx.y, level = blockIdx.z;
    const int n = min(P.ncorner[f * P.nlevels + level], P.lv[level].cl_cap);     // plain loads: written by the previous kernel of the stream
    if ((int)(blockIdx.x * FN_THREADS * 4) >= n) return;
    nms_groups(P, f, level, n, blockIdx.x * FN_THREADS + (threadIdx.x & ~31), gridDim.x * FN_THREADS, threadIdx.x & 31);
}

// tile record (host-built, frame-independent):
//   x = box word column | box row << 16          y = level | vx0 << 4 | vx1 << 12 | vh << 20   (valid detection bytes [vx0, vx1) of the 128, valid rows)
//   z = byte offset of the tile in the frame's score-map slab          w = (x - 16 of detection byte 0) & 0xFFFF | (y - 16 of detection row 0) << 16
//
// Work items of k_fast_dense.  FUSE = false (default): item = frame * ntiles + tile, the NMS is the next kernel.  FUSE = true (ORBX_OPT_FAST_DENSE = 3,
// an experiment kept with its test): as a second kernel the NMS re-reads every score map of the batch from HBM (420 MB per 128 frames of
// 1280 x 720, 0.13 ms), so here it runs inside the tile kernel while a frame's map is still in L2: the item sequence is F blocks of [ntiles tile
// items of frame q | FD_NMS_ITEMS NMS items of frame q - FD_NMS_LAG], then the NMS items of the last FD_NMS_LAG frames.  A warp that draws an NMS
// item waits until every tile of that frame has published its map and corners (done[f], a release / acquire counter); tile items never wait, a
// warp publishes its own deferred corner list before it waits and holds no drawn tile item while it waits, so the waits neither deadlock nor
// chain.  Measured: same results, but 0.565 ms against 0.32 + 0.13 ms for the two kernels — the NMS is a chain of dependent global round
// trips (entry -> 32 neighbour bytes -> slot atomic -> store, ~8 us per 128 corners on a loaded SM) whatever memory answers them, and inside the
// tile kernel each such chain occupies one of the SM's 24 one-warp CTAs instead of one of 40 cheap warps of k_fast_nms.
enum { DI_TILE = 0, DI_NMS = 1, DI_SKIP = 2, DI_END = 3 };
template <bool FUSE> __device__ __forceinline__ int dense_item(const DenseParams &P, int idx, int &f, int &t)
{
    f = 0; t = 0;
    if (idx >= P.nitems) return DI_END;
    if (!FUSE) {
        f = __float2int_rz(((float)idx + 0.5f) * P.inv_per);                   // inv_per = 1 / ntiles here
        t = idx - f * P.ntiles;
        if (t < 0) { f--; t += P.ntiles; }
        else if (t >= P.ntiles) { f++; t -= P.ntiles; }
        return DI_TILE;
    }
    const int per = P.ntiles + FD_NMS_ITEMS, body = P.nframes * per;
    if (idx >= body) {                                                          // tail: NMS of the last frames
        const int rem = idx - body, f0 = max(P.nframes - FD_NMS_LAG, 0);
        f = f0 + rem / FD_NMS_ITEMS; t = rem % FD_NMS_ITEMS;
        return DI_NMS;
    }
    int q = __float2int_rz(((float)idx + 0.5f) * P.inv_per), r = idx - q * per;
    if (r < 0) { q--; r += per; }
    else if (r >= per) { q++; r -= per; }
    if (r < P.ntiles) { f = q; t = r; return DI_TILE; }
    f = q - FD_NMS_LAG; t = r - P.ntiles;
    return f >= 0 ? DI_NMS : DI_SKIP;
}

template <bool FUSE> __global__ void __launch_bounds__(32, FD_MINB) k_fast_dense(const __grid_constant__ LevelMaps M, const __grid_constant__ DenseParams P, const __grid_constant__ NmsParams Q)
{
    extern __shared__ __align__(128) uint8_t s_raw[];
    __shared__ __align__(8) uint64_t s_full;
    __shared__ uint32_t s_fl[64];                                   // the lanes' flag words, for the redistribution before the queue walk
    ORBX_PDL_ENTRY();
    uint8_t *s_dyn = s_raw + ((128u - (smem_u32(s_raw) & 127u)) & 127u);
    uint8_t *s_img = s_dyn + 128;                                   // the word left of tile column 0 is read (never used): pad
    uint8_t *s_sc = s_img + FD_TILE_BYTES;                          // 16 x 128 score tile
    uint16_t *wq = reinterpret_cast<uint16_t *>(s_sc + FD_ROWS * 128);
    uint32_t *s_res = reinterpret_cast<uint32_t *>(wq + FD_WQ);     // 2 x FD_RES
    const uint32_t *words = reinterpret_cast<const uint32_t *>(s_img);
    const uint8_t *flb = reinterpret_cast<const uint8_t *>(s_fl);
    uint4 *sc4 = reinterpret_cast<uint4 *>(s_sc);
    const int lane = threadIdx.x;

    if (lane == 0) {
        mbar_init(&s_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
#pragma unroll
    for (int r = 0; r < 4; r++) sc4[r * 32 + lane] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    if ((int)blockIdx.x >= P.nitems) return;

    uint32_t HM, KK;
    fast_masks(P.th, HM, KK);
    const int th = P.th;
    // the lane's share of the redistributed flags: bytes lane + 32 m of the 256 flag bytes = column (lane & 3) of the word columns
    // (lane >> 3) + 4 m, rows 8 * ((lane >> 2) & 1) ...: pre-test survivors come in blobs, a blob's bytes go to different lanes
    const int walk_base = (3 + 8 * ((lane >> 2) & 1)) * FD_TP + 4 + 4 * (lane >> 3) + (lane & 3);
    int ntma = 0, nres = 0, buf = 0;
    // the previous tile: its corner list waits in shared memory for the counter atomic that reserved its slots
    bool have_prev = false;
    int pn = 0, pbase = 0, pcap = 0, pf = 0;
    uint32_t *pdst = nullptr;
    auto flush_prev = [&]() {                                       // publish the previous tile's corners, then count the tile as done
        if (!have_prev) return;
        if (pn) fast_write_out(s_res + (buf ^ 1) * FD_RES, pn, __shfl_sync(0xffffffffu, pbase, 0), pdst, pcap, P.status, lane);
        if (FUSE) {
            __syncwarp();                                           // every lane's map and list stores are ordered before lane 0's fence,
            if (lane == 0) { __threadfence(); atomicAdd(&P.done[pf], 1); }      // which is cumulative: one release per tile (32 fences cost 4x the tile)
        }
        have_prev = false; pn = 0;
    };
    auto draw = [&]() {                                             // synchronous draw of the next item index
        int v = 0;
        if (lane == 0) v = atomicAdd(P.work, 1) + (int)gridDim.x;
        return __shfl_sync(0xffffffffu, v, 0);
    };
    auto do_nms = [&](int f, int j) {                               // NMS item j of frame f: every FD_NMS_ITEMS-th group of 128 corners of every level
        if (lane == 0) {
            int spins = 0;
            while (*reinterpret_cast<volatile int32_t *>(&P.done[f]) < P.ntiles) {
                __nanosleep(256);
                if (++spins > (1 << 22)) { atomicOr(P.status, ORBX_DS_INTERNAL); break; }       // ~1 s; never seen — a wait without a bound would take the GPU with it
            }
        }
        if (lane == 0) __threadfence();                             // acquire: the frame's maps and lists are visible to the warp after the barrier
        __syncwarp();
        int nl = 0;                                                 // lane l holds level l's corner count: one round trip for all levels
        if (lane < P.nlevels) nl = min(*reinterpret_cast<volatile const int32_t *>(&Q.ncorner[f * Q.nlevels + lane]), Q.lv[lane].cl_cap);
        for (int level = 0; level < P.nlevels; level++) {
            const int n = __shfl_sync(0xffffffffu, nl, level);
            nms_groups(Q, f, level, n, j * 32, FD_NMS_ITEMS * 32, lane);
        }
    };
    // acquire(): handle the items that are not tiles until `cur` is a tile item (true) or the items are exhausted (false).  A warp that waits
    // in an NMS item must not sit on a drawn tile item (frames would complete late and the waits would chain): `nxt` is drawn only when needed
    // (-1 = not drawn), and never ahead of an NMS item.
    int cur = blockIdx.x, nxt = -1, f = 0, ti = 0;
    auto acquire = [&]() {
        for (;;) {
            const int type = dense_item<FUSE>(P, cur, f, ti);
            if (type == DI_TILE) return true;
            if (type == DI_END) return false;
            if (type == DI_NMS) { flush_prev(); do_nms(f, ti); }
            if (nxt < 0) nxt = draw();
            cur = nxt; nxt = -1;
        }
    };
    if (!acquire()) { flush_prev(); return; }
    uint4 T = __ldg(P.tiles + ti);
    if (lane == 0) {
        mbar_expect_tx(&s_full, FD_BOX_ROWS * FD_TP);
        tma_load_3d(s_img, &M.m[T.y & 15u], (int)(T.x & 0xFFFFu), (int)(T.x >> 16), f, &s_full);
    }
    for (;;) {
        // the draw for the item after next flies during the sweep (unless the next item is an NMS item: see acquire); the next item, if it is a
        // tile, is prefetched into L2
        if (nxt < 0) nxt = draw();
        int drawn = -1, nf = 0, nt = 0;
        const int next_type = dense_item<FUSE>(P, nxt, nf, nt);
        const bool next_tile = next_type == DI_TILE;
        if (next_type != DI_NMS && lane == 0) drawn = atomicAdd(P.work, 1) + (int)gridDim.x;     // consumed at the bottom of the loop
        uint4 NT = make_uint4(0, 0, 0, 0);
        if (next_tile) NT = __ldg(P.tiles + nt);
        const int level = (int)(T.y & 15u);
        const int xb = (int)(int16_t)(T.w & 0xFFFFu), yb = (int)(T.w >> 16);
        const DenseLevelDev LV = P.lv[level];
        uint32_t *res = s_res + buf * FD_RES;
        uint32_t *gdst = P.clist + (size_t)f * P.clist_slab + LV.cl_off;
        int32_t *gcnt = &P.ncorner[f * P.nlevels + level];
        mbar_wait(&s_full, (uint32_t)(ntma & 1));
        ntma++;
        if (next_tile && lane == 0) tma_prefetch_3d(&M.m[NT.y & 15u], (int)(NT.x & 0xFFFFu), (int)(NT.x >> 16), nf);   // next window -> L2
        __syncwarp();

        // ---- packed sweep: lane = word column (4 pixels) x 16 rows ----
        const uint2 raw = fast_sweep7<FD_TP>(words + 1 + lane, HM, KK, FD_ROWS);
        uint32_t p0 = raw.x, p1 = raw.y;                               // flag bit 7-k of byte j = row k (k + 8 in p1), column j
        if ((T.y >> 4) != ((128u << 8) | (16u << 16))) {              // partial tile: mask the columns and rows outside the cells
            const int vx0 = (int)((T.y >> 4) & 0xFFu), vx1 = (int)((T.y >> 12) & 0xFFu), vh = (int)((T.y >> 20) & 0x1Fu);
            const int cb = 4 * lane;
            const int lo = min(4, max(0, vx0 - cb)), hi = min(4, max(0, vx1 - cb));
            const uint32_t cm = hi > lo ? ((0xFFFFFFFFu << (8 * lo)) & (0xFFFFFFFFu >> (32 - 8 * hi))) : 0u;
            p0 &= cm & (((0xFF00u >> min(vh, 8)) & 0xFFu) * 0x01010101u);
            p1 &= cm & (((0xFF00u >> max(vh - 8, 0)) & 0xFFu) * 0x01010101u);
        }
        // ---- redistribute the flag bytes over the lanes ----
        s_fl[2 * lane] = p0; s_fl[2 * lane + 1] = p1;
        __syncwarp();
        uint32_t q0 = (uint32_t)flb[lane] | ((uint32_t)flb[lane + 32] << 8) | ((uint32_t)flb[lane + 64] << 16) | ((uint32_t)flb[lane + 96] << 24);
        uint32_t q1 = (uint32_t)flb[lane + 128] | ((uint32_t)flb[lane + 160] << 8) | ((uint32_t)flb[lane + 192] << 16) | ((uint32_t)flb[lane + 224] << 24);
        // ---- queue in lane order (prefix sum), exact scores dense over the lanes; more than FD_WQ survivors: further rounds ----
        for (;;) {
            const int cnt = __popc(q0) + __popc(q1);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) break;
            const bool fit = incl <= FD_WQ;
            int qn = total;
            if (total > FD_WQ) {
                const unsigned nofit = __ballot_sync(0xffffffffu, !fit);
                qn = __shfl_sync(0xffffffffu, incl - cnt, __ffs((int)nofit) - 1);
            }
            if (fit && cnt) {
                uint16_t *dst = wq + (incl - cnt);
#pragma unroll
                for (int hw = 0; hw < 2; hw++) {
                    uint32_t word = hw ? q1 : q0;
                    const int wb = walk_base + 64 * hw;
                    while (word) {                                             // two flag bits per trip
                        const int b0 = __ffs((int)word) - 1;
                        word &= word - 1;
                        *dst++ = (uint16_t)(wb + (7 - (b0 & 7)) * FD_TP + 2 * (b0 & 24));           // row 7 - (bit & 7), word column + 4 (bit >> 3)
                        if (word) {
                            const int b1 = __ffs((int)word) - 1;
                            word &= word - 1;
                            *dst++ = (uint16_t)(wb + (7 - (b1 & 7)) * FD_TP + 2 * (b1 & 24));
                        }
                    }
                }
                q0 = 0u; q1 = 0u;
            }
            __syncwarp();
            for (int i0 = 0; i0 < qn; i0 += 32) {
                const int i = i0 + lane;
                int val = 0, off = 0;
                if (i < qn) {
                    off = wq[i];
                    const int s = fast_score_packed<FD_TP>(s_img + off);
                    val = s > th ? s - 1 : 0;
                }
                const bool ok = val > 0;                                       // a corner whose score S-1 is 0 (th = 0) never survives the strict NMS
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (bal) {
                    const int tr = off / FD_TP, tc = off - tr * FD_TP - 4;
                    const int n = __popc(bal);
                    if (nres + n > FD_RES) { fast_publish(res, nres, gcnt, gdst, LV.cl_cap, P.status, lane); nres = 0; }
                    if (ok) {
                        s_sc[(tr - 3) * 128 + tc] = (uint8_t)val;
                        res[nres + __popc(bal & ((1u << lane) - 1u))] = orbx_pack(xb + tc, yb + tr - 3, val);
                    }
                    nres += n;
                }
            }
            __syncwarp();
            if (total <= FD_WQ) break;
        }
        // every lane is done with the tile: load the next one, then move the score tile out (and zero it for the next tile)
        if (next_tile && lane == 0) {
            mbar_expect_tx(&s_full, FD_BOX_ROWS * FD_TP);
            tma_load_3d(s_img, &M.m[NT.y & 15u], (int)(NT.x & 0xFFFFu), (int)(NT.x >> 16), nf, &s_full);
        }
        {
            uint8_t *mp = P.smap + (size_t)f * P.smap_slab + T.z;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int idx = r * 32 + lane;                                   // row = idx / 8, 16-byte segment = idx % 8
                const uint4 v = sc4[idx];
                sc4[idx] = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(mp + (idx >> 3) * LV.map_pitch + (idx & 7) * 16) = v;
            }
        }
        // write out the previous tile's corners (their counter atomic was issued one tile ago), then open this tile's
        flush_prev();
        have_prev = true; pf = f; pn = nres;
        if (nres) { if (lane == 0) pbase = atomicAdd(gcnt, nres); pdst = gdst; pcap = LV.cl_cap; }
        buf ^= 1; nres = 0;
        cur = nxt; nxt = __shfl_sync(0xffffffffu, drawn, 0);
        if (next_tile) { f = nf; ti = nt; T = NT; continue; }
        if (!acquire()) break;                                                   // NMS items (and the end) are handled there
        T = __ldg(P.tiles + ti);
        if (lane == 0) {
            mbar_expect_tx(&s_full, FD_BOX_ROWS * FD_TP);
            tma_load_3d(s_img, &M.m[T.y & 15u], (int)(T.x & 0xFFFFu), (int)(T.x >> 16), f, &s_full);
        }
    }
    flush_prev();
}

// ---- step 3: the cells iniThFAST left empty -> item list of the retry launch ----
__global__ void __launch_bounds__(256) k_fast_retry_list(const uint8_t *found, int n, int32_t *list, int32_t *count)
{
    ORBX_PDL_ENTRY();
    const int i = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31;
    const bool need = i < n && found[i] == 0;
    const unsigned bal = __ballot_sync(0xffffffffu, need);
    if (!bal) return;
    int pos = 0;
    if (lane == __ffs((int)bal) - 1) pos = atomicAdd(count, __popc(bal));
    pos = __shfl_sync(0xffffffffu, pos, __ffs((int)bal) - 1) + __popc(bal & ((1u << lane) - 1u));
    if (need) list[pos] = i;
}

// ---- host ----
static size_t up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// dense layout of a frame geometry: tiles, score maps, corner lists.  `tiles` may be null (sizing at create).
void orbx_build_fast_dense(const FrameGeom &G, int cand_divisor, DenseGeom &D, std::vector<uint4> *tiles)
{
    const int cdiv = cand_divisor > 0 ? cand_divisor : 16;
    size_t moff = 0, coff = 0;
    int ntiles = 0, cellv = 0;
    memset(&D, 0, sizeof(D));
    for (int l = 0; l < G.nlevels; l++) {
        const LevelGeom &g = G.lv[l];
        DenseLevel &d = D.lv[l];
        int nrv = 0, ncv = 0;                                          // cells the reference runs (ORBextractor.cpp:811-816), as in build_geometry
        while (nrv < g.nrows && ORBX_BORDER + nrv * g.hcell < g.h - ORBX_BORDER - 3) nrv++;
        while (ncv < g.ncols && ORBX_BORDER + ncv * g.wcell < g.w - ORBX_BORDER - 6) ncv++;
        d.ncv = ncv; d.nrv = nrv; d.cellv_first = cellv; cellv += ncv * nrv;
        const int x0 = ORBX_BORDER + 3, y0 = ORBX_BORDER + 3;
        const int xend = std::min(x0 + ncv * g.wcell, g.w - ORBX_BORDER - 3), yend = std::min(y0 + nrv * g.hcell, g.h - ORBX_BORDER - 3);
        if (ncv == 0 || nrv == 0 || xend <= x0 || yend <= y0) continue;
        d.ntx = (xend - 4 + 127) / 128; d.nty = (yend - y0 + FD_ROWS - 1) / FD_ROWS;
        d.map_pitch = d.ntx * 128; d.map_off = (int)moff; moff += (size_t)d.map_pitch * d.nty * FD_ROWS;
        d.cl_cap = (int)up((size_t)std::max(4096, (int)std::min<long long>((long long)(xend - x0) * (yend - y0), 4ll * g.w * g.h / cdiv)), 64);   // every iniThFAST corner before NMS (noise: a third of the pixels)
        d.cl_off = (int)coff; coff += (size_t)d.cl_cap;
        for (int ty = 0; ty < d.nty; ty++) for (int tx = 0; tx < d.ntx; tx++) {
            const int bx = 128 * tx + 4, by = y0 + FD_ROWS * ty;       // level column / row of detection byte 0 / row 0 of the tile
            const int vx0 = std::min(128, std::max(0, x0 - bx)), vx1 = std::min(128, std::max(0, xend - bx));
            const int vh = std::min(FD_ROWS, yend - by);
            if (vx1 <= vx0 || vh <= 0) continue;
            ntiles++;
            if (!tiles) continue;
            uint4 t;
            t.x = (uint32_t)(32 * tx) | ((uint32_t)(ORBX_BORDER + FD_ROWS * ty) << 16);
            t.y = (uint32_t)l | ((uint32_t)vx0 << 4) | ((uint32_t)vx1 << 12) | ((uint32_t)vh << 20);
            t.z = (uint32_t)(d.map_off + FD_ROWS * ty * d.map_pitch + 128 * tx);
            t.w = ((uint32_t)(bx - ORBX_BORDER) & 0xFFFFu) | ((uint32_t)(3 + FD_ROWS * ty) << 16);
            tiles->push_back(t);
        }
    }
    D.ntiles = ntiles; D.map_bytes = up(std::max<size_t>(moff, 256), 256); D.cl_entries = std::max<size_t>(coff, 64);
}

static int ensure_dense_tmaps(orbx_handle *h, int nframes)
{
    const FrameGeom &G = h->geo;
    const uint8_t *l0 = h->tmap_l0; const size_t l0_step = h->tmap_l0_step, l0_fstride = h->tmap_l0_fstride; const int l0_frames = h->tmap_l0_frames;   // as orbx_ensure_tmaps left them
    if (h->tmap_dense_pyr != h->d_pyr || h->tmap_dense_serial != h->geo_serial) {
        for (int l = 1; l < G.nlevels; l++)
            if (!orbx_encode_level(&h->tmap_dense[l], h->d_pyr + G.lv[l].off, (size_t)G.lv[l].pitch, G.lv[l].h, h->pyr_slab, h->prm.max_batch, FD_BOX_ROWS, FD_TP / 4)) return -1;
        h->tmap_dense_pyr = h->d_pyr; h->tmap_dense_serial = h->geo_serial; h->tmap_dense_l0 = nullptr;
    }
    if (h->tmap_dense_l0 != l0 || h->tmap_dense_l0_step != l0_step || h->tmap_dense_l0_fstride != l0_fstride || h->tmap_dense_l0_frames != l0_frames) {
        if (!orbx_encode_level(&h->tmap_dense[0], l0, l0_step, G.lv[0].h, l0_fstride, l0_frames, FD_BOX_ROWS, FD_TP / 4)) return -1;
        h->tmap_dense_l0 = l0; h->tmap_dense_l0_step = l0_step; h->tmap_dense_l0_fstride = l0_fstride; h->tmap_dense_l0_frames = l0_frames;
    }
    (void)nframes;
    return 0;
}

int launch_fast_dense(orbx_handle *h, int nframes)
{
    const FrameGeom &G = h->geo;
    const DenseGeom &D = h->dgeo;
    if (D.ntiles <= 0) return launch_fast_cells(h, nframes, nullptr, nullptr);
    if (ensure_dense_tmaps(h, nframes) != 0) return -1;
    const int ncells = G.total_cells_valid;
    // one memset: [work counter, retry count, 2 spare][corner counts][tiles done per frame][served flags]
    const size_t B = (size_t)h->prm.max_batch;
    int32_t *zero = h->d_dense_zero;
    int32_t *ncorner = zero + 4;
    int32_t *done = ncorner + B * ORBX_MAX_LEVELS;
    uint8_t *found = reinterpret_cast<uint8_t *>(done + B);
    cudaMemsetAsync(zero, 0, (4 + B * ORBX_MAX_LEVELS + B) * sizeof(int32_t) + (size_t)nframes * ncells, h->stream);
    const bool fuse = h->opt_dense_fuse != 0;

    LevelMaps M;
    memcpy(M.m, h->tmap_dense, sizeof(M.m));
    DenseParams P;
    P.clist = h->d_clist; P.clist_slab = D.cl_entries; P.ncorner = ncorner;
    P.smap = h->d_smap; P.smap_slab = D.map_bytes;
    P.tiles = reinterpret_cast<const uint4 *>(h->d_dtiles); P.ntiles = D.ntiles; P.nframes = nframes;
    if (fuse) {                                                       // dense_item: F blocks of (tiles, NMS items of the frame FD_NMS_LAG back), then the last frames' NMS items
        P.nitems = nframes * (D.ntiles + FD_NMS_ITEMS) + std::min(nframes, FD_NMS_LAG) * FD_NMS_ITEMS;
        P.inv_per = 1.0f / (float)(D.ntiles + FD_NMS_ITEMS);
    } else { P.nitems = D.ntiles * nframes; P.inv_per = 1.0f / (float)D.ntiles; }
    P.th = h->prm.ini_th_fast; P.nlevels = G.nlevels;
    P.status = h->d_status; P.work = zero; P.done = done;
    NmsParams Q;
    Q.clist = h->d_clist; Q.clist_slab = D.cl_entries; Q.ncorner = ncorner;
    Q.smap = h->d_smap; Q.smap_slab = D.map_bytes;
    Q.cand = h->d_cand; Q.cand_slab = G.cand_entries; Q.ncand = h->d_ncand;
    Q.found = found; Q.ncells = ncells; Q.nlevels = G.nlevels; Q.status = h->d_status;
    for (int l = 0; l < ORBX_MAX_LEVELS; l++) {
        const DenseLevel &d = D.lv[l];
        DenseLevelDev &p = P.lv[l];
        NmsLevel &q = Q.lv[l];
        memset(&p, 0, sizeof(p)); memset(&q, 0, sizeof(q));
        q.wcell = q.hcell = 1;
        if (l >= G.nlevels) continue;
        const LevelGeom &g = G.lv[l];
        p.map_off = d.map_off; p.map_pitch = d.map_pitch; p.cl_off = d.cl_off; p.cl_cap = d.cl_cap;
        q.map_off = d.map_off; q.map_pitch = d.map_pitch; q.cl_off = d.cl_off; q.cl_cap = d.cl_cap;
        q.wcell = g.wcell; q.hcell = g.hcell; q.ncv = d.ncv; q.nrv = d.nrv; q.cellv_first = d.cellv_first; q.cand_cap = g.cand_cap; q.cand_off = (unsigned)g.cand_off;
        q.x1 = std::min(d.ncv * g.wcell, g.w - 2 * (ORBX_BORDER + 3)); q.y1 = std::min(d.nrv * g.hcell, g.h - 2 * (ORBX_BORDER + 3));
        q.mw = (unsigned)((0x100000000ull + (unsigned)g.wcell - 1) / (unsigned)g.wcell); q.mh = (unsigned)((0x100000000ull + (unsigned)g.hcell - 1) / (unsigned)g.hcell);
    }
    const size_t smem = 128 + 128 + FD_TILE_BYTES + FD_ROWS * 128 + FD_WQ * 2 + 2 * FD_RES * 4;
    auto kern = fuse ? k_fast_dense<true> : k_fast_dense<false>;
    if (!h->dense_grid_cap) {
        if (!orbx_optin_smem(h, (const void *)kern, smem)) return -1;
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fast_dense<true>, 32, smem);
        h->dense_grid_cap = std::max(1, occ) * h->sm_count;             // the fused kernel's waits need every CTA of the grid resident
    }
    int grid = std::min(P.nitems, h->dense_grid_cap);
    if (!h->opt_serial && h->opt_fast_ctas > 0) grid = std::min(grid, h->opt_fast_ctas * h->sm_count);
    { ProfScope ps(h, ORBX_K_FAST_DENSE); orbx_launch_pdl(h, kern, dim3(grid), dim3(32), smem, h->stream, M, P, Q); }
    const int nit = nframes * ncells;
    {
        ProfScope ps(h, ORBX_K_FAST_NMS);
        if (!fuse) { orbx_launch_pdl(h, k_fast_nms, dim3(FN_BLOCKS_X, nframes, G.nlevels), dim3(FN_THREADS), 0, h->stream, Q); h->launches++; }
        orbx_launch_pdl(h, k_fast_retry_list, dim3((nit + 255) / 256), dim3(256), 0, h->stream, (const uint8_t *)found, nit, h->d_retry, zero + 1);
    }
    return launch_fast_cells(h, nframes, h->d_retry, zero + 1);
}
