// k_quadtree.cu — spatial distribution of FAST candidates ("OctTree"), one CTA per (frame, level).
// Replaces ORBextractor::DistributeOctTree + ExtractorNode::DivideNode + compareNodes
// (reference ORBextractor.cpp:555-779, 480-536, 538-553; details SURVEY App. A.7).
//
// The reference is a std::list algorithm (push_front children, erase parent, std::sort of the
// expandable nodes in the last rounds).  It is restated here without a list:
//   * every node gets a creation sequence number; all insertions are push_front (roots push_back),
//     so the final list order is simply DESCENDING creation number of the surviving nodes;
//   * one "round" splits a set of nodes in a given visit order; children numbering, the positions of
//     expandable children in the next round's array and the early stop `lNodes.size() >= N` are
//     prefix sums over that visit order;
//   * the std::sort call (:700) is reproduced by a single-thread restatement of libstdc++'s introsort
//     so that ties of the comparator land exactly where the reference leaves them;
//   * "best response per node, first in insertion order on ties" (:757-776) is a min over the key
//     (255-score, cell row, cell col, y, x), which is the order the cell loop pushed the candidates.
// Candidate values live in global scratch (ping-pong); node bookkeeping lives in shared memory.
#include "orbx_internal.h"

// block size NT is a template parameter: 256 threads when the batch fills the machine with CTAs (one per frame x level), 1024
// for small batches, where the kernel is a chain of dependent global-memory passes and wider blocks shorten every pass
#define QT_NONE 0xFFFFu
#ifndef QT_BATCH_THREADS
#define QT_BATCH_THREADS 256         // block size when the batch fills the machine with CTAs (measured: 128 / 512 in DESIGN.md)
#endif

struct QtParams {
    uint32_t *candA, *candB;     // per-frame slabs of packed candidates
    uint16_t *ownA, *ownB;
    uint32_t *tmp;
    size_t cand_slab;
    const int32_t *ncand;
    uint32_t *sel; int sel_slab;
    int32_t *nsel;
    int32_t *status;
    int node_cap;
};

struct QNodes {                  // struct-of-arrays views into dynamic shared memory
    short *x0, *x1, *y0, *y1;
    unsigned *beg, *cnt, *seq;
};

__device__ __forceinline__ QNodes carve_nodes(unsigned char *&p, int cap)
{
    QNodes q;
    q.beg = (unsigned *)p; p += sizeof(unsigned) * cap;
    q.cnt = (unsigned *)p; p += sizeof(unsigned) * cap;
    q.seq = (unsigned *)p; p += sizeof(unsigned) * cap;
    q.x0 = (short *)p; p += sizeof(short) * cap;
    q.x1 = (short *)p; p += sizeof(short) * cap;
    q.y0 = (short *)p; p += sizeof(short) * cap;
    q.y1 = (short *)p; p += sizeof(short) * cap;
    return q;
}

// block-wide exclusive scan of packed 3x16-bit counters held in s_val[0..n); returns total.
// s_val is overwritten with the exclusive prefix.  Uses s_warp[NT/32 + 1].
template <int NT> __device__ unsigned long long block_scan_excl(unsigned long long *s_val, int n, unsigned long long *s_warp)
{
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += NT) {
        const int i = base + threadIdx.x;
        unsigned long long v = i < n ? s_val[i] : 0ull, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned long long wv = lane < NT / 32 ? s_warp[lane] : 0ull, wx = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xffffffffu, wx, o);
                if (lane >= o) wx += y;
            }
            if (lane < NT / 32) s_warp[lane] = wx - wv;
            if (lane == NT / 32 - 1) s_warp[NT / 32] = wx;
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        if (i < n) s_val[i] = carry + s_warp[wid] + x - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_warp[NT / 32];
        __syncthreads();
    }
    return s_carry;
}

#include "orbx_sort.h"

template <int NT> __global__ void __launch_bounds__(NT) k_quadtree(QtParams P, const FrameGeom *__restrict__ G)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ unsigned long long s_warp[NT / 32 + 1];
    __shared__ int s_M, s_nleaf, s_seq0, s_size, s_nsplit, s_state;
    __shared__ int s_sortcnt[2];

    // level-major launch order: every frame's level 0 (most candidates, longest) first, the short top levels last — they fill the
    // slots the first wave frees instead of trailing behind it
    ORBX_PDL_ENTRY();
    const int level = blockIdx.y;
    const int f = blockIdx.x;
    const LevelGeom &g = G->lv[level];
    const int NC = P.node_cap;
    const int slot = f * G->nlevels + level;

    unsigned char *sp = s_raw;
    QNodes cur = carve_nodes(sp, NC), nxt = carve_nodes(sp, NC), leaf = carve_nodes(sp, NC);
    unsigned long long *s_scan = (unsigned long long *)sp; sp += sizeof(unsigned long long) * NC;
    unsigned *s_c4 = (unsigned *)sp;       sp += sizeof(unsigned) * 4 * NC;     // per-quadrant counts, then child begins
    unsigned short *s_cown = (unsigned short *)sp; sp += sizeof(unsigned short) * 4 * NC;   // child owner ids
    short *s_midx = (short *)sp;           sp += sizeof(short) * NC;
    short *s_midy = (short *)sp;           sp += sizeof(short) * NC;
    unsigned short *s_vis = (unsigned short *)sp; sp += sizeof(unsigned short) * NC;   // visit rank -> node index
    unsigned short *s_rank = (unsigned short *)sp; sp += sizeof(unsigned short) * NC;  // node index -> visit rank
    SrtE *s_srt = (SrtE *)sp;              sp += sizeof(SrtE) * NC;

    uint32_t *A = P.candA + (size_t)f * P.cand_slab + g.cand_off;
    uint32_t *B = P.candB + (size_t)f * P.cand_slab + g.cand_off;
    uint16_t *oA = P.ownA + (size_t)f * P.cand_slab + g.cand_off;
    uint16_t *oB = P.ownB + (size_t)f * P.cand_slab + g.cand_off;
    uint32_t *tmp = P.tmp + (size_t)f * P.cand_slab + g.cand_off;
    uint32_t *sel = P.sel + (size_t)f * P.sel_slab + g.sel_off;

    int n = P.ncand[slot];
    if (n > g.cand_cap) n = g.cand_cap;          // overflow already flagged by the FAST kernel
    const int N = g.N;
    if (n <= 0 || g.nini < 1) { if (threadIdx.x == 0) P.nsel[slot] = 0; return; }
    const int boxH = g.h - 2 * ORBX_BORDER;
    const int nini = g.nini;
    const float hX = g.hx;

    // ---- roots (ORBextractor.cpp:567-600).  Root i has creation number nini-1-i (push_back order). ----
    for (int i = threadIdx.x; i < nini; i += NT) s_c4[i] = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += NT) {
        const uint32_t c = A[p];
        int r = (int)((float)orbx_px(c) / hX);
        if (r >= nini) r = nini - 1;
        const unsigned k = atomicAdd(&s_c4[r], 1u);
        tmp[p] = (uint32_t)r | (k << 8);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // array order = creation order = root nini-1 first ... root 0 last
        int M = 0, nl = 0; unsigned run = 0;
        for (int i = 0; i < nini; i++) { s_scan[i] = run; run += s_c4[i]; }      // begin of root i's range
        for (int k = 0; k < nini; k++) {
            const int i = nini - 1 - k;
            const unsigned c = s_c4[i];
            if (c == 0) continue;
            QNodes &dst = c == 1 ? leaf : cur;
            const int j = c == 1 ? nl++ : M++;
            dst.beg[j] = (unsigned)s_scan[i]; dst.cnt[j] = c; dst.seq[j] = (unsigned)k;
            dst.x0[j] = (short)(int)(hX * (float)i); dst.x1[j] = (short)(int)(hX * (float)(i + 1));
            dst.y0[j] = 0; dst.y1[j] = (short)boxH;
            s_c4[i] = c == 1 ? QT_NONE : (unsigned)j;                           // root -> owner id
        }
        s_M = M; s_nleaf = nl; s_seq0 = nini; s_size = M + nl; s_state = 0;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += NT) {
        const uint32_t t = tmp[p];
        const int r = t & 255; const unsigned k = t >> 8;
        const unsigned np = (unsigned)s_scan[r] + k;
        B[np] = A[p]; oB[np] = (uint16_t)s_c4[r];
    }
    __syncthreads();
    { uint32_t *t1 = A; A = B; B = t1; uint16_t *t2 = oA; oA = oB; oB = t2; }

    // ---- rounds.  s_state: 0 = normal iteration, 1 = sorted phase, 2 = finished ----
    int guard = 0;
    while (true) {
        const int M = s_M, state = s_state, prevSize = s_size;
        if (state == 2 || ++guard > 64) break;
        // visit order: normal = reverse creation order (list order); sorted = descending (count, UL.x)
        if (state == 1) {
            for (int k = threadIdx.x; k < M; k += NT) s_srt[k] = srt_make(cur.cnt[k], cur.x0[k], k);
            __syncthreads();
            block_sort_exact<NT>(s_srt, M, (SrtE *)s_scan, s_cown, s_cown + NC, s_c4, NC, s_sortcnt);
            for (int v = threadIdx.x; v < M; v += NT) { const int k = srt_pay(s_srt[M - 1 - v]); s_vis[v] = (unsigned short)k; s_rank[k] = (unsigned short)v; }
        } else {
            for (int v = threadIdx.x; v < M; v += NT) { const int k = M - 1 - v; s_vis[v] = (unsigned short)k; s_rank[k] = (unsigned short)v; }
        }
        // split geometry (DivideNode :482-483) and zeroed quadrant counters
        for (int k = threadIdx.x; k < M; k += NT) {
            s_midx[k] = (short)(cur.x0[k] + (int)ceilf((float)(cur.x1[k] - cur.x0[k]) / 2));
            s_midy[k] = (short)(cur.y0[k] + (int)ceilf((float)(cur.y1[k] - cur.y0[k]) / 2));
            s_c4[4 * k] = s_c4[4 * k + 1] = s_c4[4 * k + 2] = s_c4[4 * k + 3] = 0;
        }
        __syncthreads();
        // quadrant of every live candidate (:515-529) and its rank inside the child
        for (int p = threadIdx.x; p < n; p += NT) {
            const unsigned k = oA[p];
            if (k == QT_NONE) continue;
            const uint32_t c = A[p];
            const int q = (orbx_px(c) < s_midx[k] ? 0 : 1) + (orbx_py(c) < s_midy[k] ? 0 : 2);
            const unsigned r = atomicAdd(&s_c4[4 * k + q], 1u);
            tmp[p] = (uint32_t)q | (r << 2);
        }
        __syncthreads();
        // per visit rank: (#non-empty children, #expandable children, #singleton children)
        for (int v = threadIdx.x; v < M; v += NT) {
            const int k = s_vis[v];
            unsigned ne = 0, nm = 0, ns = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) { const unsigned c = s_c4[4 * k + q]; ne += c > 0; nm += c > 1; ns += c == 1; }
            s_scan[v] = (unsigned long long)ne | ((unsigned long long)nm << 16) | ((unsigned long long)ns << 32);
        }
        __syncthreads();
        const unsigned long long tot = block_scan_excl<NT>(s_scan, M, s_warp);
        // early stop of the sorted phase: first visit rank after which lNodes.size() >= N (:743-744)
        if (threadIdx.x == 0) s_nsplit = M;
        __syncthreads();
        if (state == 1) {
            for (int v = threadIdx.x; v < M; v += NT) {
                const int k = s_vis[v];
                unsigned ne = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) ne += s_c4[4 * k + q] > 0;
                const int before = prevSize + (int)(s_scan[v] & 0xFFFF) - v;          // size before visiting v
                const int after = before + (int)ne - 1;
                if (after >= N && before < N) s_nsplit = v + 1;                         // unique: size is monotone
            }
            __syncthreads();
        }
        const int nsplit = s_nsplit;
        const int seq0 = s_seq0, nleaf0 = s_nleaf;
        unsigned long long totS = tot;
        if (nsplit < M) totS = s_scan[nsplit];                                          // prefix up to the stop
        const int newM = (int)((totS >> 16) & 0xFFFF), newLeaf = (int)((totS >> 32) & 0xFFFF), newNe = (int)(totS & 0xFFFF);
        const int unsplit = M - nsplit;
        if (newM > NC || nleaf0 + newLeaf + unsplit > NC) {         // cannot happen within the reference's bounds
            if (threadIdx.x == 0) { atomicOr(P.status, ORBX_DS_NODE_OVERFLOW); P.nsel[slot] = 0; }
            return;
        }
        // create children (push_front n1..n4 :639-676) / keep unsplit nodes as final nodes
        for (int v = threadIdx.x; v < M; v += NT) {
            const int k = s_vis[v];
            if (v >= nsplit) {
                const int j = nleaf0 + newLeaf + (v - nsplit);
                leaf.beg[j] = cur.beg[k]; leaf.cnt[j] = cur.cnt[k]; leaf.seq[j] = cur.seq[k];
                continue;
            }
            const unsigned long long pre = s_scan[v];
            unsigned sq = seq0 + (unsigned)(pre & 0xFFFF), mi = (unsigned)((pre >> 16) & 0xFFFF), li = nleaf0 + (unsigned)((pre >> 32) & 0xFFFF);
            unsigned b = cur.beg[k];
            const short X0 = cur.x0[k], X1 = cur.x1[k], Y0 = cur.y0[k], Y1 = cur.y1[k], MX = s_midx[k], MY = s_midy[k];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned c = s_c4[4 * k + q];
                s_c4[4 * k + q] = b;                         // child begin
                unsigned short own = QT_NONE;
                if (c > 1) {
                    nxt.beg[mi] = b; nxt.cnt[mi] = c; nxt.seq[mi] = sq++;
                    nxt.x0[mi] = (q & 1) ? MX : X0; nxt.x1[mi] = (q & 1) ? X1 : MX;
                    nxt.y0[mi] = (q & 2) ? MY : Y0; nxt.y1[mi] = (q & 2) ? Y1 : MY;
                    own = (unsigned short)mi; mi++;
                } else if (c == 1) {
                    leaf.beg[li] = b; leaf.cnt[li] = 1; leaf.seq[li] = sq++; li++;
                }
                s_cown[4 * k + q] = own;
                b += c;
            }
        }
        __syncthreads();
        // move candidates into their child's range
        for (int p = threadIdx.x; p < n; p += NT) {
            const unsigned k = oA[p];
            if (k == QT_NONE || s_rank[k] >= nsplit) { B[p] = A[p]; oB[p] = QT_NONE; continue; }
            const uint32_t t = tmp[p];
            const int q = t & 3; const unsigned np = s_c4[4 * k + q] + (t >> 2);
            B[np] = A[p]; oB[np] = s_cown[4 * k + q];
        }
        __syncthreads();
        { uint32_t *t1 = A; A = B; B = t1; uint16_t *t2 = oA; oA = oB; oB = t2; }
        { QNodes t3 = cur; cur = nxt; nxt = t3; }
        if (threadIdx.x == 0) {
            const int size = prevSize + newNe - nsplit;
            s_M = newM; s_nleaf = nleaf0 + newLeaf + unsplit; s_seq0 = seq0 + newNe; s_size = size;
            // :682-686 / :747-748
            if (size >= N || size == prevSize) s_state = 2;
            else if (state == 0 && size + 3 * newM > N) s_state = 1;
        }
        __syncthreads();
    }
    // remaining expandable nodes are final nodes too
    {
        const int M = s_M, nl = s_nleaf;
        for (int k = threadIdx.x; k < M; k += NT) { leaf.beg[nl + k] = cur.beg[k]; leaf.cnt[nl + k] = cur.cnt[k]; leaf.seq[nl + k] = cur.seq[k]; }
        __syncthreads();
        if (threadIdx.x == 0) s_nleaf = nl + M;
        __syncthreads();
    }
    const int L = s_nleaf;
    if (L > g.sel_cap) { if (threadIdx.x == 0) { atomicOr(P.status, ORBX_DS_NODE_OVERFLOW); P.nsel[slot] = 0; } return; }
    // ---- best candidate per node (:757-776), output in list order = descending creation number ----
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = wid; i < L; i += NT / 32) {
        const unsigned b = leaf.beg[i], c = leaf.cnt[i];
        unsigned long long best = ~0ull;
        for (unsigned t = lane; t < c; t += 32) {
            const uint32_t cv = A[b + t];
            const int x = orbx_px(cv), y = orbx_py(cv), s = orbx_ps(cv);
            const unsigned cell = (unsigned)((y - 3) / g.hcell) * (unsigned)g.ncols + (unsigned)((x - 3) / g.wcell);
            const unsigned long long key = ((unsigned long long)(255 - s) << 40) | ((unsigned long long)cell << 24) |
                                           ((unsigned long long)y << 12) | (unsigned long long)x;
            best = key < best ? key : best;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o); best = y < best ? y : best; }
        if (lane == 0) {
            const int x = (int)(best & 0xFFF), y = (int)((best >> 12) & 0xFFF), s = 255 - (int)(best >> 40);
            s_scan[i] = ((unsigned long long)leaf.seq[i] << 32) | orbx_pack(x, y, s);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += NT) {
        const unsigned sq = (unsigned)(s_scan[i] >> 32);
        int rank = 0;
        for (int j = 0; j < L; j++) rank += (unsigned)(s_scan[j] >> 32) > sq;
        sel[rank] = (uint32_t)s_scan[i];
    }
    if (threadIdx.x == 0) P.nsel[slot] = L;
}

static size_t quadtree_smem(int NC);
size_t orbx_quadtree_smem(int node_cap) { return quadtree_smem(node_cap); }
static size_t quadtree_smem(int NC)
{
    size_t per = 3 * (3 * sizeof(unsigned) + 4 * sizeof(short))   // cur, nxt, leaf
               + sizeof(unsigned long long) + 4 * sizeof(unsigned) + 4 * sizeof(unsigned short)
               + 2 * sizeof(short) + 2 * sizeof(unsigned short) + sizeof(SrtE);
    return per * (size_t)NC + 64;
}

int launch_quadtree_geo(orbx_handle *h, const FrameGeom *d_geo, int nlevels, int nframes, int node_cap, size_t cand_slab, int sel_slab)
{
    QtParams P;
    P.candA = h->d_cand; P.candB = h->d_cand2; P.ownA = h->d_owner; P.ownB = h->d_owner2; P.tmp = h->d_qtmp;
    P.cand_slab = cand_slab; P.ncand = h->d_ncand;
    P.sel = h->d_sel; P.sel_slab = sel_slab; P.nsel = h->d_nsel; P.status = h->d_status;
    P.node_cap = node_cap;
    const size_t smem = quadtree_smem(node_cap);
    const bool wide = nlevels * nframes <= h->sm_count;                // at most one CTA per SM: latency-bound, use 1024-thread blocks
    if (!orbx_optin_smem(h, (const void *)k_quadtree<256>, smem) || !orbx_optin_smem(h, (const void *)k_quadtree<1024>, smem)) return -1;
    dim3 grid(nframes, nlevels);
    ProfScope ps(h, ORBX_K_QUADTREE);
    if (wide) orbx_launch_pdl(h, k_quadtree<1024>, grid, dim3(1024), smem, h->stream, P, d_geo);
    else {
        if (!orbx_optin_smem(h, (const void *)k_quadtree<QT_BATCH_THREADS>, smem)) return -1;
        orbx_launch_pdl(h, k_quadtree<QT_BATCH_THREADS>, grid, dim3(QT_BATCH_THREADS), smem, h->stream, P, d_geo);
    }
    return 0;
}

int launch_quadtree(orbx_handle *h, int nframes)
{
    return launch_quadtree_geo(h, h->d_geo, h->geo.nlevels, nframes, h->geo.node_cap_max, h->geo.cand_entries, h->geo.sel_entries);
}
