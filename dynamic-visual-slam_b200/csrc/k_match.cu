// k_match.cu — brute-force Hamming matching of 256-bit descriptors (k = 1 / 2, threshold and ratio epilogues).
// Replaces cv::BFMatcher(NORM_HAMMING).match at reference frontend.cpp:1123 and :614 (+ the
// `distance < 50` loops :1126-1132, :617-623) and the per-pair match() loop of
// Backend::associateObservation (backend.cpp:1068-1077).  Semantics: SURVEY App. A.8 — integer popcount
// distance, ties to the lowest trainIdx, one DMatch per query in query order.
//
// This is integer-issue-bound work (8 LOP3 + 8 POPC per pair), not a dense FP contraction: no
// tensor cores.  Layout: one thread owns one query descriptor in registers (8 x u32); train rows
// are staged in shared memory in 128-bit pieces and broadcast to the warp; per pair the thread does
// 8 XOR + 8 POPC + 7 IADD and a 3-instruction top-2 update on the packed key
// (distance << 22 | row), whose unsigned order IS the (distance, index) lexicographic order, so the
// lowest-index tie-break needs no extra compare.  The train set is split over gridDim.y; partial
// top-2 keys are merged (lexicographic min, 64-bit keys with global row numbers) by the epilogue
// kernel, which also applies the threshold / ratio test and compacts in query order.
#include "orbx_internal.h"
#include "orbx_hamming.h"
#include <float.h>

#include "orbx_match.h"

// TOP2 = false (k = 1 without a ratio test or raw top-2 output): only the best row is tracked, one VIMNMX per pair instead of three
template <bool TOP2> __global__ void __launch_bounds__(MT_THREADS) k_match_partial(MatchParams P)
{
    __shared__ uint4 s_t[MT_TILE * 2];
    ORBX_PDL_ENTRY();
    const int prob = blockIdx.z;
    const int qs = P.qsel ? P.qsel[prob] : prob, ts = P.tsel ? P.tsel[prob] : prob;
    const int nq = P.nq_arr ? P.nq_arr[qs] : P.nq_imm;
    const int nt = P.nt_arr ? P.nt_arr[ts] : P.nt_imm;
    const int qi = blockIdx.x * MT_THREADS + threadIdx.x;
    if (blockIdx.x * MT_THREADS >= nq) return;
    // the train set is cut into gridDim.y equal parts of the problem's OWN row count (multiple of 8 rows)
    const int rps = P.nt_arr ? ((((nt + P.nsplit - 1) / P.nsplit) + 7) & ~7) : P.rows_per_split;
    const int r0 = blockIdx.y * rps;
    const int r1 = min(nt, r0 + rps);
    const uint8_t *qbase = P.q + (size_t)qs * P.q_stride;
    const uint4 *tbase = reinterpret_cast<const uint4 *>(P.t + (size_t)ts * P.t_stride);
    uint32_t q[8];
    {
        const uint4 *qp = reinterpret_cast<const uint4 *>(qbase + (size_t)(qi < nq ? qi : 0) * ORBX_DESC_BYTES);
        const uint4 a = __ldg(qp), b = __ldg(qp + 1);
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    uint32_t m0 = MT_INF, m1 = MT_INF;
    for (int base = r0; base < r1; base += MT_TILE) {
        const int cnt = min(MT_TILE, r1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 2; i += MT_THREADS) s_t[i] = __ldg(tbase + (size_t)base * 2 + i);
        __syncthreads();
        uint32_t key = (uint32_t)(base - r0);
#pragma unroll 4
        for (int j = 0; j < cnt; j++, key++) {
            const int d = hamming256(q, s_t[2 * j], s_t[2 * j + 1]);
            const uint32_t kk = ((uint32_t)d << MT_KEY_SHIFT) + key;
            if (TOP2) m1 = min(m1, max(m0, kk));
            m0 = min(m0, kk);
        }
    }
    if (qi < nq) {
        unsigned long long *o = P.part + (((size_t)prob * P.nsplit + blockIdx.y) * P.nq_max + qi) * 2;
        const unsigned long long gb = (unsigned long long)P.row_base + (unsigned long long)r0;
        o[0] = m0 == MT_INF ? ~0ull : (((unsigned long long)(m0 >> MT_KEY_SHIFT) << 32) | (gb + (m0 & ((1u << MT_KEY_SHIFT) - 1))));
        o[1] = m1 == MT_INF ? ~0ull : (((unsigned long long)(m1 >> MT_KEY_SHIFT) << 32) | (gb + (m1 & ((1u << MT_KEY_SHIFT) - 1))));
    }
}

// ---- tensor-core variant: Hamming distance as an exact integer GEMM ----
// popc(q ^ t) = popc(q) + popc(t) - 2 popc(q & t), and popc(q & t) is the dot product of the two descriptors read as 256 0/1 integers.
// sm_100a has no native binary MMA (ptxas emulates mma.sync ... b1.xor.popc with ~150 instructions around eight IMMA.16832.U8.U8), but the int8
// path itself is usable: this kernel runs the 2048 x 1M association at 1.28 T pairs/s against 0.81 T pairs/s of k_match_partial (1.6x) and the
// frame matcher at 0.090 ms against 0.129 ms per 128-frame step, bit-identical.  (tools/imma_probe.cu over-estimated the pipe: operands that never
// change let it report 4.98 T pairs/s; the real kernel shows the tensor pipe 58 % busy at 0.28 IMMA/clk/SM.)  Any fixed permutation of the 256 bit positions gives the same dot product, so the operands are unpacked the cheapest way:
// (w >> s) & 0x01010101 turns bits {s, s+8, s+16, s+24} of a descriptor word into one register of four 0/1 bytes.  k-step ks of the MMA takes
// word ks; thread t of a quad supplies shift t for k = 4t..4t+3 and shift t+4 for k = 16+4t..16+4t+3, on the query (A) and the train (B) side alike.
//   * a warp owns 16 queries: A fragments (8 k-steps x 4 registers) stay in registers for the whole kernel;
//   * the CTA (8 warps = 128 queries, the same tiling as k_match_partial) stages 64 train rows at a time: each thread unpacks one (row group,
//     lane) slice once, in fragment order, into shared memory (double buffered, one barrier per chunk), together with the packed key base
//     (popc(t) << 22 | local row);  every warp then reads its B registers with four LDS.128 per 8-row tile;
//   * per 16 x 8 tile: 8 IMMA, then per pair ONE integer multiply-add forms the packed key (distance << 22 | row) = base + (popc(q) << 22) - dot << 23
//     and the usual min/max keeps the best (two) per query; rows past the end carry distance 511 and never win.
// Output = k_match_partial's partial top-2 layout: the epilogue and every caller are unchanged, results are bit-identical.
#define MM_CHUNK 64
#define MM_DEAD (511u << MT_KEY_SHIFT)
__device__ __forceinline__ void imma_16832(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool TOP2> __global__ void __launch_bounds__(256) k_match_mma(MatchParams P)
{
    __shared__ uint4 s_b[2][MM_CHUNK / 8][4][32];                 // [buffer][8-row group][16-byte piece][lane]: 32 KB, conflict-free LDS.128 / STS.128
    __shared__ uint32_t s_tk[2][MM_CHUNK];
    ORBX_PDL_ENTRY();
    const int prob = blockIdx.z;
    const int qs = P.qsel ? P.qsel[prob] : prob, ts = P.tsel ? P.tsel[prob] : prob;
    const int nq = P.nq_arr ? P.nq_arr[qs] : P.nq_imm;
    const int nt = P.nt_arr ? P.nt_arr[ts] : P.nt_imm;
    if (blockIdx.x * MT_THREADS >= nq) return;
    const int rps = P.nt_arr ? ((((nt + P.nsplit - 1) / P.nsplit) + 7) & ~7) : P.rows_per_split;
    const int r0 = blockIdx.y * rps;
    const int r1 = min(nt, r0 + rps);
    const uint8_t *qbase = P.q + (size_t)qs * P.q_stride;
    const uint4 *tbase = reinterpret_cast<const uint4 *>(P.t + (size_t)ts * P.t_stride);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int qa = blockIdx.x * MT_THREADS + warp * 16 + g, qb = qa + 8;            // the two query rows whose results this quad holds
    const uint32_t M1 = 0x01010101u;
    uint32_t a[8][4], pqa, pqb;
    {
        const uint4 *pa = reinterpret_cast<const uint4 *>(qbase + (size_t)(qa < nq ? qa : 0) * ORBX_DESC_BYTES);
        const uint4 *pb = reinterpret_cast<const uint4 *>(qbase + (size_t)(qb < nq ? qb : 0) * ORBX_DESC_BYTES);
        const uint4 x0 = __ldg(pa), x1 = __ldg(pa + 1), y0 = __ldg(pb), y1 = __ldg(pb + 1);
        const uint32_t wa[8] = { x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w }, wb[8] = { y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w };
        int ca = 0, cb = 0;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
            a[ks][0] = (wa[ks] >> t) & M1; a[ks][1] = (wb[ks] >> t) & M1; a[ks][2] = (wa[ks] >> (t + 4)) & M1; a[ks][3] = (wb[ks] >> (t + 4)) & M1;
            ca += __popc(wa[ks]); cb += __popc(wb[ks]);
        }
        pqa = (uint32_t)ca << MT_KEY_SHIFT; pqb = (uint32_t)cb << MT_KEY_SHIFT;
    }
    uint32_t m0a = MT_INF, m1a = MT_INF, m0b = MT_INF, m1b = MT_INF;
    const int nrows = max(0, r1 - r0), nchunks = (nrows + MM_CHUNK - 1) / MM_CHUNK;
    // staging of chunk c into buffer c & 1: this thread owns the slice of (group G, lane L) = threadIdx.x.  The two global loads are issued
    // BEFORE the tiles of the current chunk are computed and unpacked AFTER them, so their latency hides behind the MMAs.
    const int G_ = threadIdx.x >> 5, L_ = threadIdx.x & 31;
    auto fetch = [&](int c, uint4 &x0, uint4 &x1) {
        const int lrow = c * MM_CHUNK + G_ * 8 + (L_ >> 2);
        x0 = make_uint4(0u, 0u, 0u, 0u); x1 = x0;
        if (lrow < nrows) { const uint4 *p = tbase + (size_t)(r0 + lrow) * 2; x0 = __ldg(p); x1 = __ldg(p + 1); }
    };
    auto stage = [&](int c, const uint4 x0, const uint4 x1) {
        const int G = G_, L = L_, lrow = c * MM_CHUNK + G * 8 + (L >> 2), tt = L & 3;
        const bool valid = lrow < nrows;
        const uint32_t w[8] = { x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w };
        uint32_t r[16];
#pragma unroll
        for (int ks = 0; ks < 8; ks++) { r[2 * ks] = (w[ks] >> tt) & M1; r[2 * ks + 1] = (w[ks] >> (tt + 4)) & M1; }
        s_b[c & 1][G][0][L] = make_uint4(r[0], r[1], r[2], r[3]); s_b[c & 1][G][1][L] = make_uint4(r[4], r[5], r[6], r[7]);
        s_b[c & 1][G][2][L] = make_uint4(r[8], r[9], r[10], r[11]); s_b[c & 1][G][3][L] = make_uint4(r[12], r[13], r[14], r[15]);
        if (tt == 0) {
            int pc = 0;
#pragma unroll
            for (int ks = 0; ks < 8; ks++) pc += __popc(w[ks]);
            s_tk[c & 1][G * 8 + (L >> 2)] = valid ? (((uint32_t)pc << MT_KEY_SHIFT) | (uint32_t)lrow) : (MM_DEAD | (uint32_t)(lrow & ((1 << MT_KEY_SHIFT) - 1)));
        }
    };
    uint4 nx0, nx1;
    if (nchunks > 0) { fetch(0, nx0, nx1); stage(0, nx0, nx1); }
    __syncthreads();
    for (int c = 0; c < nchunks; c++) {
        if (c + 1 < nchunks) fetch(c + 1, nx0, nx1);            // in flight during this chunk's tiles
        const int groups = min(MM_CHUNK / 8, (nrows - c * MM_CHUNK + 7) >> 3);
        for (int G = 0; G < groups; G++) {
            const uint4 b0 = s_b[c & 1][G][0][lane], b1 = s_b[c & 1][G][1][lane], b2 = s_b[c & 1][G][2][lane], b3 = s_b[c & 1][G][3][lane];
            // one chain of eight MMAs: two chains of four measured no faster (1.76 vs 1.68 ms on 2048 x 1M) — ncu shows the tensor pipe itself
            // 58 % busy at 0.28 IMMA/clk/SM: legacy mma.sync int8 peaks near 0.48 IMMA/clk/SM (~1.1 POP/s) on B200, a quarter of tcgen05's rate
            int acc[4] = { 0, 0, 0, 0 };
            imma_16832(acc, a[0], b0.x, b0.y); imma_16832(acc, a[1], b0.z, b0.w);
            imma_16832(acc, a[2], b1.x, b1.y); imma_16832(acc, a[3], b1.z, b1.w);
            imma_16832(acc, a[4], b2.x, b2.y); imma_16832(acc, a[5], b2.z, b2.w);
            imma_16832(acc, a[6], b3.x, b3.y); imma_16832(acc, a[7], b3.z, b3.w);
            const uint2 tk = *reinterpret_cast<const uint2 *>(&s_tk[c & 1][G * 8 + 2 * t]);      // columns 2t and 2t + 1 of this tile
            const uint32_t two23 = 1u << (MT_KEY_SHIFT + 1);
            const uint32_t k00 = pqa + tk.x - (uint32_t)acc[0] * two23, k01 = pqa + tk.y - (uint32_t)acc[1] * two23;
            const uint32_t k10 = pqb + tk.x - (uint32_t)acc[2] * two23, k11 = pqb + tk.y - (uint32_t)acc[3] * two23;
            if (TOP2) { m1a = min(m1a, max(m0a, k00)); m0a = min(m0a, k00); m1a = min(m1a, max(m0a, k01)); m0a = min(m0a, k01);
                        m1b = min(m1b, max(m0b, k10)); m0b = min(m0b, k10); m1b = min(m1b, max(m0b, k11)); m0b = min(m0b, k11); }
            else { m0a = min(m0a, min(k00, k01)); m0b = min(m0b, min(k10, k11)); }
        }
        if (c + 1 < nchunks) stage(c + 1, nx0, nx1);            // the other buffer: everyone left it at the barrier that ended chunk c - 1
        __syncthreads();
    }
    // the four lanes of a quad hold different columns of the same two query rows: merge their (best, second best)
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        const uint32_t x0 = __shfl_xor_sync(0xffffffffu, m0a, o), x1 = __shfl_xor_sync(0xffffffffu, m1a, o);
        const uint32_t y0 = __shfl_xor_sync(0xffffffffu, m0b, o), y1 = __shfl_xor_sync(0xffffffffu, m1b, o);
        m1a = min(max(m0a, x0), min(m1a, x1)); m0a = min(m0a, x0);
        m1b = min(max(m0b, y0), min(m1b, y1)); m0b = min(m0b, y0);
    }
    if (t == 0) {
        const unsigned long long gb = (unsigned long long)P.row_base + (unsigned long long)r0;
        const uint32_t lowmask = (1u << MT_KEY_SHIFT) - 1;
        if (qa < nq) {
            unsigned long long *o = P.part + (((size_t)prob * P.nsplit + blockIdx.y) * P.nq_max + qa) * 2;
            o[0] = m0a >= MM_DEAD ? ~0ull : (((unsigned long long)(m0a >> MT_KEY_SHIFT) << 32) | (gb + (m0a & lowmask)));
            o[1] = (!TOP2 || m1a >= MM_DEAD) ? ~0ull : (((unsigned long long)(m1a >> MT_KEY_SHIFT) << 32) | (gb + (m1a & lowmask)));
        }
        if (qb < nq) {
            unsigned long long *o = P.part + (((size_t)prob * P.nsplit + blockIdx.y) * P.nq_max + qb) * 2;
            o[0] = m0b >= MM_DEAD ? ~0ull : (((unsigned long long)(m0b >> MT_KEY_SHIFT) << 32) | (gb + (m0b & lowmask)));
            o[1] = (!TOP2 || m1b >= MM_DEAD) ? ~0ull : (((unsigned long long)(m1b >> MT_KEY_SHIFT) << 32) | (gb + (m1b & lowmask)));
        }
    }
}

struct MatchEpiParams {
    const unsigned long long *part; int nq_max, nsplit;
    const int32_t *nq_arr; int nq_imm; const int32_t *nt_arr; int nt_imm;
    const int32_t *qsel, *tsel;
    int k; float max_dist; int ratio_num;       // ratio = ratio_num / 1024, 0 = off
    orbx_dmatch *out; size_t out_stride;         // entries per problem
    int32_t *n_out;
    orbx_top2 *top2;                             // nullable: raw per-query top-2 instead of DMatch
};

__device__ __forceinline__ void top2_insert(unsigned long long &a, unsigned long long &b, unsigned long long v)
{
    // keep the two smallest of {a, b, v}, a <= b
    const unsigned long long hi = a > v ? a : v;
    a = a < v ? a : v;
    b = b < hi ? b : hi;
}

__global__ void __launch_bounds__(1024) k_match_epilogue(MatchEpiParams P)
{
    __shared__ int s_warp[33];                              // one CTA per problem, any block size that is a multiple of 32
    __shared__ int s_base;
    ORBX_PDL_ENTRY();
    const int prob = blockIdx.x;
    const int qs = P.qsel ? P.qsel[prob] : prob, ts = P.tsel ? P.tsel[prob] : prob;
    const int nq = P.nq_arr ? P.nq_arr[qs] : P.nq_imm;
    const int nt = P.nt_arr ? P.nt_arr[ts] : P.nt_imm;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    orbx_dmatch *out = P.out ? P.out + (size_t)prob * P.out_stride : nullptr;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    // the modes without compaction (top-2 records, knnMatch pairs) are independent per query: gridDim.y CTAs share a problem's queries
    // (a 2048-query database lookup merged by ONE CTA took 0.1 ms — a third of the 8-GPU query); the compacting modes keep one CTA
    for (int base = blockIdx.y * blockDim.x; base < nq; base += blockDim.x * gridDim.y) {
        const int qi = base + threadIdx.x;
        unsigned long long a = ~0ull, b = ~0ull;
        if (qi < nq && nt > 0) {
            for (int s = 0; s < P.nsplit; s++) {
                const unsigned long long *p = P.part + (((size_t)prob * P.nsplit + s) * P.nq_max + qi) * 2;
                top2_insert(a, b, p[0]);
                top2_insert(a, b, p[1]);
            }
        }
        if (P.top2) {
            if (qi < nq) {
                orbx_top2 r;
                r.dist0 = a == ~0ull ? 0xFFFFFFFFu : (uint32_t)(a >> 32); r.idx0 = (uint32_t)a;
                r.dist1 = b == ~0ull ? 0xFFFFFFFFu : (uint32_t)(b >> 32); r.idx1 = (uint32_t)b;
                P.top2[(size_t)prob * P.nq_max + qi] = r;
            }
            continue;
        }
        const int d0 = (int)(a >> 32), d1 = (int)(b >> 32);
        const bool has0 = a != ~0ull, has1 = b != ~0ull;
        if (P.k == 2 && P.ratio_num <= 0) {                 // knnMatch(k=2): two entries per query
            if (qi < nq) {
                orbx_dmatch m;
                m.queryIdx = qi; m.imgIdx = 0;
                m.trainIdx = has0 ? (int)(uint32_t)a : -1; m.distance = has0 ? (float)d0 : 0.f; out[2 * qi] = m;
                m.trainIdx = has1 ? (int)(uint32_t)b : -1; m.distance = has1 ? (float)d1 : 0.f; out[2 * qi + 1] = m;
            }
            continue;
        }
        bool keep = qi < nq && has0;
        if (keep && P.max_dist > 0.f) keep = (float)d0 < P.max_dist;
        if (keep && P.k == 2) keep = has1 && (d0 * 1024 < P.ratio_num * d1);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) { int run = 0; for (int w = 0; w < nw; w++) { const int c = s_warp[w]; s_warp[w] = run; run += c; } s_warp[32] = run; }
        __syncthreads();
        if (keep) {
            const int o = s_base + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
            orbx_dmatch m;
            m.queryIdx = qi; m.trainIdx = (int)(uint32_t)a; m.imgIdx = 0; m.distance = (float)d0;
            out[o] = m;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_warp[32];
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.y == 0 && P.n_out) {
        if (P.top2) P.n_out[prob] = nq;
        else if (P.k == 2 && P.ratio_num <= 0) P.n_out[prob] = 2 * nq;
        else P.n_out[prob] = s_base;
    }
}

// choose the train split so the grid covers the machine a few times over
static int pick_split(orbx_handle *h, int nq_max, int nt_max, int nproblems, int *rows_per_split)
{
    const int qtiles = (nq_max + MT_THREADS - 1) / MT_THREADS;
    const long ctas_wanted = (long)h->sm_count * 24;
    long split = (ctas_wanted + (long)qtiles * nproblems - 1) / ((long)qtiles * nproblems);
    const long max_split = (nt_max + 63) / 64;
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    long rps = (nt_max + split - 1) / split;
    rps = (rps + MT_TILE - 1) / MT_TILE * MT_TILE;
    if (rps > (1 << MT_KEY_SHIFT)) rps = 1 << MT_KEY_SHIFT;
    split = (nt_max + rps - 1) / rps;
    if (split < 1) split = 1;
    *rows_per_split = (int)rps;
    return (int)split;
}

// the tensor-memory kernel: 128-row tiles, two resident CTAs per SM, each should see many tiles
static int pick_split_umma(const orbx_handle *h, int nq_max, int nt_max, int nproblems, int *rows_per_split)
{
    const int qtiles = (nq_max + MT_THREADS - 1) / MT_THREADS;
    const long ctas_wanted = (long)h->sm_count * 4;
    long split = (ctas_wanted + (long)qtiles * nproblems - 1) / ((long)qtiles * nproblems);
    const long max_split = (nt_max + 127) / 128;
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    long rps = (nt_max + split - 1) / split;
    rps = (rps + 127) / 128 * 128;
    if (rps > (1 << MT_KEY_SHIFT)) rps = 1 << MT_KEY_SHIFT;
    split = (nt_max + rps - 1) / rps;
    if (split < 1) split = 1;
    *rows_per_split = (int)rps;
    return (int)split;
}

static int ensure_part(orbx_handle *h, size_t entries)
{
    if (entries <= h->mpart_cap) return 0;
    if (h->d_mpart) cudaFree(h->d_mpart);
    h->d_mpart = nullptr; h->mpart_cap = 0;
    if (cudaMalloc(&h->d_mpart, entries * sizeof(unsigned long long)) != cudaSuccess) return -1;
    h->mpart_cap = entries;
    return 0;
}

int launch_match_core(orbx_handle *h, const uint8_t *d_q, const int32_t *d_nq, int nq_max, size_t q_stride,
                      const uint8_t *d_t, const int32_t *d_nt, int nt_max, size_t t_stride,
                      const int32_t *d_qsel, const int32_t *d_tsel, int nproblems, uint32_t row_base,
                      int k, float max_dist, int ratio_num, orbx_dmatch *d_out, size_t out_stride, int32_t *d_n_out,
                      orbx_top2 *d_top2)
{
    if (nproblems <= 0 || nq_max <= 0) return 0;
    int rps = MT_TILE;
    // engine (ORBX_OPT_MATCH_MMA): 0 = POPC, 1 = the tensor-memory kernel once there is enough work to amortise its staging (a single frame pair, the
    // latency path of ~1 M pairs, stays on the POPC kernel), 2 = the mma.sync kernel, 3 = the tensor-memory kernel for every call
    const bool big = (double)nproblems * nq_max * nt_max >= 8e6;
    const int engine = h->opt_match_mma == 3 || (h->opt_match_mma == 1 && big) ? 3 : (h->opt_match_mma == 2 ? 2 : 0);
    const int nsplit = nt_max <= 0 ? 1 : (engine == 3 ? pick_split_umma(h, nq_max, nt_max, nproblems, &rps) : pick_split(h, nq_max, nt_max, nproblems, &rps));
    if (ensure_part(h, (size_t)nproblems * nsplit * nq_max * 2) != 0) return -1;
    if (nt_max > 0) {
        MatchParams P;
        P.q = d_q; P.nq_arr = d_nq; P.nq_imm = nq_max; P.q_stride = q_stride;
        P.t = d_t; P.nt_arr = d_nt; P.nt_imm = nt_max; P.t_stride = t_stride;
        P.qsel = d_qsel; P.tsel = d_tsel;
        P.part = (unsigned long long *)h->d_mpart; P.nq_max = nq_max; P.nsplit = nsplit; P.rows_per_split = rps;
        P.row_base = row_base;
        dim3 grid((nq_max + MT_THREADS - 1) / MT_THREADS, nsplit, nproblems);
        // every slot the epilogue reads (qi < nq, all splits) is written by the partial kernel; splits that
        // start beyond a problem's own nt write the "empty" key
        ProfScope ps(h, ORBX_K_MATCH);
        if (engine == 3) launch_match_umma(h, P, grid, k == 2 || d_top2 != nullptr);
        else if (engine == 2) {
            if (k == 2 || d_top2) orbx_launch_pdl(h, k_match_mma<true>, grid, dim3(256), 0, h->stream, P);
            else orbx_launch_pdl(h, k_match_mma<false>, grid, dim3(256), 0, h->stream, P);
        } else if (k == 2 || d_top2) orbx_launch_pdl(h, k_match_partial<true>, grid, dim3(MT_THREADS), 0, h->stream, P);
        else orbx_launch_pdl(h, k_match_partial<false>, grid, dim3(MT_THREADS), 0, h->stream, P);
    }
    MatchEpiParams E;
    E.part = (const unsigned long long *)h->d_mpart; E.nq_max = nq_max; E.nsplit = nsplit;
    E.nq_arr = d_nq; E.nq_imm = nq_max; E.nt_arr = d_nt; E.nt_imm = nt_max;
    E.qsel = d_qsel; E.tsel = d_tsel;
    E.k = k; E.max_dist = max_dist; E.ratio_num = ratio_num;
    E.out = d_out; E.out_stride = out_stride; E.n_out = d_n_out; E.top2 = d_top2;
    ProfScope ps(h, ORBX_K_MATCH_EPI);
    const bool independent = d_top2 != nullptr || (k == 2 && ratio_num <= 0);
    if (independent) orbx_launch_pdl(h, k_match_epilogue, dim3(nproblems, (nq_max + 127) / 128), dim3(128), 0, h->stream, E);
    else orbx_launch_pdl(h, k_match_epilogue, dim3(nproblems), dim3(1024), 0, h->stream, E);
    return 0;
}

// ---- merge of per-shard top-2 after the all-gather ----
__global__ void k_merge_top2(const orbx_top2 *parts, size_t stride, int nshards, int nq, orbx_top2 *out)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    unsigned long long a = ~0ull, b = ~0ull;
    for (int s = 0; s < nshards; s++) {
        const orbx_top2 p = parts[(size_t)s * stride + qi];
        if (p.dist0 != 0xFFFFFFFFu) top2_insert(a, b, ((unsigned long long)p.dist0 << 32) | p.idx0);
        if (p.dist1 != 0xFFFFFFFFu) top2_insert(a, b, ((unsigned long long)p.dist1 << 32) | p.idx1);
    }
    orbx_top2 r;
    r.dist0 = a == ~0ull ? 0xFFFFFFFFu : (uint32_t)(a >> 32); r.idx0 = (uint32_t)a;
    r.dist1 = b == ~0ull ? 0xFFFFFFFFu : (uint32_t)(b >> 32); r.idx1 = (uint32_t)b;
    out[qi] = r;
}
// parts laid out [shard][stride entries]: stride = nq after an all-gather, the mailbox slot size after the peer-memory exchange
void launch_merge_top2_strided(orbx_handle *h, const orbx_top2 *d_parts, size_t stride, int nshards, int nq, orbx_top2 *d_out)
{
    if (nq <= 0) return;
    k_merge_top2<<<(nq + 127) / 128, 128, 0, h->stream>>>(d_parts, stride, nshards, nq, d_out);
    h->launches++;
}
void launch_merge_top2(orbx_handle *h, const orbx_top2 *d_parts, int nshards, int nq, orbx_top2 *d_out)
{
    launch_merge_top2_strided(h, d_parts, (size_t)nq, nshards, nq, d_out);
}

// ---- radius query: every (query, row) with distance < max_dist (backend.cpp:1074-1076) ----
// one thread per query over a row split; hits are appended through a global counter and sorted by
// (queryIdx, trainIdx) on the host side of the ABI (the hit list is tiny: candidates within 50 bits).
struct RadiusParams { const uint8_t *q; int nq; const uint8_t *t; int nt; int rows_per_split; uint32_t row_base;
                      float max_dist; orbx_dmatch *out; int cap; int32_t *n_out; };
__global__ void __launch_bounds__(MT_THREADS) k_match_radius(RadiusParams P)
{
    __shared__ uint4 s_t[MT_TILE * 2];
    const int qi = blockIdx.x * MT_THREADS + threadIdx.x;
    const int r0 = blockIdx.y * P.rows_per_split, r1 = min(P.nt, r0 + P.rows_per_split);
    uint32_t q[8];
    {
        const uint4 *qp = reinterpret_cast<const uint4 *>(P.q + (size_t)(qi < P.nq ? qi : 0) * ORBX_DESC_BYTES);
        const uint4 a = __ldg(qp), b = __ldg(qp + 1);
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    const uint4 *tbase = reinterpret_cast<const uint4 *>(P.t);
    for (int base = r0; base < r1; base += MT_TILE) {
        const int cnt = min(MT_TILE, r1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 2; i += MT_THREADS) s_t[i] = __ldg(tbase + (size_t)base * 2 + i);
        __syncthreads();
        if (qi >= P.nq) continue;
        for (int j = 0; j < cnt; j++) {
            const int d = hamming256(q, s_t[2 * j], s_t[2 * j + 1]);
            if ((float)d < P.max_dist) {
                const int o = atomicAdd(P.n_out, 1);
                if (o < P.cap) { orbx_dmatch m; m.queryIdx = qi; m.trainIdx = (int)(P.row_base + (uint32_t)(base + j)); m.imgIdx = 0; m.distance = (float)d; P.out[o] = m; }
            }
        }
    }
}
void launch_match_radius(orbx_handle *h, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, uint32_t row_base,
                         float max_dist, orbx_dmatch *d_out, int cap, int32_t *d_n_out)
{
    if (nq <= 0 || nt <= 0) return;
    int rps;
    const int nsplit = pick_split(h, nq, nt, 1, &rps);
    RadiusParams P = { d_q, nq, d_t, nt, rps, row_base, max_dist, d_out, cap, d_n_out };
    dim3 grid((nq + MT_THREADS - 1) / MT_THREADS, nsplit);
    k_match_radius<<<grid, MT_THREADS, 0, h->stream>>>(P);
    h->launches++;
}

// ---- POPC issue-rate microbenchmark: the denominator of the matching roofline ----
__global__ void k_popc_bench(uint32_t *out, int iters)
{
    uint32_t a = threadIdx.x * 2654435761u + blockIdx.x, b = a ^ 0x9E3779B9u, c = a + 0x7F4A7C15u, d = ~a;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            s0 += __popc(a ^ s1); s1 += __popc(b ^ s2); s2 += __popc(c ^ s3); s3 += __popc(d ^ s0);
        }
    }
    if ((s0 ^ s1 ^ s2 ^ s3) == 0x12345678u) out[0] = s0;
}
double run_popc_bench(orbx_handle *h)
{
    uint32_t *d = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) return 0.0;
    const int iters = 4096, blocks = h->sm_count * 8, threads = 256;
    k_popc_bench<<<blocks, threads, 0, h->stream>>>(d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, h->stream);
    k_popc_bench<<<blocks, threads, 0, h->stream>>>(d, iters);
    cudaEventRecord(e1, h->stream);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    h->launches += 2;
    const double popc = (double)blocks * threads * (double)iters * 32.0;
    return ms > 0 ? popc / (ms * 1e-3) : 0.0;
}
