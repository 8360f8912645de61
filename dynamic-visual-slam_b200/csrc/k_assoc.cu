// k_assoc.cu — Backend::associateObservation as ONE batched call (reference backend.cpp:1064-1120, reprojectPoint :1153-1173).
// For every observation (32-byte descriptor + pixel) against every landmark of the category (32-byte descriptor + float3 position):
//   candidate  iff Hamming distance < max_descriptor_distance_ (50)                                   :1068-1077
//   winner     =   the candidate with the smallest reprojection error cv::norm(obs.pixel - reprojectPoint(X, R, t)),
//                  if that error is < max_reprojection_distance_ (5 px)                               :1091-1111
// The reference issues one 1x1 BFMatcher call per (observation, landmark) pair; here a thread owns one observation in registers,
// landmark descriptors are staged through shared memory (as in k_match.cu) and the fp64 reprojection runs only for the rare
// candidates, with the reference's operation order (R.t()*(X - t) accumulated (a0*b0 + a1*b1) + a2*b2, no FMA; u = (float)(fx*x/z + cx);
// norm in double).  The landmark range is split over gridDim.y; a small merge kernel takes the lexicographic (error, row) minimum over
// the splits — and, for a database sharded over GPUs, over the all-gathered per-shard results.
// Exactly equal errors are resolved to the lowest landmark row (the reference iterates an unordered_map: unspecified there).
#include "orbx_internal.h"
#include "orbx_hamming.h"
#include "orbx_match.h"
#include "orbx_umma.h"
#include <float.h>

#define AS_THREADS 128
#define AS_TILE 256

struct AssocParams {
    const uint8_t *q; const float *qpx; int nq;
    const uint8_t *t; const float *pos; int nt; int rows_per_split; uint32_t row_base;
    float max_dist; double max_err;
    orbx_pose pose;
    orbx_assoc *part;               // [split][nq]
};

__device__ __forceinline__ double reproj_error(const float *p, const orbx_pose &ps, float qx, float qy)
{
    const double d0 = __dsub_rn((double)p[0], ps.t[0]), d1 = __dsub_rn((double)p[1], ps.t[1]), d2 = __dsub_rn((double)p[2], ps.t[2]);
    double pc[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
        pc[i] = __dadd_rn(__dadd_rn(__dmul_rn(ps.R[i], d0), __dmul_rn(ps.R[3 + i], d1)), __dmul_rn(ps.R[6 + i], d2));
    float u = -1.f, v = -1.f;
    if (pc[2] > 0) {
        u = (float)__dadd_rn(__ddiv_rn(__dmul_rn(ps.fx, pc[0]), pc[2]), ps.cx);
        v = (float)__dadd_rn(__ddiv_rn(__dmul_rn(ps.fy, pc[1]), pc[2]), ps.cy);
    }
    const float ex = __fsub_rn(qx, u), ey = __fsub_rn(qy, v);
    return __dsqrt_rn(__dadd_rn(__dmul_rn((double)ex, (double)ex), __dmul_rn((double)ey, (double)ey)));
}

__global__ void __launch_bounds__(AS_THREADS) k_assoc_partial(AssocParams P)
{
    __shared__ uint4 s_t[AS_TILE * 2];
    const int qi = blockIdx.x * AS_THREADS + threadIdx.x;
    const int r0 = blockIdx.y * P.rows_per_split, r1 = min(P.nt, r0 + P.rows_per_split);
    uint32_t q[8];
    {
        const uint4 *qp = reinterpret_cast<const uint4 *>(P.q + (size_t)(qi < P.nq ? qi : 0) * ORBX_DESC_BYTES);
        const uint4 a = __ldg(qp), b = __ldg(qp + 1);
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    const float qx = qi < P.nq ? __ldg(P.qpx + 2 * qi) : 0.f, qy = qi < P.nq ? __ldg(P.qpx + 2 * qi + 1) : 0.f;
    const uint4 *tbase = reinterpret_cast<const uint4 *>(P.t);
    double best = DBL_MAX; int best_j = -1; float best_d = 0.f;
    for (int base = r0; base < r1; base += AS_TILE) {
        const int cnt = min(AS_TILE, r1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 2; i += AS_THREADS) s_t[i] = __ldg(tbase + (size_t)base * 2 + i);
        __syncthreads();
        if (qi >= P.nq) continue;
        for (int j = 0; j < cnt; j++) {
            const uint4 a = s_t[2 * j], b = s_t[2 * j + 1];
            const int d = hamming256(q, a, b);
            if ((float)d < P.max_dist) {
                const double e = reproj_error(P.pos + (size_t)(base + j) * 3, P.pose, qx, qy);
                if (e < P.max_err && e < best) { best = e; best_j = base + j; best_d = (float)d; }   // rows ascend: ties keep the lowest row
            }
        }
    }
    if (qi < P.nq) {
        orbx_assoc r;
        r.reproj_error = best; r.landmark = best_j < 0 ? -1 : (int32_t)(P.row_base + (uint32_t)best_j); r.distance = best_d;
        P.part[(size_t)blockIdx.y * P.nq + qi] = r;
    }
}

// ---- tensor-core variant: the Hamming gate as the exact int8 GEMM of k_match_mma (k_match.cu), the reprojection test in its epilogue ----
// popc(q ^ t) = popc(q) + popc(t) - 2 q.t with the descriptors unpacked to 0/1 bytes; a warp owns 16 observations (A fragments in registers),
// the CTA stages 64 landmark rows at a time in fragment order.  Per pair the epilogue forms the packed key (distance << 22 | local row) with one
// multiply-add; a key below (50 << 22) is a candidate (a handful per observation among a million rows) and only then the fp64 reprojection
// of k_assoc_partial runs.  A quad's four lanes see different landmark columns of the same two observations, rows ascending per lane: their
// results merge by (error, row) like the splits do.  Same partial layout, same merge kernel, bit-identical results.
#define AM_CHUNK 64
#define AM_KEY_SHIFT 22
#define AM_DEAD (511u << AM_KEY_SHIFT)
__device__ __forceinline__ void assoc_imma_16832(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
struct AssocBest { double e; int row; float d; };
__device__ __forceinline__ void assoc_try(AssocBest &b, uint32_t key, uint32_t thr, int r0, const AssocParams &P, float qx, float qy)
{
    if (key < thr) {
        const int row = r0 + (int)(key & ((1u << AM_KEY_SHIFT) - 1));
        const double e = reproj_error(P.pos + (size_t)row * 3, P.pose, qx, qy);
        if (e < P.max_err && e < b.e) { b.e = e; b.row = row; b.d = (float)(key >> AM_KEY_SHIFT); }     // a lane's rows ascend: ties keep the lowest row
    }
}
__global__ void __launch_bounds__(256) k_assoc_mma(AssocParams P)
{
    __shared__ uint4 s_b[2][AM_CHUNK / 8][4][32];                 // [buffer][8-row group][16-byte piece][lane]
    __shared__ uint32_t s_tk[2][AM_CHUNK];
    const int r0 = blockIdx.y * P.rows_per_split, r1 = min(P.nt, r0 + P.rows_per_split);
    const uint4 *tbase = reinterpret_cast<const uint4 *>(P.t);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int qa = blockIdx.x * AS_THREADS + warp * 16 + g, qb = qa + 8;
    const uint32_t M1 = 0x01010101u;
    uint32_t a[8][4], pqa, pqb;
    {
        const uint4 *pa = reinterpret_cast<const uint4 *>(P.q + (size_t)(qa < P.nq ? qa : 0) * ORBX_DESC_BYTES);
        const uint4 *pb = reinterpret_cast<const uint4 *>(P.q + (size_t)(qb < P.nq ? qb : 0) * ORBX_DESC_BYTES);
        const uint4 x0 = __ldg(pa), x1 = __ldg(pa + 1), y0 = __ldg(pb), y1 = __ldg(pb + 1);
        const uint32_t wa[8] = { x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w }, wb[8] = { y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w };
        int ca = 0, cb = 0;
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
            a[ks][0] = (wa[ks] >> t) & M1; a[ks][1] = (wb[ks] >> t) & M1; a[ks][2] = (wa[ks] >> (t + 4)) & M1; a[ks][3] = (wb[ks] >> (t + 4)) & M1;
            ca += __popc(wa[ks]); cb += __popc(wb[ks]);
        }
        pqa = (uint32_t)ca << AM_KEY_SHIFT; pqb = (uint32_t)cb << AM_KEY_SHIFT;
    }
    const float qax = qa < P.nq ? __ldg(P.qpx + 2 * qa) : 0.f, qay = qa < P.nq ? __ldg(P.qpx + 2 * qa + 1) : 0.f;
    const float qbx = qb < P.nq ? __ldg(P.qpx + 2 * qb) : 0.f, qby = qb < P.nq ? __ldg(P.qpx + 2 * qb + 1) : 0.f;
    // (float)d < max_dist for an integer d  <=>  d < ceil(max_dist); distances are at most 256
    const int dlim = max(0, min(257, (int)ceilf(P.max_dist)));
    const uint32_t thr = (uint32_t)dlim << AM_KEY_SHIFT;
    AssocBest ba = { DBL_MAX, -1, 0.f }, bb = { DBL_MAX, -1, 0.f };
    const int nrows = max(0, r1 - r0), nchunks = (nrows + AM_CHUNK - 1) / AM_CHUNK;
    const int G_ = threadIdx.x >> 5, L_ = threadIdx.x & 31;
    auto fetch = [&](int c, uint4 &x0, uint4 &x1) {
        const int lrow = c * AM_CHUNK + G_ * 8 + (L_ >> 2);
        x0 = make_uint4(0u, 0u, 0u, 0u); x1 = x0;
        if (lrow < nrows) { const uint4 *p = tbase + (size_t)(r0 + lrow) * 2; x0 = __ldg(p); x1 = __ldg(p + 1); }
    };
    auto stage = [&](int c, const uint4 x0, const uint4 x1) {
        const int G = G_, L = L_, lrow = c * AM_CHUNK + G * 8 + (L >> 2), tt = L & 3;
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
            s_tk[c & 1][G * 8 + (L >> 2)] = lrow < nrows ? (((uint32_t)pc << AM_KEY_SHIFT) | (uint32_t)lrow) : (AM_DEAD | (uint32_t)(lrow & ((1 << AM_KEY_SHIFT) - 1)));
        }
    };
    uint4 nx0, nx1;
    if (nchunks > 0) { fetch(0, nx0, nx1); stage(0, nx0, nx1); }
    __syncthreads();
    for (int c = 0; c < nchunks; c++) {
        if (c + 1 < nchunks) fetch(c + 1, nx0, nx1);            // in flight during this chunk's tiles
        const int groups = min(AM_CHUNK / 8, (nrows - c * AM_CHUNK + 7) >> 3);
        for (int G = 0; G < groups; G++) {
            const uint4 b0 = s_b[c & 1][G][0][lane], b1 = s_b[c & 1][G][1][lane], b2 = s_b[c & 1][G][2][lane], b3 = s_b[c & 1][G][3][lane];
            int acc[4] = { 0, 0, 0, 0 };
            assoc_imma_16832(acc, a[0], b0.x, b0.y); assoc_imma_16832(acc, a[1], b0.z, b0.w);
            assoc_imma_16832(acc, a[2], b1.x, b1.y); assoc_imma_16832(acc, a[3], b1.z, b1.w);
            assoc_imma_16832(acc, a[4], b2.x, b2.y); assoc_imma_16832(acc, a[5], b2.z, b2.w);
            assoc_imma_16832(acc, a[6], b3.x, b3.y); assoc_imma_16832(acc, a[7], b3.z, b3.w);
            const uint2 tk = *reinterpret_cast<const uint2 *>(&s_tk[c & 1][G * 8 + 2 * t]);      // columns 2t and 2t + 1 of this tile
            const uint32_t two23 = 1u << (AM_KEY_SHIFT + 1);
            const uint32_t k00 = pqa + tk.x - (uint32_t)acc[0] * two23, k01 = pqa + tk.y - (uint32_t)acc[1] * two23;
            const uint32_t k10 = pqb + tk.x - (uint32_t)acc[2] * two23, k11 = pqb + tk.y - (uint32_t)acc[3] * two23;
            if (min(min(k00, k01), min(k10, k11)) < thr) {                                          // rare: a descriptor within the gate
                assoc_try(ba, k00, thr, r0, P, qax, qay); assoc_try(ba, k01, thr, r0, P, qax, qay);
                assoc_try(bb, k10, thr, r0, P, qbx, qby); assoc_try(bb, k11, thr, r0, P, qbx, qby);
            }
        }
        if (c + 1 < nchunks) stage(c + 1, nx0, nx1);
        __syncthreads();
    }
    // the four lanes of a quad hold different landmark columns of the same two observations: lexicographic (error, row) minimum
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        const double ea = __shfl_xor_sync(0xffffffffu, ba.e, o), eb = __shfl_xor_sync(0xffffffffu, bb.e, o);
        const int ra = __shfl_xor_sync(0xffffffffu, ba.row, o), rb = __shfl_xor_sync(0xffffffffu, bb.row, o);
        const float da = __shfl_xor_sync(0xffffffffu, ba.d, o), db = __shfl_xor_sync(0xffffffffu, bb.d, o);
        if (ra >= 0 && (ba.row < 0 || ea < ba.e || (ea == ba.e && ra < ba.row))) { ba.e = ea; ba.row = ra; ba.d = da; }
        if (rb >= 0 && (bb.row < 0 || eb < bb.e || (eb == bb.e && rb < bb.row))) { bb.e = eb; bb.row = rb; bb.d = db; }
    }
    if (t == 0) {
        if (qa < P.nq) {
            orbx_assoc r;
            r.reproj_error = ba.e; r.landmark = ba.row < 0 ? -1 : (int32_t)(P.row_base + (uint32_t)ba.row); r.distance = ba.d;
            P.part[(size_t)blockIdx.y * P.nq + qa] = r;
        }
        if (qb < P.nq) {
            orbx_assoc r;
            r.reproj_error = bb.e; r.landmark = bb.row < 0 ? -1 : (int32_t)(P.row_base + (uint32_t)bb.row); r.distance = bb.d;
            P.part[(size_t)blockIdx.y * P.nq + qb] = r;
        }
    }
}

// ---- tensor-memory variant (tcgen05): the pipeline of k_match_umma (orbx_umma.h: um_pipeline) with the gate + reprojection as its read-back policy ----
// Keys are signed and lack the observation's own popcount, so the gate is key < (ceil(max_dist) << 22) - (popc(q) << 22).  Rows ascend per
// thread; the two threads that hold the two column halves of an observation, and the splits, merge by (error, row).
struct AssocEpi {
    const AssocParams &P; int r0; float qx, qy; int pq, thr;
    AssocBest best;
    __device__ __forceinline__ AssocEpi(const AssocParams &p, int r0_, float qx_, float qy_) : P(p), r0(r0_), qx(qx_), qy(qy_), pq(0), thr(0), best{ DBL_MAX, -1, 0.f } {}
    __device__ __forceinline__ void begin(const uint4 a, const uint4 b)
    {
        pq = (__popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w)) << MT_KEY_SHIFT;
        // (float)d < max_dist for an integer d  <=>  d < ceil(max_dist); distances are at most 256
        thr = (max(0, min(257, (int)ceilf(P.max_dist))) << MT_KEY_SHIFT) - pq;
    }
    __device__ __forceinline__ void candidate(int key)             // a landmark inside the descriptor gate: the fp64 reprojection of k_assoc_partial
    {
        if (key < thr) {
            const int row = r0 + (key & ((1 << MT_KEY_SHIFT) - 1));
            const double e = reproj_error(P.pos + (size_t)row * 3, P.pose, qx, qy);
            if (e < P.max_err && e < best.e) { best.e = e; best.row = row; best.d = (float)(((uint32_t)key + (uint32_t)pq) >> MT_KEY_SHIFT); }   // rows ascend: ties keep the lowest
        }
    }
    // inline on purpose: as an out-of-line call (by reference or by value) the rare path measured 15 % slower — the call constrains the registers of the loop
    __device__ __forceinline__ void keys8(const int (&k)[8])
    {
        const int kmin = min(min(min(k[0], k[1]), min(k[2], k[3])), min(min(k[4], k[5]), min(k[6], k[7])));
        if (kmin < thr) {                                                                              // rare
#pragma unroll
            for (int i = 0; i < 8; i++) candidate(k[i]);
        }
    }
};
__global__ void __launch_bounds__(UM_THREADS, 2) k_assoc_umma(const __grid_constant__ AssocParams P, int32_t *status)
{
    extern __shared__ __align__(16) uint8_t um_raw[];
    __shared__ double s_e[128];
    __shared__ int s_r[128];
    __shared__ float s_d[128];
    const int r0 = blockIdx.y * P.rows_per_split, r1 = min(P.nt, r0 + P.rows_per_split);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, eq = 32 * (warp & 3) + lane, qrow = blockIdx.x * 128 + eq;
    const bool reader = warp >= 8 && qrow < P.nq;
    AssocEpi epi(P, r0, reader ? __ldg(P.qpx + 2 * qrow) : 0.f, reader ? __ldg(P.qpx + 2 * qrow + 1) : 0.f);
    const bool ok = um_pipeline(um_raw, P.q, P.nq, blockIdx.x * 128, reinterpret_cast<const uint4 *>(P.t), r0, max(0, r1 - r0), epi);
    if (!ok && threadIdx.x == 0) atomicOr(status, ORBX_DS_INTERNAL);
    if (warp >= 12) { s_e[eq] = epi.best.e; s_r[eq] = epi.best.row; s_d[eq] = epi.best.d; }
    __syncthreads();
    if (warp >= 8 && warp < 12 && qrow < P.nq) {
        AssocBest best = epi.best;
        const double e1 = s_e[eq]; const int rw1 = s_r[eq];
        if (rw1 >= 0 && (best.row < 0 || e1 < best.e || (e1 == best.e && rw1 < best.row))) { best.e = e1; best.row = rw1; best.d = s_d[eq]; }
        orbx_assoc r;
        r.reproj_error = best.e; r.landmark = best.row < 0 ? -1 : (int32_t)(P.row_base + (uint32_t)best.row); r.distance = best.d;
        P.part[(size_t)blockIdx.y * P.nq + qrow] = r;
    }
}

// lexicographic (error, landmark row) minimum over `nparts` partial results laid out [part][nq]
__global__ void k_assoc_merge(const orbx_assoc *parts, size_t stride, int nparts, int nq, orbx_assoc *out)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    orbx_assoc best; best.reproj_error = DBL_MAX; best.landmark = -1; best.distance = 0.f;
    for (int s = 0; s < nparts; s++) {
        const orbx_assoc p = parts[(size_t)s * stride + qi];
        if (p.landmark < 0) continue;
        if (best.landmark < 0 || p.reproj_error < best.reproj_error || (p.reproj_error == best.reproj_error && p.landmark < best.landmark)) best = p;
    }
    out[qi] = best;
}

int launch_assoc(orbx_handle *h, const uint8_t *d_q, const float *d_qpx, int nq, const uint8_t *d_t, const float *d_pos, int nt, uint32_t row_base,
                 const orbx_pose *pose, float max_dist, double max_err, orbx_assoc *d_out)
{
    if (nq <= 0) return 0;
    // engine (ORBX_OPT_MATCH_MMA, as for the matcher): 3 / 1 = the tensor-memory kernel (always / for calls of >= 8 M pairs), 2 = mma.sync, 0 = POPC
    const bool big = (double)nq * (double)nt >= 8e6;
    const int engine = h->opt_match_mma == 3 || (h->opt_match_mma == 1 && big) ? 3 : (h->opt_match_mma == 2 ? 2 : 0);
    // split the landmark range so that the grid covers the machine a few times over (the tensor-memory kernel: two resident CTAs per SM, 128-row tiles)
    const int qtiles = (nq + AS_THREADS - 1) / AS_THREADS;
    const int unit = engine == 3 ? UM_TILE : AS_TILE;
    long split = ((long)h->sm_count * (engine == 3 ? 4 : 16) + qtiles - 1) / qtiles;
    const long max_split = std::max(1, (nt + unit - 1) / unit);
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    long rps = (std::max(nt, 1) + split - 1) / split;
    rps = (rps + unit - 1) / unit * unit;
    split = (std::max(nt, 1) + rps - 1) / rps;
    const size_t need = (size_t)split * nq * sizeof(orbx_assoc);
    if (need > h->mpart_cap * sizeof(unsigned long long)) {
        if (h->d_mpart) { cudaStreamSynchronize(h->stream); cudaFree(h->d_mpart); }
        h->d_mpart = nullptr; h->mpart_cap = 0;
        if (cudaMalloc(&h->d_mpart, need) != cudaSuccess) return -1;
        h->mpart_cap = need / sizeof(unsigned long long);
    }
    AssocParams P;
    P.q = d_q; P.qpx = d_qpx; P.nq = nq; P.t = d_t; P.pos = d_pos; P.nt = nt; P.rows_per_split = (int)rps; P.row_base = row_base;
    P.max_dist = max_dist; P.max_err = max_err; P.pose = *pose; P.part = (orbx_assoc *)h->d_mpart;
    dim3 grid(qtiles, (unsigned)split);
    {
        ProfScope ps(h, ORBX_K_OTHER);
        if (engine == 3) {
            orbx_optin_smem(h, (const void *)k_assoc_umma, UM_SMEM);
            k_assoc_umma<<<grid, UM_THREADS, UM_SMEM, h->stream>>>(P, h->d_status);
        } else if (engine == 2) k_assoc_mma<<<grid, 256, 0, h->stream>>>(P);
        else k_assoc_partial<<<grid, AS_THREADS, 0, h->stream>>>(P);
    }
    ProfScope ps(h, ORBX_K_OTHER);
    k_assoc_merge<<<(nq + 127) / 128, 128, 0, h->stream>>>((const orbx_assoc *)h->d_mpart, (size_t)nq, (int)split, nq, d_out);
    return 0;
}

void launch_assoc_merge_strided(orbx_handle *h, const orbx_assoc *d_parts, size_t stride, int nparts, int nq, orbx_assoc *d_out)
{
    if (nq <= 0) return;
    ProfScope ps(h, ORBX_K_OTHER);
    k_assoc_merge<<<(nq + 127) / 128, 128, 0, h->stream>>>(d_parts, stride, nparts, nq, d_out);
}
void launch_assoc_merge(orbx_handle *h, const orbx_assoc *d_parts, int nparts, int nq, orbx_assoc *d_out)
{
    launch_assoc_merge_strided(h, d_parts, (size_t)nq, nparts, nq, d_out);
}
