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
    // split the landmark range so that the grid covers the machine a few times over
    const int qtiles = (nq + AS_THREADS - 1) / AS_THREADS;
    long split = ((long)h->sm_count * 16 + qtiles - 1) / qtiles;
    const long max_split = std::max(1, (nt + AS_TILE - 1) / AS_TILE);
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    long rps = (std::max(nt, 1) + split - 1) / split;
    rps = (rps + AS_TILE - 1) / AS_TILE * AS_TILE;
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
        k_assoc_partial<<<grid, AS_THREADS, 0, h->stream>>>(P);
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
