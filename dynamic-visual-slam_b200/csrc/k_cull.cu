// k_cull.cu — the frontend's feature culling for the backend (reference frontend.cpp:1168-1218), one frame per launch.
//   backend set = the keypoint of every geometrically consistent match, in match order (:1181-1190), followed by the unmatched
//   keypoints (collected in index order, :1193-1198) sorted by response with std::sort and the comparator `a.first > b.first`
//   (:1201-1202), taken while fewer than MAX_NEW_FEATURES = 200 were added and response >= MIN_RESPONSE = 50 (:1205-1219).
// FAST responses are small integers, so the sort is tie-heavy and WHICH 200 features survive depends on where libstdc++'s introsort
// leaves equal elements: the sort is the exact block-parallel restatement of orbx_sort.h (shared with the quadtree).
// One CTA: the whole job is a few thousand elements and sits on the latency path of a keyframe, not on the throughput path.
#include "orbx_internal.h"
#include "orbx_sort.h"

#define CULL_THREADS 256

struct CullParams {
    const orbx_keypoint *kps; const uint8_t *desc; int n;
    const int32_t *mq; int nm;
    int max_new; float min_response;
    orbx_keypoint *out_kps; uint8_t *out_desc; int32_t *out_index; int cap;
    int32_t *n_out; int32_t *status;
    int range_cap;                       // capacity of the sort's range lists
};

// descending response as an ascending 32-bit key (-0 counts as +0, like the float comparator)
__device__ __forceinline__ unsigned cull_key(float response)
{
    const unsigned u = __float_as_uint(response + 0.0f);
    return ~(u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u));
}

__device__ __forceinline__ void cull_copy(const CullParams &P, int dst, int src)
{
    // one keypoint = 7 words, one descriptor = 8 words: 15 word moves per element, spread over the block by the caller
    const uint32_t *sk = reinterpret_cast<const uint32_t *>(P.kps + src), *sd = reinterpret_cast<const uint32_t *>(P.desc + (size_t)src * ORBX_DESC_BYTES);
    uint32_t *dk = reinterpret_cast<uint32_t *>(P.out_kps + dst), *dd = reinterpret_cast<uint32_t *>(P.out_desc + (size_t)dst * ORBX_DESC_BYTES);
#pragma unroll
    for (int w = 0; w < 7; w++) dk[w] = sk[w];
#pragma unroll
    for (int w = 0; w < 8; w++) dd[w] = sd[w];
    if (P.out_index) P.out_index[dst] = src;
}

__global__ void __launch_bounds__(CULL_THREADS) k_cull(CullParams P)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_cnt[2], s_warp[CULL_THREADS / 32], s_base, s_bad, s_nge;
    const int n = P.n, nm = P.nm;
    unsigned char *sp = s_raw;
    SrtE *a = (SrtE *)sp;                     sp += sizeof(SrtE) * n;
    SrtE *tmp = (SrtE *)sp;                   sp += sizeof(SrtE) * n;
    unsigned *lists = (unsigned *)sp;         sp += sizeof(unsigned) * 4 * P.range_cap;
    unsigned *flags = (unsigned *)sp;         sp += sizeof(unsigned) * ((n + 31) / 32);
    unsigned short *Ls = (unsigned short *)sp; sp += sizeof(unsigned short) * n;
    unsigned short *Ra = (unsigned short *)sp;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int i = tid; i < (n + 31) / 32; i += CULL_THREADS) flags[i] = 0u;
    if (tid == 0) { s_base = 0; s_bad = 0; s_nge = 0; }
    __syncthreads();
    // matched_indices (std::set<int>, :1175-1178)
    for (int i = tid; i < nm; i += CULL_THREADS) {
        const int q = P.mq[i];
        if (q < 0 || q >= n) s_bad = 1;
        else atomicOr(&flags[q >> 5], 1u << (q & 31));
    }
    __syncthreads();
    if (s_bad) {
        if (tid == 0) { atomicOr(P.status, ORBX_DS_BAD_INDEX); *P.n_out = 0; }
        return;
    }
    // priority 1: every match's keypoint, in match order
    for (int i = tid; i < nm; i += CULL_THREADS) if (i < P.cap) cull_copy(P, i, P.mq[i]);
    // unmatched (response, index) pairs in index order
    for (int i0 = 0; i0 < n; i0 += CULL_THREADS) {
        const int i = i0 + tid;
        const bool un = i < n && !((flags[i >> 5] >> (i & 31)) & 1u);
        const unsigned bal = __ballot_sync(0xffffffffu, un);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < wid; w++) off += s_warp[w];
        if (un) a[off + __popc(bal & ((1u << lane) - 1u))] = ((SrtE)cull_key(P.kps[i].response) << 32) | (SrtE)(unsigned)i;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < CULL_THREADS / 32; w++) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    const int nu = s_base;
    block_sort_exact<CULL_THREADS>(a, nu, tmp, Ls, Ra, lists, P.range_cap, s_cnt);
    // sorted descending: the loop of :1209-1219 takes min(max_new, #(response >= min_response)) elements
    int cnt = 0;
    for (int i = tid; i < nu; i += CULL_THREADS) cnt += !(P.kps[srt_pay(a[i])].response < P.min_response);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) atomicAdd(&s_nge, cnt);
    __syncthreads();
    // (a NaN response compares false either way: it neither stops the reference's loop nor sorts anywhere definite — not produced by FAST)
    const int n_new = max(0, min(P.max_new, s_nge));
    for (int j = tid; j < n_new; j += CULL_THREADS) if (nm + j < P.cap) cull_copy(P, nm + j, srt_pay(a[j]));
    if (tid == 0) {
        *P.n_out = nm + n_new;
        if (nm + n_new > P.cap) atomicOr(P.status, ORBX_DS_KP_OVERFLOW);
    }
}

size_t cull_smem_bytes(int n, int range_cap)
{
    return (size_t)n * (2 * sizeof(SrtE) + 2 * sizeof(unsigned short)) + sizeof(unsigned) * (4 * (size_t)range_cap + (size_t)(n + 31) / 32) + 64;
}

int launch_cull(orbx_handle *h, const orbx_keypoint *d_kps, const uint8_t *d_desc, int n, const int32_t *d_mq, int nm, int max_new, float min_response,
                orbx_keypoint *d_out_kps, uint8_t *d_out_desc, int32_t *d_out_index, int cap, int32_t *d_n_out)
{
    CullParams P;
    P.kps = d_kps; P.desc = d_desc; P.n = n; P.mq = d_mq; P.nm = nm; P.max_new = max_new; P.min_response = min_response;
    P.out_kps = d_out_kps; P.out_desc = d_out_desc; P.out_index = d_out_index; P.cap = cap; P.n_out = d_n_out; P.status = h->d_status;
    P.range_cap = n / 16 + 4;
    const size_t smem = cull_smem_bytes(n, P.range_cap);
    if (!orbx_optin_smem(h, (const void *)k_cull, smem)) return -1;
    ProfScope ps(h, ORBX_K_OTHER);
    k_cull<<<1, CULL_THREADS, smem, h->stream>>>(P);
    return 0;
}
