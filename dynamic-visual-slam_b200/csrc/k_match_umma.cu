// k_match_umma.cu — brute-force Hamming matching on the fifth-generation tensor cores (tcgen05.mma kind::i8, accumulators in tensor memory).
// Replaces the same reference code as k_match.cu (cv::BFMatcher(NORM_HAMMING) at frontend.cpp:1123 / :614, the per-pair match() loop of
// Backend::associateObservation backend.cpp:1068-1077) with the same results: integer popcount distance, ties to the lowest trainIdx.
//
// popc(q ^ t) = popc(q) + popc(t) - 2 q.t with the descriptors unpacked to 256 0/1 bytes (k_match_mma's identity).  Here the products run as
// 128 x 128 x 256 tiles: A = 128 queries, staged once per CTA; B = 128 train rows per tile, double buffered; eight tcgen05.mma (M 128, N 128,
// K 32, u8 x u8 -> s32) per tile, issued by one thread, accumulate into one of two 128-column TMEM buffers and signal an mbarrier through
// tcgen05.commit.  Three roles run concurrently on mbarrier rings (orbx_umma.h: um_pipeline): producer warps unpack train tiles, one lane issues
// the MMAs, read-back warps fetch finished accumulators (tcgen05.ld 32x32b: a thread = one query, 32 train columns per load), turn each dot product
// into the packed key (distance << 22 | row) with one multiply-add and keep the best (two) per query.  Operands sit in shared memory in the canonical K-major no-swizzle layout (8-row x 16-byte core matrices): any fixed
// permutation of the 256 bit positions gives the same dot product, so byte k = 32 w + 4 s + b of a row is bit s + 8 b of descriptor word w — one
// (w >> s) & 0x01010101 per four bytes.  tools/umma_probe.cu checks these encodings against a CPU product (and measured 732 cycles per tile even
// with the issue -> commit -> wait chain exposed: 6.3 T pairs/s, against 1.28 T pairs/s of the mma.sync kernel).
// Output = k_match_partial's partial top-2 layout: the epilogue kernel and every caller are unchanged.
#include "orbx_match.h"
#include "orbx_umma.h"

// read-back policy of the matcher: the best (two) keys of the thread's query
template <bool TOP2> struct MatchEpi {
    int m0 = 0x7FFFFFFF, m1 = 0x7FFFFFFF, pq = 0;
    __device__ __forceinline__ void begin(const uint4 a, const uint4 b)
    {
        pq = (__popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w)) << MT_KEY_SHIFT;
    }
    __device__ __forceinline__ void keys8(const int (&k)[8])
    {
        const int kmin = min(min(min(k[0], k[1]), min(k[2], k[3])), min(min(k[4], k[5]), min(k[6], k[7])));
        if (TOP2) {
            // the second best only moves when a key undercuts it — ~2 ln(n) times per query over n rows: test the smallest of eight keys against
            // it (four three-input minima and a compare per eight pairs) and update (best, second best) on that rare path
            if (kmin < m1) {
#pragma unroll
                for (int i = 0; i < 8; i++) { m1 = min(m1, max(m0, k[i])); m0 = min(m0, k[i]); }
            }
        } else m0 = min(m0, kmin);
    }
};

template <bool TOP2> __global__ void __launch_bounds__(UM_THREADS, 2) k_match_umma(MatchParams P, int32_t *status)
{
    extern __shared__ __align__(16) uint8_t um_raw[];
    __shared__ uint32_t s_m[2][128];                              // merge of the two column halves at the end
    ORBX_PDL_ENTRY();
    const int prob = blockIdx.z;
    const int qs = P.qsel ? P.qsel[prob] : prob, ts = P.tsel ? P.tsel[prob] : prob;
    const int nq = P.nq_arr ? P.nq_arr[qs] : P.nq_imm;
    const int nt = P.nt_arr ? P.nt_arr[ts] : P.nt_imm;
    if ((int)(blockIdx.x * 128) >= nq) return;                    // before any allocation: the whole CTA leaves
    const int rps = P.nt_arr ? ((((nt + P.nsplit - 1) / P.nsplit) + 7) & ~7) : P.rows_per_split;
    const int r0 = blockIdx.y * rps, r1 = min(nt, r0 + rps);
    MatchEpi<TOP2> epi;
    const bool ok = um_pipeline(um_raw, P.q + (size_t)qs * P.q_stride, nq, blockIdx.x * 128, reinterpret_cast<const uint4 *>(P.t + (size_t)ts * P.t_stride),
                                r0, max(0, r1 - r0), epi);
    if (!ok && threadIdx.x == 0) atomicOr(status, ORBX_DS_INTERNAL);
    // the two warps that hold the two column halves of a query: merge, then write like k_match_partial
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, eq = 32 * (warp & 3) + lane;
    if (warp >= 12) { s_m[0][eq] = (uint32_t)epi.m0; s_m[1][eq] = (uint32_t)epi.m1; }
    __syncthreads();
    if (warp >= 8 && warp < 12) {
        int m0 = epi.m0, m1 = epi.m1;
        const int x0 = (int)s_m[0][eq], x1 = (int)s_m[1][eq];
        m1 = min(max(m0, x0), min(m1, x1)); m0 = min(m0, x0);
        const int qrow = blockIdx.x * 128 + eq;
        if (qrow < nq) {
            // back to unsigned keys with the query's popcount: distance << 22 | local row; a dead row carries >= 511 << 22
            const uint32_t u0 = m0 == 0x7FFFFFFF ? MT_INF : (uint32_t)m0 + (uint32_t)epi.pq, u1 = m1 == 0x7FFFFFFF ? MT_INF : (uint32_t)m1 + (uint32_t)epi.pq;
            const unsigned long long gb = (unsigned long long)P.row_base + (unsigned long long)r0;
            const uint32_t lowmask = (1u << MT_KEY_SHIFT) - 1;
            unsigned long long *o = P.part + (((size_t)prob * P.nsplit + blockIdx.y) * P.nq_max + qrow) * 2;
            o[0] = u0 >= UM_DEAD ? ~0ull : (((unsigned long long)(u0 >> MT_KEY_SHIFT) << 32) | (gb + (u0 & lowmask)));
            o[1] = (!TOP2 || u1 >= UM_DEAD) ? ~0ull : (((unsigned long long)(u1 >> MT_KEY_SHIFT) << 32) | (gb + (u1 & lowmask)));
        }
    }
}

void launch_match_umma(orbx_handle *h, const MatchParams &P, dim3 grid, bool top2)
{
    auto kern = top2 ? k_match_umma<true> : k_match_umma<false>;
    orbx_optin_smem(h, (const void *)kern, UM_SMEM);
    orbx_launch_pdl(h, kern, grid, dim3(UM_THREADS), UM_SMEM, h->stream, P, h->d_status);
}
