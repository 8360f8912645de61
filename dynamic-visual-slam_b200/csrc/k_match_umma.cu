// k_match_umma.cu — brute-force Hamming matching on the fifth-generation tensor cores (tcgen05.mma kind::i8, accumulators in tensor memory).
// Replaces the same reference code as k_match.cu (cv::BFMatcher(NORM_HAMMING) at frontend.cpp:1123 / :614, the per-pair match() loop of
// Backend::associateObservation backend.cpp:1068-1077) with the same results: integer popcount distance, ties to the lowest trainIdx.
//
// popc(q ^ t) = popc(q) + popc(t) - 2 q.t with the descriptors unpacked to 256 0/1 bytes (k_match_mma's identity).  Here the products run as
// 128 x 128 x 256 tiles: A = 128 queries, staged once per CTA; B = 128 train rows per tile, double buffered; eight tcgen05.mma (M 128, N 128,
// K 32, u8 x u8 -> s32) per tile, issued by one thread, accumulate into one of two 128-column TMEM buffers and signal an mbarrier through
// tcgen05.commit.  While the tensor core works on tile t, the CTA's eight warps read tile t-1 back (tcgen05.ld 32x32b: a thread = one query, 32
// train columns per load), turn each dot product into the packed key (distance << 22 | row) with one multiply-add and keep the best (two) per
// query, then unpack tile t+1.  Operands sit in shared memory in the canonical K-major no-swizzle layout (8-row x 16-byte core matrices): any fixed
// permutation of the 256 bit positions gives the same dot product, so byte k = 32 w + 4 s + b of a row is bit s + 8 b of descriptor word w — one
// (w >> s) & 0x01010101 per four bytes.  tools/umma_probe.cu checks these encodings against a CPU product (and measured 732 cycles per tile even
// with the issue -> commit -> wait chain exposed: 6.3 T pairs/s, against 1.28 T pairs/s of the mma.sync kernel).
// Output = k_match_partial's partial top-2 layout: the epilogue kernel and every caller are unchanged.
#include "orbx_match.h"

#define UM_TILE 128
#define UM_THREADS 256
#define UM_A_BYTES (128 * 256)
#define UM_B_BYTES (UM_TILE * 256)
#define UM_SMEM (UM_A_BYTES + 2 * UM_B_BYTES + 1024)             // + alignment slack; 97 KB (+ 2 KB static): at most two CTAs per SM = 2 x 256 of the 512 TMEM columns
#define UM_DEAD (511u << MT_KEY_SHIFT)

__device__ __forceinline__ uint32_t um_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor, K-major, no swizzle: start address >> 4 | LBO (next 16-byte k chunk: 128 B) | SBO (next 8-row group: 256 B) | version 1
__device__ __forceinline__ uint64_t um_desc(uint32_t saddr) { return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46); }
// instruction descriptor: D = s32 (2 << 4), A = B = u8, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
#define UM_IDESC ((2u << 4) | ((uint32_t)(UM_TILE >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))

// unpack 16 bytes of a descriptor (words 4 half .. 4 half + 3 of row `row`) into the operand tile: word w = k-block, shifts 0-3 -> chunk 0, 4-7 -> chunk 1
__device__ __forceinline__ void um_unpack(uint8_t *tile, int row, int half, const uint4 x)
{
    const uint32_t M1 = 0x01010101u;
    const uint32_t w[4] = { x.x, x.y, x.z, x.w };
    uint8_t *base = tile + (row >> 3) * 256 + (row & 7) * 16;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint8_t *p = base + (4 * half + i) * (128 * 32);
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[i] & M1, (w[i] >> 1) & M1, (w[i] >> 2) & M1, (w[i] >> 3) & M1);
        *reinterpret_cast<uint4 *>(p + 128) = make_uint4((w[i] >> 4) & M1, (w[i] >> 5) & M1, (w[i] >> 6) & M1, (w[i] >> 7) & M1);
    }
}

template <bool TOP2> __global__ void __launch_bounds__(UM_THREADS) k_match_umma(MatchParams P, int32_t *status)
{
    extern __shared__ __align__(16) uint8_t um_raw[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) int32_t s_tk[3][UM_TILE];            // per train row of a tile: popc(t) << 22 | local row  (dead rows: 511 << 22); three tiles are
                                                                  // live at once: t - 1 (being read back), t (in the tensor core), t + 1 (being staged)
    __shared__ uint32_t s_m[2][128];                              // merge of the two column halves at the end
    ORBX_PDL_ENTRY();
    const int prob = blockIdx.z;
    const int qs = P.qsel ? P.qsel[prob] : prob, ts = P.tsel ? P.tsel[prob] : prob;
    const int nq = P.nq_arr ? P.nq_arr[qs] : P.nq_imm;
    const int nt = P.nt_arr ? P.nt_arr[ts] : P.nt_imm;
    if ((int)(blockIdx.x * 128) >= nq) return;                    // before any allocation: the whole CTA leaves
    const int rps = P.nt_arr ? ((((nt + P.nsplit - 1) / P.nsplit) + 7) & ~7) : P.rows_per_split;
    const int r0 = blockIdx.y * rps, r1 = min(nt, r0 + rps);
    const int nrows = max(0, r1 - r0), ntiles = (nrows + UM_TILE - 1) / UM_TILE;
    const uint8_t *qbase = P.q + (size_t)qs * P.q_stride;
    const uint4 *tbase = reinterpret_cast<const uint4 *>(P.t + (size_t)ts * P.t_stride);
    uint8_t *sa = um_raw + ((1024u - (um_smem(um_raw) & 1023u)) & 1023u), *sb = sa + UM_A_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int srow = tid >> 1, shalf = tid & 1;                   // staging role: 16 bytes of one row

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(um_smem(&s_bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(um_smem(&s_bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem(&s_tmem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- A: this CTA's 128 queries (rows past nq: zeros, their results are not written) ----
    int pq;
    {
        const int qrow = blockIdx.x * 128 + srow;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (qrow < nq) x = __ldg(reinterpret_cast<const uint4 *>(qbase + (size_t)qrow * ORBX_DESC_BYTES) + shalf);
        um_unpack(sa, srow, shalf, x);
    }
    // the epilogue role: warp w reads TMEM lanes 32 (w & 3) .. + 31 (a thread = query 32 (w & 3) + lane), columns 64 (w >> 2) .. + 63 of a tile
    const int eq = 32 * (warp & 3) + lane, ehalf = warp >> 2;
    {
        const int qrow = blockIdx.x * 128 + eq;
        const uint4 *qp = reinterpret_cast<const uint4 *>(qbase + (size_t)(qrow < nq ? qrow : 0) * ORBX_DESC_BYTES);
        const uint4 a = __ldg(qp), b = __ldg(qp + 1);
        pq = (__popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w)) << MT_KEY_SHIFT;
    }
    // staging of train tile t into buffer t & 1: 16 bytes per thread, the row's key base from the two halves' popcounts
    auto fetch = [&](int t) {
        const int lrow = t * UM_TILE + srow;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (lrow < nrows) x = __ldg(tbase + (size_t)(r0 + lrow) * 2 + shalf);
        return x;
    };
    auto stage = [&](int t, const uint4 x) {
        um_unpack(sb + (t & 1) * UM_B_BYTES, srow, shalf, x);
        int pc = __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
        pc += __shfl_xor_sync(0xffffffffu, pc, 1);
        const int lrow = t * UM_TILE + srow;
        if (shalf == 0) s_tk[t % 3][srow] = lrow < nrows ? ((pc << MT_KEY_SHIFT) | lrow) : (int32_t)(UM_DEAD | (uint32_t)(lrow & ((1 << MT_KEY_SHIFT) - 1)));
    };
    // generic-proxy writes of the operands -> visible to the tensor core; TMEM address and barriers -> visible to everyone
    auto publish = [&]() {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };
    if (ntiles > 0) stage(0, fetch(0));
    publish();
    const uint32_t tm = s_tmem;
    // signed keys without the query's own popcount (a per-thread constant: added at the end): popc(t) - 2 q.t in [-256, 256] << 22 fits an int
    int m0 = 0x7FFFFFFF, m1 = 0x7FFFFFFF;
    bool failed = false;
    for (int t = 0; t <= ntiles; t++) {
        if (t < ntiles && tid == 0) {                             // tile t -> accumulator buffer t & 1
            const uint32_t a0 = um_smem(sa), b0 = um_smem(sb + (t & 1) * UM_B_BYTES), tc = tm + (uint32_t)((t & 1) * UM_TILE);
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                const uint64_t da = um_desc(a0 + kb * (128 * 32)), db = um_desc(b0 + kb * (UM_TILE * 32));
                const uint32_t acc = kb > 0 ? 1u : 0u;
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}"
                             ::"r"(tc), "l"(da), "l"(db), "r"(UM_IDESC), "r"(acc), "r"(0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(um_smem(&s_bar[t & 1])) : "memory");
        }
        uint4 nx = make_uint4(0u, 0u, 0u, 0u);
        if (t + 1 < ntiles) nx = fetch(t + 1);                    // in flight during the epilogue of tile t - 1
        if (t > 0) {                                              // epilogue of tile t - 1 (its MMAs were committed one iteration ago)
            const int e = t - 1, buf = e & 1;
            const uint32_t parity = (uint32_t)((e >> 1) & 1);
            uint32_t ok = 0;
            for (int spin = 0; spin < (1 << 24) && !ok; spin++)   // bounded: a lost commit must not take the GPU with it
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok) : "r"(um_smem(&s_bar[buf])), "r"(parity) : "memory");
            if (!ok) failed = true;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t v[32];
                const int col0 = 64 * ehalf + 32 * c;
                const uint32_t taddr = tm + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(buf * UM_TILE + col0);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                               "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                               "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int4 *tk4 = reinterpret_cast<const int4 *>(&s_tk[e % 3][col0]);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int4 k4 = tk4[j];                                            // one broadcast load: the key bases of four train rows
                    const int kk[4] = { k4.x, k4.y, k4.z, k4.w };
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int key = (int)((uint32_t)kk[i] - (v[4 * j + i] << (MT_KEY_SHIFT + 1)));     // (popc(t) - 2 q.t) << 22 | row
                        if (TOP2) m1 = min(m1, max(m0, key));
                        m0 = min(m0, key);
                    }
                }
            }
        }
        if (t + 1 < ntiles) stage(t + 1, nx);                     // buffer (t + 1) & 1: tile t - 1's MMAs have read it (their commit was waited for)
        publish();                                                // also orders this iteration's TMEM reads before the next tile's MMAs
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
    if (failed && lane == 0) atomicOr(status, ORBX_DS_INTERNAL);
    // the two warps that hold the two column halves of a query: merge, then write like k_match_partial
    if (ehalf == 1) { s_m[0][eq] = (uint32_t)m0; s_m[1][eq] = (uint32_t)m1; }
    __syncthreads();
    if (ehalf == 0) {
        const int x0 = (int)s_m[0][eq], x1 = (int)s_m[1][eq];
        m1 = min(max(m0, x0), min(m1, x1)); m0 = min(m0, x0);
        const int qrow = blockIdx.x * 128 + eq;
        if (qrow < nq) {
            // back to unsigned keys with the query's popcount: distance << 22 | local row; a dead row carries >= 511 << 22
            const uint32_t u0 = m0 == 0x7FFFFFFF ? MT_INF : (uint32_t)m0 + (uint32_t)pq, u1 = m1 == 0x7FFFFFFF ? MT_INF : (uint32_t)m1 + (uint32_t)pq;
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
