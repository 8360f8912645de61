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
#include "orbx_umma.h"

#define UM_THREADS_X 512             // 16 warps: 0-3 unpack train tiles (a thread = one row), 4 issues the MMAs (one lane), 8-15 read the accumulators back
                                     // (warp w reads TMEM lanes 32 (w % 4) .. + 31: a thread = one query; two warps per lane quarter, 64 columns each)

// Roles on mbarrier rings, no CTA barrier per tile (the lock-step version spent 30 % of its time there):
//   producers --b_full[2]--> MMA issuer --acc_full[2]--> read-back warps        (tcgen05.commit arrives on acc_full and on b_empty)
//   producers <--b_empty[2]-- MMA issuer <--acc_empty[2]-- read-back warps
// key bases (popc(t) << 22 | row) live in a ring of eight tiles: a tile's entry is read by the read-back warps up to four tiles after it was written.
template <bool TOP2> __global__ void __launch_bounds__(UM_THREADS_X, 2) k_match_umma(MatchParams P, int32_t *status)
{
    extern __shared__ __align__(16) uint8_t um_raw[];
    __shared__ __align__(8) uint64_t s_bfull[2], s_bempty[2], s_afull[2], s_aempty[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    __shared__ __align__(16) int32_t s_tk[8][UM_TILE];
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
    volatile int *abort_flag = &s_abort;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) { um_bar_init(&s_bfull[i], 4); um_bar_init(&s_bempty[i], 1); um_bar_init(&s_afull[i], 1); um_bar_init(&s_aempty[i], 8); }
        s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) um_tmem_alloc(&s_tmem, 256);
    // ---- A: this CTA's 128 queries (rows past nq: zeros, their results are not written), 16 bytes per thread of the first eight warps ----
    if (tid < 256) {
        const int srow = tid >> 1, shalf = tid & 1, qrow = blockIdx.x * 128 + srow;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (qrow < nq) x = __ldg(reinterpret_cast<const uint4 *>(qbase + (size_t)qrow * ORBX_DESC_BYTES) + shalf);
        um_unpack(sa, srow, shalf, x);
    }
    um_publish();                                                 // operands, barriers and the TMEM address visible to every role
    const uint32_t tm = s_tmem;
    int m0 = 0x7FFFFFFF, m1 = 0x7FFFFFFF, pq = 0;
    const int eq = 32 * (warp & 3) + lane, ehalf = (warp >> 2) & 1;

    if (warp < 4) {
        // ---- producers: thread = train row `tid` of every tile ----
        const int row = tid;
        uint4 x0 = make_uint4(0u, 0u, 0u, 0u), x1 = x0;
        if (row < nrows) { const uint4 *p = tbase + (size_t)(r0 + row) * 2; x0 = __ldg(p); x1 = __ldg(p + 1); }
        for (int t = 0; t < ntiles; t++) {
            const int s = t & 1, lrow = t * UM_TILE + row, nlrow = lrow + UM_TILE;
            uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;                  // the next tile's row: in flight while this one is unpacked
            if (t + 1 < ntiles && nlrow < nrows) { const uint4 *p = tbase + (size_t)(r0 + nlrow) * 2; n0 = __ldg(p); n1 = __ldg(p + 1); }
            if (t >= 2 && !um_wait_role(&s_bempty[s], (uint32_t)(((t - 2) >> 1) & 1), abort_flag)) break;      // the MMAs of tile t - 2 have read this buffer
            uint8_t *tile = sb + s * UM_B_BYTES;
            um_unpack(tile, row, 0, x0); um_unpack(tile, row, 1, x1);
            const int pc = __popc(x0.x) + __popc(x0.y) + __popc(x0.z) + __popc(x0.w) + __popc(x1.x) + __popc(x1.y) + __popc(x1.z) + __popc(x1.w);
            s_tk[t & 7][row] = lrow < nrows ? ((pc << MT_KEY_SHIFT) | lrow) : (int32_t)(UM_DEAD | (uint32_t)(lrow & ((1 << MT_KEY_SHIFT) - 1)));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // this thread's operand bytes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) um_arrive(&s_bfull[s]);
            x0 = n0; x1 = n1;
        }
    } else if (warp == 4) {
        // ---- MMA issuer: one lane ----
        if (lane == 0) {
            for (int t = 0; t < ntiles; t++) {
                const int s = t & 1;
                if (!um_wait_role(&s_bfull[s], (uint32_t)((t >> 1) & 1), abort_flag)) break;
                if (t >= 2 && !um_wait_role(&s_aempty[s], (uint32_t)(((t - 2) >> 1) & 1), abort_flag)) break;     // tile t - 2 has been read back
                const uint32_t a0 = um_smem(sa), b0 = um_smem(sb + s * UM_B_BYTES), tc = tm + (uint32_t)(s * UM_TILE);
#pragma unroll
                for (int kb = 0; kb < 8; kb++) {
                    const uint64_t da = um_desc(a0 + kb * (128 * 32)), db = um_desc(b0 + kb * (UM_TILE * 32));
                    const uint32_t acc = kb > 0 ? 1u : 0u;
                    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}"
                                 ::"r"(tc), "l"(da), "l"(db), "r"(UM_IDESC), "r"(acc), "r"(0u) : "memory");
                }
                um_commit(&s_bempty[s]);
                um_commit(&s_afull[s]);
            }
        }
    } else if (warp >= 8) {
        // ---- read-back: a thread = query eq, columns 64 ehalf .. + 63 of every tile ----
        {
            const int qrow = blockIdx.x * 128 + eq;
            const uint4 *qp = reinterpret_cast<const uint4 *>(qbase + (size_t)(qrow < nq ? qrow : 0) * ORBX_DESC_BYTES);
            const uint4 a = __ldg(qp), b = __ldg(qp + 1);
            pq = (__popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w)) << MT_KEY_SHIFT;
        }
        // signed keys without the query's own popcount (a per-thread constant: added at the end): popc(t) - 2 q.t in [-256, 256] << 22 fits an int
        for (int t = 0; t < ntiles; t++) {
            const int s = t & 1;
            if (!um_wait_role(&s_afull[s], (uint32_t)((t >> 1) & 1), abort_flag)) break;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                uint32_t v[32];
                const int col0 = 64 * ehalf + 32 * c;
                UM_TMEM_LD32(v, tm + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(s * UM_TILE + col0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c == 1) {                                                      // both loads done: the accumulator buffer may be overwritten
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) um_arrive(&s_aempty[s]);
                }
                const int4 *tk4 = reinterpret_cast<const int4 *>(&s_tk[t & 7][col0]);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int4 k4 = tk4[j];                                        // one broadcast load: the key bases of four train rows
                    const uint32_t *vv = &v[4 * j];
                    // (popc(t) - 2 q.t) << 22 | row
                    const int ka = (int)((uint32_t)k4.x - (vv[0] << (MT_KEY_SHIFT + 1))), kb = (int)((uint32_t)k4.y - (vv[1] << (MT_KEY_SHIFT + 1)));
                    const int kc = (int)((uint32_t)k4.z - (vv[2] << (MT_KEY_SHIFT + 1))), kd = (int)((uint32_t)k4.w - (vv[3] << (MT_KEY_SHIFT + 1)));
                    if (TOP2) {
                        // the second best only moves when a key undercuts it — ~2 ln(n) times per query over n rows: test the smallest of four keys
                        // against it (two three-input minima and a compare per four pairs) and update (best, second best) on that rare path
                        const int k4min = min(min(ka, kb), min(kc, kd));
                        if (k4min < m1) {
                            m1 = min(m1, max(m0, ka)); m0 = min(m0, ka);
                            m1 = min(m1, max(m0, kb)); m0 = min(m0, kb);
                            m1 = min(m1, max(m0, kc)); m0 = min(m0, kc);
                            m1 = min(m1, max(m0, kd)); m0 = min(m0, kd);
                        }
                    } else m0 = min(min(m0, min(ka, kb)), min(kc, kd));
                }
            }
        }
        if (ehalf == 1) { s_m[0][eq] = (uint32_t)m0; s_m[1][eq] = (uint32_t)m1; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                              // every role is done (or has given up)
    if (warp == 0) um_tmem_free(tm, 256);
    if (s_abort && tid == 0) atomicOr(status, ORBX_DS_INTERNAL);
    // the two warps that hold the two column halves of a query: merge, then write like k_match_partial
    if (warp >= 8 && ehalf == 0) {
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
    orbx_launch_pdl(h, kern, grid, dim3(UM_THREADS_X), UM_SMEM, h->stream, P, h->d_status);
}
