// orbx_umma.h — device pieces shared by the tensor-memory (tcgen05) kernels: k_match_umma.cu (top-2 matcher) and k_assoc.cu (association with
// the reprojection gate).  Operand encodings checked by tools/umma_probe.cu.
#pragma once
#include "orbx_internal.h"

#define UM_TILE 128
#define UM_THREADS 512               // 16 warps: 0-3 unpack train tiles (a thread = one row), 4 issues the MMAs (one lane), 8-15 read the accumulators back
                                     // (warp w reads TMEM lanes 32 (w % 4) .. + 31: a thread = one query; two warps per lane quarter, 64 columns each)
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


// one tcgen05.ld 32x32b.x32: the calling warp's 32 TMEM lanes (a thread = one lane = one query), 32 consecutive 32-bit columns from taddr
#define UM_TMEM_LD32(v, taddr) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), \
                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
                 : "r"(taddr))

// tile t of a CTA: eight MMAs (K = 8 x 32 bytes) of A (128 rows at a0) x B (128 rows at b0) into TMEM columns tc .. tc + 127, then the commit that
// makes `bar` complete a phase when they are done.  One thread calls this.
__device__ __forceinline__ void um_issue_tile(uint32_t a0, uint32_t b0, uint32_t tc, uint64_t *bar)
{
#pragma unroll
    for (int kb = 0; kb < 8; kb++) {
        const uint64_t da = um_desc(a0 + kb * (128 * 32)), db = um_desc(b0 + kb * (UM_TILE * 32));
        const uint32_t acc = kb > 0 ? 1u : 0u;
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}"
                     ::"r"(tc), "l"(da), "l"(db), "r"(UM_IDESC), "r"(acc), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(um_smem(bar)) : "memory");
}
// bounded wait for a phase of `bar` (a lost commit must not take the GPU with it): false = timed out
__device__ __forceinline__ bool um_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 24) && !ok; spin++)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(um_smem(bar)), "r"(parity) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return ok != 0;
}
// generic-proxy writes of the operands -> visible to the tensor core; orders the CTA's TMEM reads before the next MMAs
__device__ __forceinline__ void um_publish()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void um_tmem_alloc(uint32_t *dst, int ncols)          // one warp; ncols a power of two >= 32
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem(dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void um_tmem_free(uint32_t taddr, int ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void um_bar_init(uint64_t *bar, int count = 1)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(um_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void um_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(um_smem(bar)) : "memory");
}
__device__ __forceinline__ void um_commit(uint64_t *bar)            // arrives on `bar` when every tcgen05.mma issued so far by this thread is complete
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(um_smem(bar)) : "memory");
}
// wait of a pipeline role: bounded, and gives up as soon as any role of the CTA has given up (*abort), so that a lost arrival costs one bound, not
// one per tile.  false = the pipeline is broken: leave the role loop.
__device__ __forceinline__ bool um_wait_role(uint64_t *bar, uint32_t parity, volatile int *abort)
{
    uint32_t ok = 0;
    for (int spin = 0; !ok; spin++) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(um_smem(bar)), "r"(parity) : "memory");
        if (!ok && (spin & 1023) == 1023 && (*abort || spin > (1 << 24))) { *abort = 1; return false; }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return true;
}

// ---- the pipeline both kernels run ----
// One CTA: 128 queries (A, unpacked once) against the train rows [r0, r0 + nrows) in tiles of 128 (B, two buffers), two 128-column TMEM accumulators.
// Roles on mbarrier rings, no CTA barrier per tile (a lock-step version spent 30 % of its time at that barrier):
//   producers --b_full[2]--> MMA issuer --acc_full[2]--> read-back warps        (tcgen05.commit arrives on acc_full and on b_empty)
//   producers <--b_empty[2]-- MMA issuer <--acc_empty[2]-- read-back warps
// Key bases (popc(t) << 22 | local row; 511 << 22 for rows past the end) live in a ring of eight tiles: an entry is read up to four tiles after it
// was written.  The read-back hands eight keys at a time to the policy EPI:
//   keys are SIGNED and lack the query's own popcount (a per-thread constant):  key = (popc(t) - 2 q.t) << 22 | local row, in [-2^30, 2^31)
//   EPI::begin(q words)          once per read-back thread (its query's descriptor)
//   EPI::keys8(k[8])             eight consecutive train rows, ascending
// Every wait is bounded and gives up when any role has given up: returns false if the pipeline broke (results are then undefined; no hang).
template <class EPI>
__device__ __forceinline__ bool um_pipeline(uint8_t *um_raw, const uint8_t *qbase, int nq, int q0, const uint4 *tbase, int r0, int nrows, EPI &epi)
{
    __shared__ __align__(8) uint64_t s_bfull[2], s_bempty[2], s_afull[2], s_aempty[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    __shared__ __align__(16) int32_t s_tk[8][UM_TILE];
    const int ntiles = (nrows + UM_TILE - 1) / UM_TILE;
    uint8_t *sa = um_raw + ((1024u - (um_smem(um_raw) & 1023u)) & 1023u), *sb = sa + UM_A_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    volatile int *abort_flag = &s_abort;
    if (tid == 0) {
        for (int i = 0; i < 2; i++) { um_bar_init(&s_bfull[i], 4); um_bar_init(&s_bempty[i], 1); um_bar_init(&s_afull[i], 1); um_bar_init(&s_aempty[i], 8); }
        s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) um_tmem_alloc(&s_tmem, 256);
    // A: the CTA's 128 queries (rows past nq: zeros, their results are not written), 16 bytes per thread of the first eight warps
    if (tid < 256) {
        const int srow = tid >> 1, shalf = tid & 1, qrow = q0 + srow;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (qrow < nq) x = __ldg(reinterpret_cast<const uint4 *>(qbase + (size_t)qrow * ORBX_DESC_BYTES) + shalf);
        um_unpack(sa, srow, shalf, x);
    }
    um_publish();                                                 // operands, barriers and the TMEM address visible to every role
    const uint32_t tm = s_tmem;
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
        // ---- read-back: a thread = query 32 (warp % 4) + lane, columns 64 ((warp >> 2) & 1) .. + 63 of every tile ----
        const int eq = 32 * (warp & 3) + lane, ehalf = (warp >> 2) & 1, qrow = q0 + eq;
        {
            const uint4 *qp = reinterpret_cast<const uint4 *>(qbase + (size_t)(qrow < nq ? qrow : 0) * ORBX_DESC_BYTES);
            epi.begin(__ldg(qp), __ldg(qp + 1));
        }
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
                for (int j = 0; j < 4; j++) {
                    const int4 ka4 = tk4[2 * j], kb4 = tk4[2 * j + 1];             // two broadcast loads: the key bases of eight train rows
                    const uint32_t *vv = &v[8 * j];
                    int k[8];
                    k[0] = (int)((uint32_t)ka4.x - (vv[0] << (MT_KEY_SHIFT + 1))); k[1] = (int)((uint32_t)ka4.y - (vv[1] << (MT_KEY_SHIFT + 1)));
                    k[2] = (int)((uint32_t)ka4.z - (vv[2] << (MT_KEY_SHIFT + 1))); k[3] = (int)((uint32_t)ka4.w - (vv[3] << (MT_KEY_SHIFT + 1)));
                    k[4] = (int)((uint32_t)kb4.x - (vv[4] << (MT_KEY_SHIFT + 1))); k[5] = (int)((uint32_t)kb4.y - (vv[5] << (MT_KEY_SHIFT + 1)));
                    k[6] = (int)((uint32_t)kb4.z - (vv[6] << (MT_KEY_SHIFT + 1))); k[7] = (int)((uint32_t)kb4.w - (vv[7] << (MT_KEY_SHIFT + 1)));
                    epi.keys8(k);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                              // every role is done (or has given up)
    if (warp == 0) um_tmem_free(tm, 256);
    return s_abort == 0;
}
