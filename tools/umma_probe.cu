// umma_probe.cu — checks the tcgen05.mma kind::i8 operand encodings the tensor-memory matcher relies on (shared-memory matrix descriptors for
// K-major operands without swizzle, the instruction descriptor, the TMEM accumulator layout) against a CPU product, then times the MMA issue rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/build/umma_probe tools/umma_probe.cu     (run on the GPU box)
// C[m][n] = sum_k A[m][k] * B[n][k], A: 128 x 256 u8, B: 128 x 256 u8 (0/1 bytes in the matcher), C: s32 in TMEM lane m, column n.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define M_ 128
#define N_ 128
#define K_ 256
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major, no swizzle: core matrix = 8 rows x 16 bytes, rows 16 bytes apart; element (row, k) of a [rows x 256 B] operand split in eight 32-byte
// k-blocks:  kb * (rows * 32) + (row / 8) * 256 + ((k % 32) / 16) * 128 + (row % 8) * 16 + k % 16   ->  LBO (k chunk) = 128 B, SBO (row group) = 256 B
__host__ __device__ inline int umma_off(int rows, int row, int k) { return (k >> 5) * (rows * 32) + (row >> 3) * 256 + ((k >> 4) & 1) * 128 + (row & 7) * 16 + (k & 15); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) { return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (8ull << 16) | (16ull << 32) | (1ull << 46); }
// instruction descriptor: c_format S32 (2) at bit 4, a/b format UINT8 (0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ inline uint32_t umma_idesc(int m, int n) { return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__global__ void __launch_bounds__(128) k_probe(const uint8_t *A, const uint8_t *B, int32_t *C, int reps, long long *cycles, int *err)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint8_t *sa = smem, *sb = smem + M_ * K_;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M_ * K_; i += 128) { const int row = i / K_, k = i % K_; sa[umma_off(M_, row, k)] = A[i]; }
    for (int i = tid; i < N_ * K_; i += 128) { const int row = i / K_, k = i % K_; sb[umma_off(N_, row, k)] = B[i]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes of the operands -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    const uint32_t idesc = umma_idesc(M_, N_);
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < reps; r++) {
        if (tid == 0) {
            for (int kb = 0; kb < K_ / 32; kb++) {
                const uint64_t da = umma_desc(smem_u32(sa) + kb * (M_ * 32)), db = umma_desc(smem_u32(sb) + kb * (N_ * 32));
                const uint32_t acc = kb > 0 ? 1u : 0u;
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}"
                             ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        // bounded wait: a wrong encoding must not hang the GPU
        uint32_t ok = 0;
        for (int spin = 0; spin < (1 << 22) && !ok; spin++)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        if (!ok) { if (tid == 0) atomicExch(err, 1); break; }
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long t1 = clock64();
    if (tid == 0) *cycles = t1 - t0;
    // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31, 32 columns at a time
    for (int c0 = 0; c0 < N_; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                       "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; j++) C[tid * N_ + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128) : "memory");
}

// variant: A in tensor memory (tcgen05.mma with [tmem_a]): thread = row m writes its 256 bytes as 64 32-bit columns (column c = bytes 4c .. 4c+3 of the row)
#define ST32(taddr, v) \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" \
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), \
                    "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory")
__global__ void __launch_bounds__(128) k_probe_ts(const uint8_t *A, const uint8_t *B, int32_t *C, int reps, long long *cycles, int *err)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint8_t *sb = smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < N_ * K_; i += 128) { const int row = i / K_, k = i % K_; sb[umma_off(N_, row, k)] = B[i]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base, ta = tm + 128u;
    {
        const uint32_t *arow = reinterpret_cast<const uint32_t *>(A + (size_t)tid * K_);
        uint32_t v[32];
        for (int h = 0; h < 2; h++) {
            for (int j = 0; j < 32; j++) v[j] = arow[32 * h + j];
            ST32(ta + ((uint32_t)(warp * 32) << 16) + (uint32_t)(32 * h), v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = umma_idesc(M_, N_);
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < reps; r++) {
        if (tid == 0) {
            for (int kb = 0; kb < K_ / 32; kb++) {
                const uint64_t db = umma_desc(smem_u32(sb) + kb * (N_ * 32));
                const uint32_t acc = kb > 0 ? 1u : 0u;
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}"
                             ::"r"(tm), "r"(ta + (uint32_t)(kb * 8)), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        uint32_t ok = 0;
        for (int spin = 0; spin < (1 << 22) && !ok; spin++)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        if (!ok) { if (tid == 0) atomicExch(err, 1); break; }
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long t1 = clock64();
    if (tid == 0) *cycles = t1 - t0;
    for (int c0 = 0; c0 < N_; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                       "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; j++) C[tid * N_ + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
}

int main()
{
    std::vector<uint8_t> A(M_ * K_), B(N_ * K_);
    srand(7);
    for (auto &x : A) x = rand() & 1;
    for (auto &x : B) x = rand() & 1;
    for (int k = 0; k < K_; k++) { A[5 * K_ + k] = (uint8_t)(k & 3); B[9 * K_ + k] = (uint8_t)(200 + (k & 7)); }      // a few general u8 values
    uint8_t *dA, *dB; int32_t *dC; long long *dcyc; int *derr;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dC, M_ * N_ * 4); cudaMalloc(&dcyc, 8); cudaMalloc(&derr, 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xFF, M_ * N_ * 4); cudaMemset(derr, 0, 4);
    const int smem = (M_ + N_) * K_ + 1024;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int reps : { 1, 2000 }) {
        k_probe<<<1, 128, smem>>>(dA, dB, dC, reps, dcyc, derr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
        int err = 0; long long cyc = 0;
        cudaMemcpy(&err, derr, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
        if (err) { printf("mbarrier wait timed out (commit never arrived)\n"); return 3; }
        std::vector<int32_t> C(M_ * N_);
        cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < M_; m++) for (int n = 0; n < N_; n++) {
            int s = 0;
            for (int k = 0; k < K_; k++) s += (int)A[m * K_ + k] * (int)B[n * K_ + k];
            if (s != C[m * N_ + n]) { if (bad < 5) printf("  C[%d][%d] = %d, want %d\n", m, n, C[m * N_ + n], s); bad++; }
        }
        printf("reps %d: %d mismatches of %d; %lld cycles -> %.1f cycles per 128x128x256 tile (%.2f T pairs/s per SM-clock-GHz x 148 SMs at 1.9 GHz: %.2f T pairs/s)\n",
               reps, bad, M_ * N_, cyc, (double)cyc / reps, 0.0, 148.0 * 1.9e9 * reps * M_ * N_ / (double)cyc / 1e12);
        if (bad) return 1;
    }
    // A in tensor memory
    cudaFuncSetAttribute(k_probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int reps : { 1, 2000 }) {
        cudaMemset(dC, 0xFF, M_ * N_ * 4); cudaMemset(derr, 0, 4);
        k_probe_ts<<<1, 128, smem>>>(dA, dB, dC, reps, dcyc, derr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("TS: CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
        int err = 0; long long cyc = 0;
        cudaMemcpy(&err, derr, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
        if (err) { printf("TS: mbarrier wait timed out\n"); return 3; }
        std::vector<int32_t> C(M_ * N_);
        cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < M_; m++) for (int n = 0; n < N_; n++) {
            int s = 0;
            for (int k = 0; k < K_; k++) s += (int)A[m * K_ + k] * (int)B[n * K_ + k];
            if (s != C[m * N_ + n]) { if (bad < 5) printf("  TS C[%d][%d] = %d, want %d\n", m, n, C[m * N_ + n], s); bad++; }
        }
        printf("A in TMEM, reps %d: %d mismatches of %d; %.1f cycles per tile\n", reps, bad, M_ * N_, (double)cyc / reps);
    }
    return 0;
}
