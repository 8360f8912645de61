// imma_probe.cu — go/no-go measurement for a tensor-core Hamming matcher.
// Hamming distance of two 256-bit descriptors = |q| + |t| - 2 q.t with the bits taken as 0/1 integers, so a brute-force matcher is an
// exact integer GEMM.  On sm_100a the binary MMA (mma.sync ... b1 xor.popc) is EMULATED by ptxas (a ~150-instruction routine around
// eight IMMA.16832.U8.U8 per call — see profiles/r02_imma_probe.md), so the candidate is the int8 path itself on descriptors unpacked to bytes.
// This probe measures (1) the issue rate of mma.sync.m16n8k32.s32.u8.u8 per SM (independent accumulators, operands in registers) and
// (2) the same loop with the per-pair epilogue a matcher needs (distance from the dot product + packed-key minimum), against the
// 5-POPC CUDA-core kernel's 790 G pairs/s.        nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_probe imma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <bool EPI> __global__ void __launch_bounds__(256) k_probe(const uint32_t *in, uint32_t *out, int iters)
{
    uint32_t a[8][4], b[2];
    for (int k = 0; k < 8; k++) for (int i = 0; i < 4; i++) a[k][i] = in[(threadIdx.x * 32 + k * 4 + i) & 1023] & 0x01010101u;   // a 16 x 256 query tile, 0/1 bytes
    b[0] = in[threadIdx.x & 1023] & 0x01010101u; b[1] = in[(threadIdx.x + 7) & 1023] & 0x01010101u;
    uint32_t best0 = 0xFFFFFFFFu, best1 = 0xFFFFFFFFu, best2 = 0xFFFFFFFFu, best3 = 0xFFFFFFFFu;
    int sink = 0;
    for (int it = 0; it < iters; it++) {
        int c[4] = { 0, 0, 0, 0 };
#pragma unroll
        for (int k = 0; k < 8; k++) { b[0] ^= (uint32_t)k; imma(c, a[k], b); }                   // one 16 x 8 tile of dot products over K = 256
        if (EPI) {
            // distance = |q| + |t| - 2 dot, packed key (distance << 22 | row), running minimum per query row held by this thread
            const uint32_t base = (uint32_t)it * 8u + (threadIdx.x & 3u) * 2u;
            const uint32_t k0 = ((uint32_t)(120 + 130 - 2 * c[0]) << 22) | base, k1 = ((uint32_t)(120 + 130 - 2 * c[1]) << 22) | (base + 1);
            const uint32_t k2 = ((uint32_t)(120 + 130 - 2 * c[2]) << 22) | base, k3 = ((uint32_t)(120 + 130 - 2 * c[3]) << 22) | (base + 1);
            best0 = min(best0, k0); best1 = min(best1, k1); best2 = min(best2, k2); best3 = min(best3, k3);
        } else sink += c[0] + c[1] + c[2] + c[3];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = EPI ? (best0 ^ best1 ^ best2 ^ best3) : (uint32_t)sink;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    uint32_t *in, *out;
    cudaMalloc(&in, 4096); cudaMemset(in, 0x55, 4096);
    const int grid = p.multiProcessorCount * 8, iters = 20000;
    cudaMalloc(&out, (size_t)grid * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int epi = 0; epi < 2; epi++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (epi) k_probe<true><<<grid, 256>>>(in, out, iters); else k_probe<false><<<grid, 256>>>(in, out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double tiles = (double)grid * 8 * iters;                     // 16 x 8 tiles (8 warps per CTA)
        const double pairs = tiles * 128, ops = tiles * 8 * 16 * 8 * 32 * 2;
        printf("%s: %.3f ms  %.1f G pairs/s  %.1f int8 TOP/s  (%d SMs, err %s)\n", epi ? "imma + top-1 epilogue" : "imma only", ms, pairs / ms / 1e6, ops / ms / 1e9,
               p.multiProcessorCount, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
