// ubench.cu — issue-rate microbenchmarks of the integer instructions the FAST / blur kernels lean on (sm_100a).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{
    if (OP == 0) return __vabsdiffu4(a, b);
    if (OP == 1) return __vminu2(a, b);
    if (OP == 2) return __vminu2(__vminu2(a, b), c);          // VIMNMX3.U16x2
    if (OP == 3) return __byte_perm(a, b, c);
    if (OP == 4) return (a | b) & c;                            // LOP3
    if (OP == 5) return a + b;                                  // IADD3 / IMAD.IADD
    if (OP == 6) return a * b + c;                              // IMAD
    if (OP == 7) return min(a, b);                              // VIMNMX.U32
    if (OP == 8) return min(min(a, b), c);                      // VIMNMX3.U32
    if (OP == 9) return __dp4a(a, b, c);                        // IDP.4A
    if (OP == 10) return __popc(a ^ b);                         // LOP3 + POPC
    if (OP == 11) return __funnelshift_r(a, b, 8);              // SHF
    if (OP == 12) return __vsadu4(a, b) + c;                    // VABSDIFF4 accumulate
    if (OP == 13) return __umulhi(a, b) + c;                    // IMAD.HI
    if (OP == 14) return __dp2a_lo(a, b, c);                    // IDP.2A
    return a;
}
template <int OP> __global__ void k(uint32_t *out, uint32_t seed)
{
    uint32_t x0 = threadIdx.x * 2654435761u + seed, x1 = x0 ^ 0x9E3779B9u, x2 = x0 + 0x7F4A7C15u, x3 = ~x0;
    uint32_t y0 = x1 * 3, y1 = x2 * 5, y2 = x3 * 7, y3 = x0 * 11;
    const uint32_t c = seed | 0x01010101u;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            x0 = op<OP>(x0, y0, c); x1 = op<OP>(x1, y1, c); x2 = op<OP>(x2, y2, c); x3 = op<OP>(x3, y3, c);
            y0 = op<OP>(y0, x1, c); y1 = op<OP>(y1, x2, c); y2 = op<OP>(y2, x3, c); y3 = op<OP>(y3, x0, c);
        }
    }
    if ((x0 ^ x1 ^ x2 ^ x3 ^ y0 ^ y1 ^ y2 ^ y3) == 0x12345678u) out[0] = x0;
}
__global__ void k_lds(uint32_t *out, int width)
{
    __shared__ uint32_t s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = i * 7u;
    __syncthreads();
    uint32_t acc = 0; int idx = threadIdx.x;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (width == 1) acc += ((volatile uint8_t *)s)[(idx * 4 + u * 1031 + i) & 16383];
            else acc += ((volatile uint32_t *)s)[(idx + u * 257 + i) & 4095];
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int OP> static void run(const char *name, uint32_t *d, int sms)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8, threads = 256;
    k<OP><<<blocks, threads>>>(d, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<OP><<<blocks, threads>>>(d, 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * ITERS * 32.0;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.1f Gop/s  = %6.2f lanes/clk/SM at %d MHz nominal\n", name, ops / ms / 1e6, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *d; cudaMalloc(&d, 4);
    run<0>("VABSDIFF4.U8", d, sms); run<12>("VABSDIFF4 + acc (vsadu4)", d, sms);
    run<1>("VIMNMX.U16x2", d, sms); run<2>("VIMNMX3.U16x2", d, sms);
    run<7>("VIMNMX.U32", d, sms); run<8>("VIMNMX3.U32", d, sms);
    run<3>("PRMT", d, sms); run<4>("LOP3", d, sms); run<5>("IADD", d, sms); run<6>("IMAD", d, sms);
    run<9>("IDP.4A", d, sms); run<13>("IMAD.HI", d, sms); run<14>("IDP.2A", d, sms); run<10>("LOP3+POPC", d, sms); run<11>("SHF", d, sms);
    for (int w = 1; w <= 4; w += 3) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k_lds<<<sms * 8, 256>>>(d, w); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_lds<<<sms * 8, 256>>>(d, w); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)sms * 8 * 256 * ITERS * 8.0;
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        printf("LDS.%s (+IADD, addr math)      %8.1f Gld/s  = %6.2f lanes/clk/SM\n", w == 1 ? "U8 " : "32 ", ops / ms / 1e6, ops / (ms * 1e-3) / sms / (clk * 1e3));
    }
    return 0;
}
