// k_color.cu — cv::cvtColor(BGR2GRAY) on CV_8UC3 (reference frontend.cpp:1084), the step right before the hot path.
// OpenCV's 8-bit path: gray = (3735*B + 19235*G + 9798*R + 16384) >> 15 (pinned against cv2 in the tests).
// Pure streaming: a thread turns 4 pixels (three aligned 32-bit loads) into one 32-bit store; rows of the destination are pitched.
#include "orbx_internal.h"

__global__ void __launch_bounds__(256) k_bgr2gray(const uint8_t *__restrict__ src, size_t sstep, size_t sfstride,
                                                   uint8_t *__restrict__ dst, size_t dstep, size_t dfstride, int w, int h)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, f = blockIdx.z;
    if (x4 >= w) return;
    const uint8_t *s = src + (size_t)f * sfstride + (size_t)y * sstep + 3 * (size_t)x4;
    uint8_t *d = dst + (size_t)f * dfstride + (size_t)y * dstep + x4;
    uint32_t out = 0;
    if (x4 + 3 < w && (((uintptr_t)s) & 3) == 0) {
        const uint32_t a = __ldg(reinterpret_cast<const uint32_t *>(s)), b = __ldg(reinterpret_cast<const uint32_t *>(s) + 1), c = __ldg(reinterpret_cast<const uint32_t *>(s) + 2);
        // bytes: a = B0 G0 R0 B1 | b = G1 R1 B2 G2 | c = R2 B3 G3 R3
        const uint32_t g0 = (3735u * (a & 255u) + 19235u * ((a >> 8) & 255u) + 9798u * ((a >> 16) & 255u) + 16384u) >> 15;
        const uint32_t g1 = (3735u * (a >> 24) + 19235u * (b & 255u) + 9798u * ((b >> 8) & 255u) + 16384u) >> 15;
        const uint32_t g2 = (3735u * ((b >> 16) & 255u) + 19235u * (b >> 24) + 9798u * (c & 255u) + 16384u) >> 15;
        const uint32_t g3 = (3735u * ((c >> 8) & 255u) + 19235u * ((c >> 16) & 255u) + 9798u * (c >> 24) + 16384u) >> 15;
        out = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
    } else {
        for (int i = 0; i < 4 && x4 + i < w; i++)
            out |= ((3735u * s[3 * i] + 19235u * s[3 * i + 1] + 9798u * s[3 * i + 2] + 16384u) >> 15) << (8 * i);
    }
    *reinterpret_cast<uint32_t *>(d) = out;          // destination rows are pitched to a multiple of 4 bytes
}

void launch_bgr2gray(orbx_handle *h, const uint8_t *d_bgr, size_t sstep, size_t sfstride, uint8_t *d_gray, size_t dstep, size_t dfstride,
                     int w, int hgt, int nframes, cudaStream_t st)
{
    dim3 grid((w + 1023) / 1024, hgt, nframes);
    ProfScope ps(h, ORBX_K_OTHER, st);
    k_bgr2gray<<<grid, 256, 0, st>>>(d_bgr, sstep, sfstride, d_gray, dstep, dfstride, w, hgt);
}
