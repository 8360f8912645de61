// k_synth.cu — seeded synthetic inputs generated in HBM (bench / tests).  Integer-only so the bytes are
// identical to the oracle's host generator (oracle/orb_oracle.c, orc_synth_*): world-anchored value
// noise + hashed rectangles, frame f = world shifted by (4f, 2f) px; depth in mm with dropout blocks;
// landmark-database rows from a counter-based hash.  SURVEY §8(d).
#include "orbx_internal.h"

__device__ __forceinline__ uint32_t syn_hash(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t h = a * 0x9E3779B1u ^ b * 0x85EBCA77u ^ c * 0xC2B2AE3Du;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}
__device__ __forceinline__ uint32_t syn_vnoise(uint32_t X, uint32_t Y, int shift, uint32_t seed)
{
    const uint32_t Pw = 1u << shift, gx = X >> shift, gy = Y >> shift, fx = X & (Pw - 1), fy = Y & (Pw - 1);
    const uint32_t v00 = syn_hash(gx, gy, seed) & 255, v10 = syn_hash(gx + 1, gy, seed) & 255;
    const uint32_t v01 = syn_hash(gx, gy + 1, seed) & 255, v11 = syn_hash(gx + 1, gy + 1, seed) & 255;
    const uint32_t top = v00 * (Pw - fx) + v10 * fx, bot = v01 * (Pw - fx) + v11 * fx;
    return (top * (Pw - fy) + bot * fy) >> (2 * shift);
}
__device__ __forceinline__ uint8_t syn_gray_px(uint32_t seed, uint32_t X, uint32_t Y)
{
    const uint32_t acc = 8 * syn_vnoise(X, Y, 6, seed + 1) + 4 * syn_vnoise(X, Y, 5, seed + 2) +
                         2 * syn_vnoise(X, Y, 4, seed + 3) + 2 * syn_vnoise(X, Y, 3, seed + 4);
    const int base = 40 + (int)((acc >> 4) * 5 >> 3);
    int val = base;
    const uint32_t cx0 = (X >> 5) - 1, cy0 = (Y >> 5) - 1;
    for (uint32_t dy = 0; dy < 2; dy++)
        for (uint32_t dx = 0; dx < 2; dx++) {
            const uint32_t cx = cx0 + dx, cy = cy0 + dy;
            const uint32_t hsh = syn_hash(cx, cy, seed ^ 0xABCD1234u);
            if ((hsh & 3u) == 0) continue;
            const uint32_t x0 = (cx << 5) + ((hsh >> 2) & 31), y0 = (cy << 5) + ((hsh >> 7) & 31);
            const uint32_t rw = 5 + ((hsh >> 12) & 31) % 27, rh = 5 + ((hsh >> 17) & 31) % 27;
            if (X - x0 < rw && Y - y0 < rh) {
                val = (int)(hsh >> 24) + ((base - 128) >> 2);
                val = val < 0 ? 0 : val > 255 ? 255 : val;
            }
        }
    return (uint8_t)val;
}

__global__ void k_synth_gray(uint32_t seed, int first, int w, int hgt, uint8_t *out, size_t step, size_t fstride)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= w) return;
    const uint32_t OX = (1u << 20) + 4u * (uint32_t)(first + f), OY = (1u << 20) + 2u * (uint32_t)(first + f);
    out[(size_t)f * fstride + (size_t)y * step + x] = syn_gray_px(seed, OX + (uint32_t)x, OY + (uint32_t)y);
    (void)hgt;
}
__global__ void k_synth_depth(uint32_t seed, int first, int w, int hgt, uint16_t *out, size_t step, size_t fstride)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= w) return;
    const uint32_t X = (1u << 20) + 4u * (uint32_t)(first + f) + (uint32_t)x, Y = (1u << 20) + 2u * (uint32_t)(first + f) + (uint32_t)y;
    uint32_t d = 600 + 12 * syn_vnoise(X, Y, 6, seed ^ 0x0D0D0D0Du);
    if (syn_hash(X >> 4, Y >> 4, seed ^ 0xDDDD0001u) % 100u < 15u) d = 0;
    *(uint16_t *)((uint8_t *)out + (size_t)f * fstride + (size_t)y * step + (size_t)x * 2) = (uint16_t)d;
    (void)hgt;
}
__global__ void k_synth_desc(uint32_t seed, uint64_t first_row, int64_t nrows, uint32_t *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * 8) return;
    const uint64_t row = first_row + (uint64_t)(i >> 3);
    const uint32_t k = (uint32_t)(i & 7);
    out[i] = syn_hash((uint32_t)row, (uint32_t)(row >> 32) * 8u + k, seed);
}

void launch_synth_gray(orbx_handle *h, uint32_t seed, int first, int n, int w, int hh, uint8_t *d, size_t step, size_t fstride)
{
    dim3 grid((w + 255) / 256, hh, n);
    k_synth_gray<<<grid, 256, 0, h->stream>>>(seed, first, w, hh, d, step, fstride); h->launches++;
}
void launch_synth_depth(orbx_handle *h, uint32_t seed, int first, int n, int w, int hh, uint16_t *d, size_t step, size_t fstride)
{
    dim3 grid((w + 255) / 256, hh, n);
    k_synth_depth<<<grid, 256, 0, h->stream>>>(seed, first, w, hh, d, step, fstride); h->launches++;
}
void launch_synth_desc(orbx_handle *h, uint32_t seed, uint64_t first_row, int64_t nrows, uint8_t *d)
{
    const int64_t n = nrows * 8;
    k_synth_desc<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(seed, first_row, nrows, (uint32_t *)d); h->launches++;
}
