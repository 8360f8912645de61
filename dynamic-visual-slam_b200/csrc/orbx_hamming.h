// orbx_hamming.h — the 256-bit Hamming distance shared by the matcher and the landmark association kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// POPC is the scarce instruction (measured 15.8 lanes/clk/SM = 8.1 cycles per warp instruction on a scheduler; LOP3 takes 2), so the
// eight XOR words go through carry-save adders first: three full adders turn seven of the words into one sum word (weight 1) and
// three carry words (weight 2) — five POPC instead of eight for six extra LOP3.  Measured per 128-frame step (k_match_partial):
// 8 POPC 0.185 ms, 6 POPC 0.150, 5 POPC 0.136, the full Harley-Seal tree with 4 POPC (14 extra LOP3) 0.167 — past five the ALU pipe
// becomes the bound.
#ifndef ORBX_MATCH_CSA
#define ORBX_MATCH_CSA 2
#endif
__device__ __forceinline__ void csa3(uint32_t &carry, uint32_t &sum, uint32_t a, uint32_t b, uint32_t c)
{
    sum = a ^ b ^ c;
    carry = (a & b) | (c & (a ^ b));
}
__device__ __forceinline__ int hamming256(const uint32_t q[8], const uint4 a, const uint4 b)
{
#if ORBX_MATCH_CSA == 3
    uint32_t ones, twos, t2a, t2b, t2c, f4a;
    csa3(t2a, ones, q[0] ^ a.x, q[1] ^ a.y, q[2] ^ a.z);
    csa3(t2b, ones, ones, q[3] ^ a.w, q[4] ^ b.x);
    csa3(t2c, ones, ones, q[5] ^ b.y, q[6] ^ b.z);
    const uint32_t x7 = q[7] ^ b.w;
    const uint32_t t2d = ones & x7;
    ones ^= x7;
    csa3(f4a, twos, t2a, t2b, t2c);
    const uint32_t f4b = twos & t2d;
    twos ^= t2d;
    const uint32_t fours = f4a ^ f4b, eights = f4a & f4b;
    return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
#elif ORBX_MATCH_CSA == 2
    uint32_t s0, c0, s1, c1, s2, c2;
    csa3(c0, s0, q[0] ^ a.x, q[1] ^ a.y, q[2] ^ a.z);
    csa3(c1, s1, q[3] ^ a.w, q[4] ^ b.x, q[5] ^ b.y);
    csa3(c2, s2, s0, s1, q[6] ^ b.z);
    return __popc(s2) + __popc(q[7] ^ b.w) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
#elif ORBX_MATCH_CSA == 1
    uint32_t s0, c0, s1, c1;
    csa3(c0, s0, q[0] ^ a.x, q[1] ^ a.y, q[2] ^ a.z);
    csa3(c1, s1, q[3] ^ a.w, q[4] ^ b.x, q[5] ^ b.y);
    return __popc(s0) + __popc(s1) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w) + 2 * (__popc(c0) + __popc(c1));
#else
    return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
           __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
#endif
}

