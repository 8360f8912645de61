// orbx_tma.h — TMA / mbarrier device primitives and the per-level tensor maps shared by the tiled kernels.
// Every pyramid level is described to the TMA unit as a 3-D tensor of u32 elements
//   (row pitch / 4) x rows x frames,  box = ORBX_TMA_BOX_WORDS x (hCell + 6) x 1,  zero fill outside,
// so one cp.async.bulk.tensor.3d moves a 288-byte x (hCell + 6)-row window of any frame into shared memory.
// Box start columns must be multiples of 16 bytes (4 elements) — an unaligned inner coordinate is an illegal
// instruction on sm_100a.
#pragma once
#include "orbx_internal.h"

#define ORBX_TMA_BOX_WORDS 72            // 288 bytes
#define ORBX_TMA_BOX_BYTES (ORBX_TMA_BOX_WORDS * 4)
#define ORBX_RZ_BOX_ROWS 80               // source rows staged per resize tile (k_pyramid.cu)
#define ORBX_RZ2_BOX_WORDS 48             // 192 bytes: source window of a 128-column output tile (k_resize_linear2)
#define ORBX_RZ2_BOX_BYTES (ORBX_RZ2_BOX_WORDS * 4)

struct LevelMaps { CUtensorMap m[ORBX_MAX_LEVELS]; };

// ---- TMA / mbarrier primitives (sm_90+ PTX; SASS: UTMALDG, SYNCS) ----
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
#endif
