// orbx_match.h — parameters and key packing shared by the matcher kernels (k_match.cu: POPC and mma.sync; k_match_umma.cu: tcgen05)
#pragma once
#include "orbx_internal.h"

#define MT_THREADS 128
#define MT_TILE 256                 // train rows staged per shared-memory tile (8 KB)
#define MT_KEY_SHIFT 22
#define MT_INF 0xFFFFFFFFu

struct MatchParams {
    const uint8_t *q; const int32_t *nq_arr; int nq_imm; size_t q_stride;     // stride between problems' query sets (bytes)
    const uint8_t *t; const int32_t *nt_arr; int nt_imm; size_t t_stride;
    const int32_t *qsel, *tsel;     // problem -> set index (nullable: identity)
    unsigned long long *part;        // [problem][split][nq_max][2] 64-bit keys (dist<<32 | global row)
    int nq_max, nsplit, rows_per_split;
    uint32_t row_base;               // global index of train row 0 (database shards)
};

// k_match_umma.cu: the tensor-memory matcher (tcgen05.mma kind::i8); same partial layout as k_match_partial / k_match_mma
void launch_match_umma(orbx_handle *h, const MatchParams &P, dim3 grid, bool top2);
