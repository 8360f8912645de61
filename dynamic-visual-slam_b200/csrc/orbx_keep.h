// orbx_keep.h — the post-selection keep test of a keypoint, shared by k_filter (stable compaction of described rows, the reference's order) and
// k_keep_list (the same test on the selected positions, before any descriptor work).  ONE function, so both orders take the same decision:
//   * Frontend::isValidDepth / filterDepth (reference frontend.cpp:457-473, 503-527): keep iff x = round(pt.x), y = round(pt.y) (half away
//     from zero) lies inside the depth image and d = depth_u16 * 0.001f satisfies depth_min <= d <= depth_max;
//   * Backend::categorizeObservation + filtered_objects_ (backend.cpp:1011-1029, 746-751): the first box containing the pixel (inclusive
//     bounds, fp64 compares) gives the class; drop iff that class is in the drop mask.
#pragma once
#include "orbx_internal.h"

struct KeepParams {
    const uint16_t *depth; size_t dstep, dfstride; int dw, dh;     // steps in BYTES; depth == nullptr: no depth test
    float dmin, dmax;
    const orbx_box *boxes; const int32_t *box_offsets; int box_base; int nboxes; unsigned long long drop_mask;   // box_offsets (nullable): frame f owns boxes [off[f], off[f+1])
};

__device__ __forceinline__ bool orbx_keep(const KeepParams &P, int f, float kx, float ky)
{
    if (P.depth) {
        const int x = (int)roundf(kx), y = (int)roundf(ky);
        if (x < 0 || y < 0 || x >= P.dw || y >= P.dh) return false;
        const uint16_t raw = *(const uint16_t *)((const uint8_t *)P.depth + (size_t)f * P.dfstride + (size_t)y * P.dstep + (size_t)x * 2);
        const float d = __fmul_rn((float)raw, 0.001f);
        if (d < P.dmin || d > P.dmax) return false;
    }
    if (P.nboxes > 0) {
        const double px = (double)kx, py = (double)ky;
        const int b0 = P.box_offsets ? P.box_offsets[f] - P.box_base : 0, b1 = P.box_offsets ? P.box_offsets[f + 1] - P.box_base : P.nboxes;
        for (int b = b0; b < b1; b++) {
            const orbx_box bx = P.boxes[b];
            if (px >= bx.cx - bx.w / 2 && px <= bx.cx + bx.w / 2 && py >= bx.cy - bx.h / 2 && py <= bx.cy + bx.h / 2)   // first containing box decides
                return !(bx.class_id >= 0 && bx.class_id < 64 && ((P.drop_mask >> bx.class_id) & 1ull));
        }
    }
    return true;
}
