// k_filter.cu — post-selection filters, fused into one stable compaction per frame.
//   * Frontend::isValidDepth / filterDepth (reference frontend.cpp:457-473, 503-527): keep a keypoint iff
//     x=round(pt.x), y=round(pt.y) (half away from zero) is inside the depth image and
//     d = depth_u16 * 0.001f satisfies depth_min <= d <= depth_max;
//   * Backend::categorizeObservation + filtered_objects_ (backend.cpp:1011-1029, 746-751): the first
//     box containing the pixel (inclusive bounds, fp64 compares) gives the class; drop the keypoint
//     if that class is in the drop mask.
// Both run AFTER the quadtree, exactly where the reference applies them, so the
// retained set is identical to extract-then-filter.  One CTA per frame; order preserved.
#include "orbx_internal.h"

struct FilterParams {
    const orbx_keypoint *kin; const uint8_t *din; const int32_t *nin; int cap_in;
    const uint16_t *depth; size_t dstep, dfstride; int dw, dh;     // steps in BYTES
    float dmin, dmax;
    const orbx_box *boxes; const int32_t *box_offsets; int box_base; int nboxes; unsigned long long drop_mask;   // box_offsets (nullable): frame f owns boxes [off[f], off[f+1])
    orbx_keypoint *kout; uint8_t *dout; int32_t *nout; int cap_out;
    int32_t *status;
};

__global__ void __launch_bounds__(1024) k_filter(FilterParams P)
{
    __shared__ int s_warp[33];                              // one CTA per frame, any block size that is a multiple of 32
    __shared__ int s_base;
    ORBX_PDL_ENTRY();
    const int f = blockIdx.x;
    const int n = P.nin[f];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const orbx_keypoint *kin = P.kin + (size_t)f * P.cap_in;
    const uint8_t *din = P.din + (size_t)f * P.cap_in * ORBX_DESC_BYTES;
    orbx_keypoint *kout = P.kout + (size_t)f * P.cap_out;
    uint8_t *dout = P.dout + (size_t)f * P.cap_out * ORBX_DESC_BYTES;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        bool keep = false;
        orbx_keypoint kp;
        if (i < n) {
            kp = kin[i];
            keep = true;
            if (P.depth) {
                const int x = (int)roundf(kp.x), y = (int)roundf(kp.y);
                if (x < 0 || y < 0 || x >= P.dw || y >= P.dh) keep = false;
                else {
                    const uint16_t raw = *(const uint16_t *)((const uint8_t *)P.depth + (size_t)f * P.dfstride + (size_t)y * P.dstep + (size_t)x * 2);
                    const float d = __fmul_rn((float)raw, 0.001f);
                    if (d < P.dmin || d > P.dmax) keep = false;
                }
            }
            if (keep && P.nboxes > 0) {
                const double px = (double)kp.x, py = (double)kp.y;
                const int b0 = P.box_offsets ? P.box_offsets[f] - P.box_base : 0, b1 = P.box_offsets ? P.box_offsets[f + 1] - P.box_base : P.nboxes;
                for (int b = b0; b < b1; b++) {
                    const orbx_box bx = P.boxes[b];
                    if (px >= bx.cx - bx.w / 2 && px <= bx.cx + bx.w / 2 && py >= bx.cy - bx.h / 2 && py <= bx.cy + bx.h / 2) {
                        if (bx.class_id >= 0 && bx.class_id < 64 && ((P.drop_mask >> bx.class_id) & 1ull)) keep = false;
                        break;                                  // first containing box decides
                    }
                }
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) { int run = 0; for (int w = 0; w < nw; w++) { const int c = s_warp[w]; s_warp[w] = run; run += c; } s_warp[32] = run; }
        __syncthreads();
        const int o = s_base + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
        if (keep) {
            if (o < P.cap_out) {
                kout[o] = kp;
                const uint4 *s = reinterpret_cast<const uint4 *>(din + (size_t)i * ORBX_DESC_BYTES);
                uint4 *d = reinterpret_cast<uint4 *>(dout + (size_t)o * ORBX_DESC_BYTES);
                d[0] = s[0]; d[1] = s[1];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_warp[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (s_base > P.cap_out) { atomicOr(P.status, ORBX_DS_KP_OVERFLOW); P.nout[f] = 0; }
        else P.nout[f] = s_base;
    }
}

void launch_filter(orbx_handle *h, int nframes, const uint16_t *d_depth, size_t dstep, size_t dfstride,
                   const orbx_box *d_boxes, const int32_t *d_box_offsets, int box_base, int nboxes, uint64_t drop_mask,
                   orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_counts)
{
    FilterParams P;
    P.kin = h->d_kps_all; P.din = h->d_desc_all; P.nin = h->d_count_all; P.cap_in = h->max_kp;
    P.depth = d_depth; P.dstep = dstep; P.dfstride = dfstride; P.dw = h->geo.width; P.dh = h->geo.height;
    P.dmin = h->prm.depth_min; P.dmax = h->prm.depth_max;
    P.boxes = d_boxes; P.box_offsets = d_box_offsets; P.box_base = box_base; P.nboxes = nboxes; P.drop_mask = drop_mask;
    P.kout = d_kps; P.dout = d_desc; P.nout = d_counts; P.cap_out = cap; P.status = h->d_status;
    ProfScope ps(h, ORBX_K_FILTER);
    orbx_launch_pdl(h, k_filter, dim3(nframes), dim3(1024), 0, h->stream, P);
}

// ---- keyframe packing: the landmark / observation loop of Frontend::publishKeyframe (reference frontend.cpp:731-776) ----
// per keypoint: depth at round(pt) (half away from zero) * 0.001f, back-projection in FLOAT with the float intrinsics
// ((pt.x - cx) * d / fx, ...), kept iff z > 0.3 && z < 3.0 (float against double literals), world = R * p + t in double with
// cv::Mat's accumulation order and no FMA; one 80-byte record (Landmark + Observation of Keyframe.msg) per kept keypoint,
// order preserved (stable ballot compaction, one CTA per frame).
struct PackParams {
    const orbx_keypoint *kps; const uint8_t *desc; const int32_t *counts; int cap_in;
    const uint16_t *depth; size_t dstep, dfstride; int dw, dh;     // steps in BYTES
    orbx_kfparams K;
    orbx_kfrecord *out; int32_t *nout; int cap_out;
    int32_t *status;
};

__global__ void __launch_bounds__(1024) k_pack_keyframe(PackParams P)
{
    __shared__ int s_warp[33];
    __shared__ int s_base;
    const int f = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = min(P.counts[f], P.cap_in);
    const orbx_keypoint *kin = P.kps + (size_t)f * P.cap_in;
    const uint8_t *din = P.desc + (size_t)f * P.cap_in * ORBX_DESC_BYTES;
    orbx_kfrecord *out = P.out + (size_t)f * P.cap_out;
    const uint8_t *dimg = reinterpret_cast<const uint8_t *>(P.depth) + (size_t)f * P.dfstride;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        bool keep = false;
        float px = 0.f, py = 0.f, X = 0.f, Y = 0.f, Z = 0.f;
        if (i < n) {
            px = kin[i].x; py = kin[i].y;
            const int x = (int)roundf(px), y = (int)roundf(py);
            if (x >= 0 && y >= 0 && x < P.dw && y < P.dh) {
                Z = __fmul_rn((float)__ldg(reinterpret_cast<const uint16_t *>(dimg + (size_t)y * P.dstep) + x), 0.001f);
                X = __fdiv_rn(__fmul_rn(__fsub_rn(px, P.K.cx), Z), P.K.fx);
                Y = __fdiv_rn(__fmul_rn(__fsub_rn(py, P.K.cy), Z), P.K.fy);
                keep = (double)Z > 0.3 && (double)Z < 3.0;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) { int run = 0; for (int w = 0; w < nw; w++) { const int c = s_warp[w]; s_warp[w] = run; run += c; } s_warp[32] = run; }
        __syncthreads();
        if (keep) {
            const int o = s_base + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
            if (o < P.cap_out) {
                orbx_kfrecord r;
                r.landmark_id = (uint64_t)i;
#pragma unroll
                for (int k = 0; k < 3; k++)
                    r.position[k] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.K.R[3 * k], (double)X), __dmul_rn(P.K.R[3 * k + 1], (double)Y)),
                                                        __dmul_rn(P.K.R[3 * k + 2], (double)Z)), P.K.t[k]);
                r.pixel_x = (double)px; r.pixel_y = (double)py;
                const uint4 *dp = reinterpret_cast<const uint4 *>(din + (size_t)i * ORBX_DESC_BYTES);
                *reinterpret_cast<uint4 *>(r.descriptor) = dp[0];
                *reinterpret_cast<uint4 *>(r.descriptor + 16) = dp[1];
                out[o] = r;
            } else atomicOr(P.status, ORBX_DS_KP_OVERFLOW);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_warp[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) P.nout[f] = min(s_base, P.cap_out);
}

void launch_pack_keyframe(orbx_handle *h, int nframes, const orbx_keypoint *d_kps, const uint8_t *d_desc, const int32_t *d_counts, int cap_in,
                          const uint16_t *d_depth, size_t dstep, size_t dfstride, int dw, int dh, const orbx_kfparams *K,
                          orbx_kfrecord *d_out, int32_t *d_nout, int cap_out)
{
    PackParams P;
    P.kps = d_kps; P.desc = d_desc; P.counts = d_counts; P.cap_in = cap_in;
    P.depth = d_depth; P.dstep = dstep; P.dfstride = dfstride; P.dw = dw; P.dh = dh;
    P.K = *K; P.out = d_out; P.nout = d_nout; P.cap_out = cap_out; P.status = h->d_status;
    ProfScope ps(h, ORBX_K_OTHER);
    k_pack_keyframe<<<nframes, 1024, 0, h->stream>>>(P);
}
