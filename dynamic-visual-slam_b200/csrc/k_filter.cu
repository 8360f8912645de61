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
#include "orbx_keep.h"

struct FilterParams {
    const orbx_keypoint *kin; const uint8_t *din; const int32_t *nin; int cap_in;
    KeepParams K;                                                 // depth + box test (orbx_keep.h)
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
            keep = orbx_keep(P.K, f, kp.x, kp.y);
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

static KeepParams keep_params(const orbx_handle *h, const uint16_t *d_depth, size_t dstep, size_t dfstride,
                              const orbx_box *d_boxes, const int32_t *d_box_offsets, int box_base, int nboxes, uint64_t drop_mask)
{
    KeepParams K;
    K.depth = d_depth; K.dstep = dstep; K.dfstride = dfstride; K.dw = h->geo.width; K.dh = h->geo.height;
    K.dmin = h->prm.depth_min; K.dmax = h->prm.depth_max;
    K.boxes = d_boxes; K.box_offsets = d_box_offsets; K.box_base = box_base; K.nboxes = nboxes; K.drop_mask = drop_mask;
    return K;
}

// ---- filter-first order: the same test on the SELECTED positions, before any descriptor work ----
// The filters look at a keypoint's position only: pt = (selected pixel + border) * the level's scale (ORBextractor.cpp:886-887, 1147-1149) is known
// as soon as the quadtree has selected it.  One CTA per frame walks the frame's selected list in output order (levels concatenated, :1123),
// forms pt exactly as the descriptor kernel does, applies orbx_keep and leaves the surviving list indices, in order, in map[f][0 .. nout[f]) —
// k_describe_fused then works on those alone and writes their final rows.  Same capacity semantics as k_filter.
struct KeepListParams {
    const uint32_t *sel; int sel_slab; const int32_t *nsel;
    KeepParams K;
    int32_t *map; int map_slab; int32_t *nout; int cap_out;
    int32_t *status;
};

__global__ void __launch_bounds__(1024) k_keep_list(KeepListParams P, const FrameGeom *__restrict__ G)
{
    __shared__ int s_warp[33];
    __shared__ int s_base;
    __shared__ int s_lend[ORBX_MAX_LEVELS + 1];                 // s_lend[l] = selected keypoints of levels < l
    ORBX_PDL_ENTRY();
    const int f = blockIdx.x;
    const int nl = G->nlevels;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        int run = 0;
        for (int l = 0; l < nl; l++) { s_lend[l] = run; run += P.nsel[f * nl + l]; }
        s_lend[nl] = run; s_base = 0;
    }
    __syncthreads();
    const int n = s_lend[nl];
    const uint32_t *sel = P.sel + (size_t)f * P.sel_slab;
    int32_t *map = P.map + (size_t)f * P.map_slab;
    const int nw = blockDim.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        bool keep = false;
        if (i < n) {
            int level = 0;
            while (level + 1 < nl && i >= s_lend[level + 1]) level++;
            const LevelGeom &g = G->lv[level];
            const uint32_t c = sel[g.sel_off + (i - s_lend[level])];
            float kx = (float)(orbx_px(c) + ORBX_BORDER), ky = (float)(orbx_py(c) + ORBX_BORDER);
            if (level != 0) { kx = __fmul_rn(kx, g.scale); ky = __fmul_rn(ky, g.scale); }
            keep = orbx_keep(P.K, f, kx, ky);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        if (threadIdx.x == 0) { int run = 0; for (int w = 0; w < nw; w++) { const int c = s_warp[w]; s_warp[w] = run; run += c; } s_warp[32] = run; }
        __syncthreads();
        const int o = s_base + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
        if (keep && o < P.cap_out && o < P.map_slab) map[o] = i;
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_warp[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // (n > map_slab: more selected keypoints than the handle's max_keypoints — the reference's order reports that from the descriptor kernel)
        if (s_base > P.cap_out || n > P.map_slab) { atomicOr(P.status, ORBX_DS_KP_OVERFLOW); P.nout[f] = 0; }
        else P.nout[f] = s_base;
    }
}

void launch_keep_list(orbx_handle *h, int nframes, const uint16_t *d_depth, size_t dstep, size_t dfstride,
                      const orbx_box *d_boxes, const int32_t *d_box_offsets, int box_base, int nboxes, uint64_t drop_mask,
                      int32_t *d_map, int map_slab, int32_t *d_counts, int cap)
{
    KeepListParams P;
    P.sel = h->d_sel; P.sel_slab = h->geo.sel_entries; P.nsel = h->d_nsel;
    P.K = keep_params(h, d_depth, dstep, dfstride, d_boxes, d_box_offsets, box_base, nboxes, drop_mask);
    P.map = d_map; P.map_slab = map_slab; P.nout = d_counts; P.cap_out = cap; P.status = h->d_status;
    ProfScope ps(h, ORBX_K_FILTER);
    orbx_launch_pdl(h, k_keep_list, dim3(nframes), dim3(1024), 0, h->stream, P, (const FrameGeom *)h->d_geo);
}

void launch_filter(orbx_handle *h, int nframes, const uint16_t *d_depth, size_t dstep, size_t dfstride,
                   const orbx_box *d_boxes, const int32_t *d_box_offsets, int box_base, int nboxes, uint64_t drop_mask,
                   orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_counts)
{
    FilterParams P;
    P.kin = h->d_kps_all; P.din = h->d_desc_all; P.nin = h->d_count_all; P.cap_in = h->max_kp;
    P.K = keep_params(h, d_depth, dstep, dfstride, d_boxes, d_box_offsets, box_base, nboxes, drop_mask);
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
