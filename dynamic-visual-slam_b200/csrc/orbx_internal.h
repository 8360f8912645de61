// orbx_internal.h — shared declarations of the sm_100a ORB path (not part of the ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/orbx.h"

#define ORBX_EDGE 19                 // EDGE_THRESHOLD, reference ORBextractor.cpp:73
#define ORBX_BORDER 16               // EDGE_THRESHOLD - 3  (minBorderX, ORBextractor.cpp:787)
#define ORBX_CELL_W 35               // W, ORBextractor.cpp:783
#define ORBX_HALF_PATCH 15
#define ORBX_PATCH 31
#define ORBX_BLUR_TW 256             // blur tile: 256 px x hCell rows per CTA (k_blur.cu)
#define ORBX_FAST_MAX_W 240          // max detection width (px) of one FAST strip: 16-byte aligned TMA box of 288 bytes (k_fast.cu)

// candidate / selected-keypoint packing: x:12 | y:12 | score:8, coordinates relative to the border box
__host__ __device__ inline uint32_t orbx_pack(int x, int y, int s) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24); }
__host__ __device__ inline int orbx_px(uint32_t c) { return (int)(c & 0xFFFu); }
__host__ __device__ inline int orbx_py(uint32_t c) { return (int)((c >> 12) & 0xFFFu); }
__host__ __device__ inline int orbx_ps(uint32_t c) { return (int)(c >> 24); }

// Geometry of one pyramid level for the current (width, height); lives in __constant__-like param structs.
struct LevelGeom {
    int w, h;                // level image size (ORBextractor.cpp:1173-1174)
    int pitch;               // bytes per row in the pyramid slab (multiple of 128); level 0 aliases the input frame
    size_t off;              // byte offset of the level inside a frame's pyramid slab (levels >= 1)
    int bpitch;              // bytes per row in the blurred slab (all levels)
    size_t boff;             // byte offset inside a frame's blurred slab
    int ncols, nrows;        // cell grid (ORBextractor.cpp:799-802)
    int wcell, hcell;
    int cell_first;          // index of this level's first cell in the flattened all-level cell list
    int cells_per_strip, strips_per_row, strip_first;   // FAST strips: runs of cells of one cell row
    int blur_first, blur_tx, blur_ty;   // blur tiles
    int N;                   // mnFeaturesPerLevel
    int nini;                // quadtree roots (ORBextractor.cpp:559)
    float hx;                // root width
    int cand_cap;            // capacity of the candidate list
    size_t cand_off;         // offset (in entries) of this level's candidate list inside a frame's candidate slab
    int sel_cap;             // capacity of the selected list = node cap
    int sel_off;             // offset (entries) inside a frame's selected slab
    float scale;             // mvScaleFactor[level]
    float size;              // (float)(int)(31*scale)
    // resize tables (level l from l-1): offsets into the table buffers
    int xtab_off, ytab_off;
};

struct FrameGeom {
    int nlevels;
    int width, height;
    int total_cells, total_blur_tiles, total_strips, max_hcell, max_wcell, total_cells_valid;
    int rz_tw, rz_th;        // output tile of the resize kernel: its source window fits the 288-byte x 80-row TMA box at every level
    int rz2_ok;              // the 128 x 64 tile of the batch kernel fits a 192-byte x 80-row box at every level (scale factors up to ~1.35)
    size_t pyr_bytes;        // per-frame pyramid slab size (levels 1..)
    size_t blur_bytes;       // per-frame blurred slab size (levels 0..)
    size_t cand_entries;     // per-frame candidate slab entries
    int sel_entries;         // per-frame selected slab entries
    int node_cap_max;
    LevelGeom lv[ORBX_MAX_LEVELS];
};

// dense FAST formulation (k_fast_dense.cu): score maps, corner lists and tiles of a frame geometry
struct DenseLevel {
    int map_off, map_pitch;      // score map of the level inside a frame's map slab (bytes); map column 0 = level column 4, map row 0 = level row 19
    int cl_off, cl_cap;          // corner list (entries) inside a frame's corner slab
    int ntx, nty;                // tiles of 128 x 16 detection pixels
    int ncv, nrv, cellv_first;   // cells the reference runs: columns, rows, index of the level's first cell record
};
struct DenseGeom { DenseLevel lv[ORBX_MAX_LEVELS]; int ntiles; size_t map_bytes, cl_entries; };
#define ORBX_FAST_DENSE_MIN_FRAMES 8

enum { ORBX_K_RESIZE = 0, ORBX_K_FAST, ORBX_K_QUADTREE, ORBX_K_BLUR, ORBX_K_DESCRIBE, ORBX_K_FILTER,
       ORBX_K_MATCH, ORBX_K_MATCH_EPI, ORBX_K_OTHER, ORBX_K_FAST_DENSE, ORBX_K_FAST_NMS, ORBX_K_FAST_RETRY, ORBX_K_COUNT };

struct ResizeTab { int32_t ofs; int16_t a0, a1; };   // 8 bytes

// ORBX_OPT_OVERLAP: a second set of everything one pipeline run writes besides the caller's buffers.  A large batch is cut in two halves
// that run on two streams, the second one stage behind the first, so that kernels bound on different units (FAST: ALU issue, descriptor
// kernel: shared-memory wavefronts, matcher: POPC / ALU, quadtree: latency) share the SMs instead of running back to back.  The kernels
// and their launchers are untouched: the handle's active fields are swapped with this set around the second half.
struct OrbxLane {
    bool allocated;
    cudaStream_t stream;
    uint8_t *d_pyr; uint32_t *d_cand, *d_cand2, *d_qtmp; uint16_t *d_owner, *d_owner2; int32_t *d_ncand, *d_nsel; uint32_t *d_sel;
    orbx_keypoint *d_kps_all; uint8_t *d_desc_all; int32_t *d_count_all; uint32_t *d_mpart; size_t mpart_cap;
    CUtensorMap tmap[ORBX_MAX_LEVELS], tmap_rz[ORBX_MAX_LEVELS], tmap_rz2[ORBX_MAX_LEVELS], tmap_cell[ORBX_MAX_LEVELS]; bool tmap_valid;
    const uint8_t *tmap_l0; size_t tmap_l0_step, tmap_l0_fstride; int tmap_l0_frames;
};

struct orbx_cvorb;                     // profile C (cv::ORB) state, orbx_cvorb.cu
struct orbx_handle {
    orbx_params prm;
    orbx_cvorb *cv;              // non-null iff prm.profile == ORBX_PROFILE_CVORB
    int device;
    cudaStream_t stream, copy_stream, out_stream, aux_stream;   // aux: blur runs beside FAST + quadtree
    cudaEvent_t ev_fork, ev_fork0, ev_join;
    int opt_serial;
    int opt_pdl;                 // 1 (default): programmatic stream serialization where it measured faster (below)
    bool pdl_chain;              // this call's kernels use it: small batches only — at 128 frames the early-resident CTAs of the next kernel cost
                                 // 1 % of throughput, at one frame they take 9 % off the latency; the resize chain always uses it (+0.5 %)
    int opt_filter_first;        // 1 (default): depth / box filter on the selected positions BEFORE the descriptor kernel (k_keep_list), which then works on the survivors only
    int opt_fused_blur;          // 1 (default): the Gaussian is evaluated inside the descriptor kernel, no blurred pyramid is written
    bool blur_valid;             // d_blur holds the blurred levels of the last batch
    int opt_fast_ctas;           // FAST warps per SM in the overlapped schedule (0 = as many as fit)
    int opt_match_mma;           // 1 (default): Hamming matching as an int8 tensor-core GEMM (k_match_mma); 0: the POPC kernel (k_match_partial)
    int opt_overlap;             // 1 (default): batches of >= 32 frames run as two half-batches on two streams (OrbxLane)
    OrbxLane alt;                // the second lane's scratch (allocated on first use)
    bool in_overlap;             // a half-batch run is being enqueued (FAST caps its resident warps so the other lane's kernels fit beside it)
    cudaEvent_t ev_lane_fork, ev_lane_join, ev_prev_ready;
    cudaEvent_t ev_after_pyramid;            // when set, run_pipeline records it right after the pyramid launches
    cudaEvent_t ev_a, ev_b;
    cudaEvent_t ev_in[2], ev_comp[2], ev_out[2];   // host-batch pipeline: input slot filled / kernels done / outputs copied
    int chunk;                                      // frames per pipeline chunk of the synchronous host-buffer batch calls
    uint64_t seq;                                   // chunks enqueued so far; chunk n uses staging slot n & 1
    struct { bool active; uint64_t ticket; const int32_t *counts; int nframes, cap; } pending[2];   // asynchronous submissions
    std::string err;
    int64_t launches;
    // extractor tables (ORBextractor.cpp:409-469)
    float scale[ORBX_MAX_LEVELS], inv_scale[ORBX_MAX_LEVELS], sigma2[ORBX_MAX_LEVELS], inv_sigma2[ORBX_MAX_LEVELS];
    int nfeat[ORBX_MAX_LEVELS];
    int umax[16];
    int max_kp;
    // geometry of the current frame size (rebuilt when width/height change)
    FrameGeom geo;
    FrameGeom *d_geo;
    ResizeTab *d_xtab, *d_ytab; int tab_cap;
    uint32_t *d_strips; int strip_cap;   // FAST strip descriptors (k_fast.cu)
    uint32_t *d_blur_tiles; int blur_tile_cap;   // blur tile descriptors (k_blur.cu)
    // TMA tensor maps of the pyramid levels (k_fast.cu): levels >= 1 depend on the geometry only, level 0 on the caller's frames
    CUtensorMap tmap[ORBX_MAX_LEVELS]; bool tmap_valid;      // box rows = hCell + 6 (FAST strips, blur tiles)
    CUtensorMap tmap_rz[ORBX_MAX_LEVELS];                    // box rows = ORBX_RZ_BOX_ROWS (resize source windows)
    CUtensorMap tmap_rz2[ORBX_MAX_LEVELS];                   // the same with a 192-byte box: source windows of the 128-column batch tiles
    CUtensorMap tmap_cell[ORBX_MAX_LEVELS];                  // box = 96 bytes x (hCell + 6) rows (FAST cell windows)
    void *d_cells; int cell_cap;                             // FAST cell records, 32 bytes each (k_fast.cu: orbx_build_fast_cells)
    const uint8_t *tmap_l0; size_t tmap_l0_step, tmap_l0_fstride; int tmap_l0_frames;
    int pyr_grid_cap;                                     // resident CTAs of the cooperative pyramid kernel (0 = not probed, -1 = unavailable)
    size_t smem_optin_max;                                // cudaDevAttrMaxSharedMemoryPerBlockOptin of the handle's device
    int fast_grid_cap; size_t fast_smem; int fast_tp;   // resident CTAs / dynamic smem / tile pitch of the persistent FAST kernel
    // dense FAST formulation (k_fast_dense.cu), used for batches: ORBX_OPT_FAST_DENSE 0 = never, 1 (default) = batches of >= ORBX_FAST_DENSE_MIN_FRAMES, 2 = always
    int opt_fast_dense; bool dense_ok; int dense_grid_cap; unsigned geo_serial;
    int opt_dense_fuse;          // 0 (default): the NMS runs as a second kernel; 1: as work items inside k_fast_dense while the frame's score map is in L2 (measured slower)
    DenseGeom dgeo;
    void *d_dtiles; int dtile_cap;                        // tile records, 16 bytes each
    uint8_t *d_smap; size_t smap_cap;                     // score maps [batch]
    uint32_t *d_clist; size_t clist_cap;                  // corner lists [batch] (entries)
    int32_t *d_dense_zero; size_t dense_zero_bytes;       // [work counter, retry count, 2 spare][corner counts][served-cell flags]: zeroed per launch
    int32_t *d_retry;                                     // (frame, cell) items of the minThFAST retry launch
    CUtensorMap tmap_dense[ORBX_MAX_LEVELS];              // box = 144 bytes x 22 rows
    const uint8_t *tmap_dense_pyr; unsigned tmap_dense_serial;
    const uint8_t *tmap_dense_l0; size_t tmap_dense_l0_step, tmap_dense_l0_fstride; int tmap_dense_l0_frames;
    // arenas, sized for max_width x max_height x max_batch
    uint8_t *d_pyr, *d_blur;     size_t pyr_slab, blur_slab;          // current per-frame strides
    size_t pyr_cap, blur_cap;                                          // arena bytes
    uint8_t *d_in; size_t in_cap;          // staging for host-API inputs (gray)
    uint8_t *d_bgr; size_t bgr_cap;        // staging for the BGR single-frame call (grown on first use)
    uint16_t *d_depth_in; size_t depth_cap;
    uint32_t *d_cand, *d_cand2;  size_t cand_cap;        // candidate values (entries), ping-pong for the quadtree
    uint32_t *d_qtmp; uint16_t *d_owner, *d_owner2;
    int32_t *d_ncand;            // [batch][levels]
    uint32_t *d_sel; size_t sel_cap;        // selected per (frame, level), entries
    int32_t *d_nsel;             // [batch][levels]
    orbx_keypoint *d_kps_all; uint8_t *d_desc_all;   // unfiltered per-frame outputs [batch][max_kp]
    int32_t *d_count_all;        // [batch]
    orbx_keypoint *d_kps_out; uint8_t *d_desc_out; int32_t *d_count_out;   // host-API outputs [batch][max_kp]
    orbx_box *d_boxes; int boxes_cap;        // staged box lists: two slots of boxes_cap boxes (host-batch pipeline), slot 0 for single-frame calls
    int32_t *d_box_off;                      // staged per-frame offsets: two slots of max_batch + 1
    int32_t *d_status;           // device-side error flags: the word the kernels launched now report into
    int32_t *d_status_base;      // [0] synchronous and device calls, [1 + slot] the chunk in staging slot `slot` of the host-batch pipeline
    // pinned staging for the host API
    uint8_t *h_out; size_t h_out_bytes;
    int32_t *h_status;
    // matcher scratch
    uint32_t *d_mpart; size_t mpart_cap;    // partial top-2 keys
    uint8_t *d_mq, *d_mt; size_t mq_cap, mt_cap;
    orbx_dmatch *d_mout; size_t mout_cap; int32_t *d_mcount;
    int last_batch;
    const uint8_t *last_l0; size_t last_l0_step, last_l0_fstride;   // level 0 of the last batch (may alias caller memory)
    int sm_count;
    // previous-frame descriptors for the stream API (the frontend's prev_descriptors_, frontend.cpp:1258-1259)
    uint8_t *d_prev_desc; int32_t *d_prev_count; int prev_valid;
    // per-kernel CUDA-event profiling (bench roofline): pairs of events around every launch
    int prof_on; int prof_n;
    std::vector<cudaEvent_t> prof_ev; std::vector<int> prof_id;
    double prof_ms[ORBX_K_COUNT]; long prof_cnt[ORBX_K_COUNT];
};

struct ProfScope {
    orbx_handle *h; int idx; cudaStream_t st;
    ProfScope(orbx_handle *hh, int id, cudaStream_t s = nullptr) : h(hh), idx(-1), st(s ? s : hh->stream) {
        if (h->prof_on && (size_t)(2 * h->prof_n + 1) < h->prof_ev.size()) {
            idx = h->prof_n++; h->prof_id[idx] = id; cudaEventRecord(h->prof_ev[2 * idx], st);
        }
    }
    ~ProfScope() { if (idx >= 0) cudaEventRecord(h->prof_ev[2 * idx + 1], st); h->launches++; }
};

struct orbx_db {
    orbx_handle *h;
    uint8_t *d_rows; int64_t cap, rows; uint32_t first_index;
    float *d_pos;                      // landmark positions (xyz per row), allocated by orbx_db_set_positions
    float *d_qpx; size_t qpx_cap;      // staged query pixels of the host association call
    uint32_t *d_part; size_t part_cap;
    uint8_t *d_q; size_t q_cap; orbx_top2 *d_out; size_t out_cap;
};

#define ORBX_CUDA(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_); return ORBX_E_CUDA; } } while (0)

// ---- programmatic dependent launch ----
// Kernels of the per-step chain begin with ORBX_PDL_ENTRY(): wait until the previous launch of the stream is complete and visible, then
// release the next one.  Launched through orbx_launch_pdl the next kernel's CTAs become resident while the previous kernel drains, so a
// kernel boundary costs a barrier instead of a drain + launch + ramp (this is what the single-frame latency pays for: seven boundaries
// of kernels that run 5-35 us each).  Without the launch attribute (profiling pass, other call sites) both instructions are no-ops.
#define ORBX_PDL_ENTRY() do { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); } while (0)
template <typename... KArgs, typename... Args>
static inline void orbx_launch_pdl(orbx_handle *h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (h->prof_on || !h->opt_pdl || !h->pdl_chain) ? 0 : 1;
    if (cudaLaunchKernelEx(&cfg, kern, args...) != cudaSuccess) {           // e.g. a driver without the attribute: plain launch
        cudaGetLastError();
        cfg.numAttrs = 0;
        cudaLaunchKernelEx(&cfg, kern, args...);
    }
}

// Dynamic shared memory above 48 KB needs an opt-in that CUDA keeps per FUNCTION and per DEVICE, not per handle: the largest size asked
// for so far is remembered process-wide and only ever raised, so that a second handle on the same GPU can never lower what an earlier
// one relies on.  false: `smem` exceeds the device's opt-in limit or the driver refused.
bool orbx_optin_smem(orbx_handle *h, const void *func, size_t smem);
size_t orbx_quadtree_smem(int node_cap);              // k_quadtree.cu: dynamic shared memory of one quadtree CTA

// device status bits
#define ORBX_DS_CAND_OVERFLOW 1
#define ORBX_DS_NODE_OVERFLOW 2
#define ORBX_DS_KP_OVERFLOW   4
#define ORBX_DS_INTERNAL     16      // a device-side wait ran into its bound (k_fast_dense)
#define ORBX_DS_BAD_INDEX     8      // a caller-supplied index list pointed outside its array (k_cull)

// ---- profile C (cv::ORB), orbx_cvorb.cu ----
orbx_status orbx_cvorb_create(orbx_handle *h);
void orbx_cvorb_destroy(orbx_handle *h);
orbx_status orbx_cvorb_run(orbx_handle *h, const uint8_t *d_gray, int width, int height, size_t step,
                           orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_count);
orbx_status orbx_cvorb_get_level(orbx_handle *h, int level, int blurred, uint8_t *out, size_t out_step, const uint8_t *l0, size_t l0_step);
void orbx_cvorb_level_size(const orbx_handle *h, int w, int hgt, int level, int *lw, int *lh);

// ---- kernel launchers (one per .cu) ----
int  launch_pyramid(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride);
int  launch_resize_level(orbx_handle *h, int level, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride);
void orbx_build_fast_cells(const FrameGeom &G, const std::vector<uint32_t> &ctab, std::vector<uint4> &out);   // k_fast.cu: 32-byte FAST cell records
int  orbx_ensure_tmaps(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride);
int  launch_fast(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride);   // -1: TMA descriptor encode failed
int  launch_fast_cells(orbx_handle *h, int nframes, const int32_t *d_items, const int32_t *d_nitems);   // k_fast.cu: every (frame, cell), or the listed ones at minThFAST
int  launch_fast_dense(orbx_handle *h, int nframes);                                                    // k_fast_dense.cu
void orbx_build_fast_dense(const FrameGeom &G, int cand_divisor, DenseGeom &D, std::vector<uint4> *tiles);
bool orbx_encode_level(CUtensorMap *m, const uint8_t *base, size_t pitch, int rows, size_t fstride, int frames, int box_rows, int box_words);
int  launch_quadtree(orbx_handle *h, int nframes);     // -1: the node table does not fit the shared-memory opt-in limit
int  launch_quadtree_geo(orbx_handle *h, const FrameGeom *d_geo, int nlevels, int nframes, int node_cap, size_t cand_slab, int sel_slab);
int  launch_cull(orbx_handle *h, const orbx_keypoint *d_kps, const uint8_t *d_desc, int n, const int32_t *d_mq, int nm, int max_new, float min_response,
                 orbx_keypoint *d_out_kps, uint8_t *d_out_desc, int32_t *d_out_index, int cap, int32_t *d_n_out);   // k_cull.cu; -1: shared memory opt-in failed
int  launch_blur(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride, cudaStream_t st, int tile_first = 0, int ntiles = -1);
// d_map (nullable, fused kernel only): filter-first order — launch_keep_list has left, per frame, the indices (in the selected list) of the
// keypoints the depth / box filter keeps and their count in d_counts; only those are described, straight into their final rows
void launch_describe_to(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride,
                        orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_counts, const int32_t *d_map = nullptr, int map_slab = 0);
void launch_keep_list(orbx_handle *h, int nframes, const uint16_t *d_depth, size_t dstep, size_t dfstride,
                      const orbx_box *d_boxes, const int32_t *d_box_offsets, int box_base, int nboxes, uint64_t drop_mask,
                      int32_t *d_map, int map_slab, int32_t *d_counts, int cap);                  // k_filter.cu
void upload_umax(const int *umax);
void launch_filter(orbx_handle *h, int nframes, const uint16_t *d_depth, size_t dstep, size_t dfstride,
                   const orbx_box *d_boxes, const int32_t *d_box_offsets, int box_base, int nboxes, uint64_t drop_mask,
                   orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_counts);
int  launch_match_core(orbx_handle *h, const uint8_t *d_q, const int32_t *d_nq, int nq_max, size_t q_stride,
                       const uint8_t *d_t, const int32_t *d_nt, int nt_max, size_t t_stride,
                       const int32_t *d_qsel, const int32_t *d_tsel, int nproblems, uint32_t row_base,
                       int k, float max_dist, int ratio_num, orbx_dmatch *d_out, size_t out_stride, int32_t *d_n_out,
                       orbx_top2 *d_top2);
void launch_match_radius(orbx_handle *h, const uint8_t *d_q, int nq, const uint8_t *d_t, int nt, uint32_t row_base,
                         float max_dist, orbx_dmatch *d_out, int cap, int32_t *d_n_out);
int  launch_assoc(orbx_handle *h, const uint8_t *d_q, const float *d_qpx, int nq, const uint8_t *d_t, const float *d_pos, int nt, uint32_t row_base,
                  const orbx_pose *pose, float max_dist, double max_err, orbx_assoc *d_out);
void launch_assoc_merge(orbx_handle *h, const orbx_assoc *d_parts, int nparts, int nq, orbx_assoc *d_out);
void launch_assoc_merge_strided(orbx_handle *h, const orbx_assoc *d_parts, size_t stride, int nparts, int nq, orbx_assoc *d_out);
void launch_pack_keyframe(orbx_handle *h, int nframes, const orbx_keypoint *d_kps, const uint8_t *d_desc, const int32_t *d_counts, int cap_in,
                          const uint16_t *d_depth, size_t dstep, size_t dfstride, int dw, int dh, const orbx_kfparams *K,
                          orbx_kfrecord *d_out, int32_t *d_nout, int cap_out);
void launch_harris(orbx_handle *h, const uint8_t *img, size_t step, int w, int hgt, const int32_t *d_xy, int n, int bs, float k, float *d_out);
void launch_bgr2gray(orbx_handle *h, const uint8_t *d_bgr, size_t sstep, size_t sfstride, uint8_t *d_gray, size_t dstep, size_t dfstride,
                     int w, int hgt, int nframes, cudaStream_t st);
void launch_merge_top2(orbx_handle *h, const orbx_top2 *d_parts, int nshards, int nq, orbx_top2 *d_out);
void launch_merge_top2_strided(orbx_handle *h, const orbx_top2 *d_parts, size_t stride, int nshards, int nq, orbx_top2 *d_out);
void launch_synth_gray(orbx_handle *h, uint32_t seed, int first, int n, int w, int hh, uint8_t *d, size_t step, size_t fstride);
void launch_synth_depth(orbx_handle *h, uint32_t seed, int first, int n, int w, int hh, uint16_t *d, size_t step, size_t fstride);
void launch_synth_desc(orbx_handle *h, uint32_t seed, uint64_t first_row, int64_t nrows, uint8_t *d);
void launch_test_trig(orbx_handle *h, const float *d_in, int n, float *d_c, float *d_s);
void launch_test_atan2(orbx_handle *h, const float *d_y, const float *d_x, int n, float *d_o);
void launch_trig_checksum(orbx_handle *h, uint32_t first, uint32_t last, unsigned long long *d_sums);
double run_popc_bench(orbx_handle *h);
void launch_fmat_hypotheses(orbx_handle *h, const float *d_p1, const float *d_p2, int n, int nh, uint32_t seed, double *d_F);
void launch_pnp_score(orbx_handle *h, const float *d_p3, const float *d_p2, int n, const double *d_Rt, int nh, double fx, double fy, double cx, double cy,
                      float t2, int32_t *d_counts, uint8_t *d_masks, int32_t *d_best, uint8_t *d_best_mask);   // k_ransac.cu
void launch_pnp_points(orbx_handle *h, const orbx_keypoint *d_prev, int nprev, const orbx_keypoint *d_curr, int ncurr, const orbx_dmatch *d_m, int nm,
                       const uint16_t *d_depth, int w, int hgt, size_t dstep, float fx, float fy, float cx, float cy, float *d_p3, float *d_p2, int32_t *d_n);
void launch_fmat_score(orbx_handle *h, const float *d_p1, const float *d_p2, int n, const double *d_F, int nh, float t2,
                       int32_t *d_counts, uint8_t *d_masks, int32_t *d_best, uint8_t *d_best_mask);
