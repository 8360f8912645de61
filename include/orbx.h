/* orbx.h — C ABI of the B200-native ORB extract + Hamming match path.
 *
 * Drop-in boundary for the ONE data-parallel hot path of andrewkwolek/dynamic-visual-slam
 * (paths below are relative to the reference's dynamic_visual_slam/ directory):
 *
 *   seam 1  ORB_SLAM3::ORBextractor                    include/dynamic_visual_slam/ORBextractor.hpp:44-111
 *           ctor (nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)            ORBextractor.hpp:50-51
 *           int operator()(image, mask, keypoints, descriptors, vLappingArea)       ORBextractor.hpp:58-60
 *           getters + public mvImagePyramid                                         ORBextractor.hpp:62-84
 *           callers: frontend.cpp:1094, :1285
 *   seam 2  cv::BFMatcher(NORM_HAMMING).match(query, train, matches) + `distance < 50`
 *           callers: frontend.cpp:1123-1132, :614-623 ; backend.cpp:1072-1076
 *   post-filters fused behind the same call: Frontend::filterDepth (frontend.cpp:457-527) and
 *           Backend::categorizeObservation + filtered_objects_ (backend.cpp:1011-1029, 746-751)
 *   backend association, descriptor stage: Backend::associateObservation  backend.cpp:1064-1083
 *
 * The reference has no FFI layer; this header IS the binding surface a maintainer would call from
 * the two nodes (see INTEGRATION.md and dynamic-visual-slam_b200/host/ORBextractor.hpp for the
 * C++ adapter that reproduces the two call shapes over it).  Plain pointers and sizes only; no
 * torch / OpenCV / CUDA types.  Status codes, never exceptions or aborts.  One handle = one CUDA
 * device + one stream, single caller at a time (the reference's extractor is not re-entrant either).
 *
 * Memory-space convention: functions ending in `_device` take DEVICE pointers and are asynchronous
 * on the handle's stream (call orbx_sync); all others take HOST pointers and are synchronous.
 */
#ifndef ORBX_H
#define ORBX_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_MAX_LEVELS 16
#define ORBX_DESC_BYTES 32

typedef enum {
    ORBX_OK = 0,
    ORBX_E_INVALID = 1,      /* bad argument */
    ORBX_E_CUDA = 2,         /* CUDA runtime error, see orbx_last_error */
    ORBX_E_CAPACITY = 3,     /* an output or an internal candidate list was too small; nothing truncated silently */
    ORBX_E_EMPTY = 4,        /* empty image: the reference's operator() returns -1 here (ORBextractor.cpp:1090) */
    ORBX_E_NOMEM = 5,
    ORBX_E_UNSUPPORTED = 6   /* e.g. aspect ratio for which the reference itself divides by zero */
} orbx_status;

/* 28-byte cv::KeyPoint layout (pt.x, pt.y, size, angle, response, octave, class_id) */
typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } orbx_keypoint;
/* 16-byte cv::DMatch layout (queryIdx, trainIdx, imgIdx, distance) */
typedef struct { int32_t queryIdx, trainIdx, imgIdx; float distance; } orbx_dmatch;
/* one YOLO detection box, yolo_msgs bbox convention: centre + size in pixels (backend.cpp:1017-1020) */
typedef struct { double cx, cy, w, h; int32_t class_id; int32_t pad_; } orbx_box;
/* per-shard top-2 of a landmark-database query (association, backend.cpp:1064-1083) */
typedef struct { uint32_t dist0, idx0, dist1, idx1; } orbx_top2;
/* camera pose and intrinsics of Backend::reprojectPoint (backend.cpp:1153-1173): point_camera = R.t() * (X - t), R row-major */
typedef struct { double R[9]; double t[3]; double fx, fy, cx, cy; } orbx_pose;
/* keyframe packing (Frontend::publishKeyframe, frontend.cpp:731-776): float intrinsics (frontend.cpp:278) and the camera-to-world pose */
typedef struct { double R[9]; double t[3]; float fx, fy, cx, cy; } orbx_kfparams;
/* one Landmark + Observation pair of Keyframe.msg (msg/Landmark.msg:5-8, msg/Observation.msg:5-12): 80 bytes */
typedef struct { uint64_t landmark_id; double position[3]; double pixel_x, pixel_y; uint8_t descriptor[32]; } orbx_kfrecord;
/* result of the reprojection-gated association of one observation: landmark = global row or -1 (backend.cpp:1064-1120) */
typedef struct { double reproj_error; int32_t landmark; float distance; } orbx_assoc;

typedef struct {
    /* ORBextractor ctor arguments — defaults are the reference's literals, frontend.cpp:205-211 */
    int32_t nfeatures;        /* 1000 */
    float   scale_factor;     /* 1.2f */
    int32_t nlevels;          /* 8    */
    int32_t ini_th_fast;      /* 20   */
    int32_t min_th_fast;      /* 7    */
    /* Frontend depth validity range in metres — frontend.cpp:241-242 */
    float   depth_min;        /* 0.3f */
    float   depth_max;        /* 3.0f */
    /* arena sizing */
    int32_t max_width;        /* 1280 */
    int32_t max_height;       /* 720  */
    int32_t max_batch;        /* frames resident per batch call (1) */
    int32_t max_keypoints;    /* per-frame output capacity; 0 = nfeatures + 4*nlevels + 64 */
    int32_t cand_divisor;     /* candidate-list capacity per level = max(4096, pixels/cand_divisor); 0 = 16 */
    int32_t device;           /* CUDA device ordinal */
    int32_t host_chunk;       /* frames per pipeline chunk of the host-buffer batch calls, 0 = auto (max_batch/4, 1..32) */
    int32_t profile;          /* ORBX_PROFILE_SLAM (0, default) | ORBX_PROFILE_CVORB (1), see below */
    int32_t reserved_;
} orbx_params;

/* Extraction profiles: the two ORB extractors the reference uses.
 *   ORBX_PROFILE_SLAM   the frontend's ORB_SLAM3::ORBextractor (ORBextractor.cpp; frontend.cpp:1094): INTER_LINEAR pyramid, per-cell FAST with the
 *                       20 -> 7 retry, quadtree distribution, FAST score as the response, fixed-point blur.  Bit-exact.
 *   ORBX_PROFILE_CVORB  cv::ORB (test/test_dbow2_integration.cpp:19,38; BASELINE configs[0]; north_star stage 3): INTER_LINEAR_EXACT pyramid,
 *                       whole-level FAST(ini_th_fast) + retainBest(2N), Harris response (block 7, k 0.04) + retainBest(N), float-path blur.
 *                       min_th_fast is unused.  Inside a level the keypoints come sorted by (response descending, y, x) — OpenCV leaves
 *                       std::nth_element's order — so parity with cv::ORB is on sets; descriptors follow the AVX2/FMA build of OpenCV 4.13
 *                       (>= 99.9 % of rows identical against other builds, as north_star states).  One frame at a time: the batch entry
 *                       points loop over the frames; stage access and the stream (track) calls work as for the other profile.        */
#define ORBX_PROFILE_SLAM 0
#define ORBX_PROFILE_CVORB 1

typedef struct orbx_handle orbx_handle;
typedef struct orbx_db orbx_db;

void        orbx_default_params(orbx_params *p);
orbx_status orbx_create(const orbx_params *p, orbx_handle **out);
void        orbx_destroy(orbx_handle *h);
const char *orbx_last_error(const orbx_handle *h);   /* h may be NULL: last error of a failed create */
const char *orbx_version(void);
orbx_status orbx_sync(orbx_handle *h);
void       *orbx_stream(orbx_handle *h);              /* the handle's cudaStream_t */

/* ---- ORBextractor getters (ORBextractor.hpp:62-82) ---- */
int32_t orbx_get_levels(const orbx_handle *h);
float   orbx_get_scale_factor(const orbx_handle *h);
void    orbx_get_scale_factors(const orbx_handle *h, float *out /*nlevels*/);
void    orbx_get_inverse_scale_factors(const orbx_handle *h, float *out);
void    orbx_get_scale_sigma_squares(const orbx_handle *h, float *out);
void    orbx_get_inverse_scale_sigma_squares(const orbx_handle *h, float *out);
void    orbx_get_features_per_level(const orbx_handle *h, int32_t *out);
/* level geometry for a given input size (ComputePyramid, ORBextractor.cpp:1173-1174) */
orbx_status orbx_level_size(const orbx_handle *h, int32_t width, int32_t height, int32_t level,
                            int32_t *lw, int32_t *lh);

/* ---- extraction: ORBextractor::operator() (ORBextractor.cpp:1086-1167) ----
 * gray: CV_8UC1 rows of `step` bytes.  depth (nullable): CV_16UC1 millimetres, rows of `dstep` BYTES;
 * when given, keypoints without valid depth are dropped after selection (Frontend::filterDepth).
 * boxes (nullable) + drop_class_mask: keypoints whose first containing box has a class_id c with
 * bit c set in the mask are dropped after selection (categorizeObservation + filtered_objects_).
 * Outputs are caller-allocated with capacity `cap`; *n_out receives the count
 * (ORBX_E_CAPACITY if it would exceed cap).  Featureless frames give *n_out = 0 and ORBX_OK.     */
orbx_status orbx_extract(orbx_handle *h, const uint8_t *gray, int32_t width, int32_t height, size_t step,
                         orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *n_out);
orbx_status orbx_extract_filtered(orbx_handle *h, const uint8_t *gray, int32_t width, int32_t height, size_t step,
                                  const uint16_t *depth, size_t dstep,
                                  const orbx_box *boxes, int32_t nboxes, uint64_t drop_class_mask,
                                  orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *n_out);

/* ---- ingest: cv::cvtColor(BGR2GRAY) on the device (reference frontend.cpp:1084, the step before the hot path) ----
 * orbx_bgr2gray_device converts nframes packed-BGR (CV_8UC3) device frames into gray frames (dst step % 4 == 0), asynchronously.
 * orbx_extract_bgr = H2D of ONE host BGR frame + conversion + orbx_extract_filtered's pipeline (depth / boxes nullable).      */
orbx_status orbx_bgr2gray_device(orbx_handle *h, const uint8_t *d_bgr, int32_t nframes, int32_t width, int32_t height, size_t step,
                                 size_t frame_stride, uint8_t *d_gray, size_t gray_step, size_t gray_frame_stride);
orbx_status orbx_extract_bgr(orbx_handle *h, const uint8_t *bgr, int32_t width, int32_t height, size_t step,
                             const uint16_t *depth, size_t dstep, const orbx_box *boxes, int32_t nboxes, uint64_t drop_class_mask,
                             orbx_keypoint *kps, uint8_t *desc, int32_t cap, int32_t *n_out);

/* frame-parallel batch, HOST buffers: frames tightly packed (frame f at gray + f*height*step).
 * depth nullable.  kps/desc have room for cap_per_frame entries per frame; counts[nframes].
 * Internally cut into chunks that are pipelined over three streams (H2D of chunk i+1, kernels of chunk i,
 * D2H of chunk i-1).  A depth buffer in pinned host memory (orbx_alloc_pinned / cudaHostRegister) is not
 * copied at all: the depth filter gathers its samples from it in place.                            */
orbx_status orbx_extract_batch(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                               size_t step, const uint16_t *depth, size_t dstep,
                               orbx_keypoint *kps, uint8_t *desc, int32_t cap_per_frame, int32_t *counts);

/* frame-parallel batch, DEVICE buffers, asynchronous.  nframes <= max_batch.  Frame f starts at
 * d_gray + f*frame_stride (bytes); rows are `step` bytes, step % 16 == 0 and 16-byte aligned base.
 * d_depth nullable (frame stride dframe_stride BYTES, row step dstep BYTES).  d_kps / d_desc hold
 * cap_per_frame entries per frame; d_counts[nframes].  Errors detected on the device (capacity) are
 * reported by the next orbx_sync.  Per-frame YOLO boxes: the *_boxes variants below.              */
orbx_status orbx_extract_batch_device(orbx_handle *h, const uint8_t *d_gray, int32_t nframes,
                                      int32_t width, int32_t height, size_t step, size_t frame_stride,
                                      const uint16_t *d_depth, size_t dstep, size_t dframe_stride,
                                      orbx_keypoint *d_kps, uint8_t *d_desc, int32_t cap_per_frame,
                                      int32_t *d_counts);

/* ---- per-frame YOLO box lists in the batch calls (BASELINE configs[4]: semantic-masked extraction) ----
 * Backend::categorizeObservation + the filtered_objects_ test (backend.cpp:1011-1029, 746-751) for every frame of a batch: frame f owns
 * boxes[box_offsets[f] .. box_offsets[f+1]) (nframes + 1 non-decreasing offsets; nboxes_total = box_offsets[nframes] for the device
 * variants); a keypoint whose FIRST containing box (inclusive bounds, fp64 compares) has a class_id c with bit c of drop_class_mask set
 * is dropped — after selection, so the retained set equals extract-then-filter.  boxes == NULL: no box filter.
 * Otherwise identical to the calls without `_boxes`.                                                                              */
orbx_status orbx_extract_batch_boxes_device(orbx_handle *h, const uint8_t *d_gray, int32_t nframes,
                                            int32_t width, int32_t height, size_t step, size_t frame_stride,
                                            const uint16_t *d_depth, size_t dstep, size_t dframe_stride,
                                            const orbx_box *d_boxes, const int32_t *d_box_offsets, int32_t nboxes_total, uint64_t drop_class_mask,
                                            orbx_keypoint *d_kps, uint8_t *d_desc, int32_t cap_per_frame, int32_t *d_counts);
orbx_status orbx_extract_batch_boxes(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                     size_t step, const uint16_t *depth, size_t dstep,
                                     const orbx_box *boxes, const int32_t *box_offsets, uint64_t drop_class_mask,
                                     orbx_keypoint *kps, uint8_t *desc, int32_t cap_per_frame, int32_t *counts);

/* ---- stream step: the hot part of Frontend::syncCallback (frontend.cpp:1094-1132, first frame :1277-1317) ----
 * For every frame of the batch: extraction, depth filter (d_depth / depth nullable), then
 * matcher_.match(filtered_descriptors, prev_descriptors_) + `distance < max_dist` against the PREVIOUS
 * frame's filtered descriptors.  Frame 0 of a call is matched against the handle's carried state (the
 * last frame of the previous call; none after create / orbx_track_reset => 0 matches, as on the
 * reference's first frame).  matches: cap_per_frame entries per frame, query order; match_counts[nframes].
 * max_dist <= 0 keeps every match (exactly match()).                                               */
orbx_status orbx_track_batch_device(orbx_handle *h, const uint8_t *d_gray, int32_t nframes,
                                    int32_t width, int32_t height, size_t step, size_t frame_stride,
                                    const uint16_t *d_depth, size_t dstep, size_t dframe_stride,
                                    orbx_keypoint *d_kps, uint8_t *d_desc, int32_t cap_per_frame, int32_t *d_counts,
                                    orbx_dmatch *d_matches, int32_t *d_match_counts, float max_dist);
orbx_status orbx_track_batch(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                             size_t step, const uint16_t *depth, size_t dstep,
                             orbx_keypoint *kps, uint8_t *desc, int32_t cap_per_frame, int32_t *counts,
                             orbx_dmatch *matches, int32_t *match_counts, float max_dist);
/* the same with per-frame YOLO boxes (see orbx_extract_batch_boxes_device): descriptors of dropped keypoints never reach the matcher */
orbx_status orbx_track_batch_boxes_device(orbx_handle *h, const uint8_t *d_gray, int32_t nframes,
                                          int32_t width, int32_t height, size_t step, size_t frame_stride,
                                          const uint16_t *d_depth, size_t dstep, size_t dframe_stride,
                                          const orbx_box *d_boxes, const int32_t *d_box_offsets, int32_t nboxes_total, uint64_t drop_class_mask,
                                          orbx_keypoint *d_kps, uint8_t *d_desc, int32_t cap_per_frame, int32_t *d_counts,
                                          orbx_dmatch *d_matches, int32_t *d_match_counts, float max_dist);
orbx_status orbx_track_batch_boxes(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                   size_t step, const uint16_t *depth, size_t dstep,
                                   const orbx_box *boxes, const int32_t *box_offsets, uint64_t drop_class_mask,
                                   orbx_keypoint *kps, uint8_t *desc, int32_t cap_per_frame, int32_t *counts,
                                   orbx_dmatch *matches, int32_t *match_counts, float max_dist);
void        orbx_track_reset(orbx_handle *h);

/* Asynchronous variants of the two host-buffer batch calls: enqueue ONE batch (1..max_batch frames) and return a ticket;
 * orbx_batch_wait(ticket) blocks until that batch's results are in the caller's buffers and reports its status.  Two batches
 * may be in flight, collected in submission order: submit(k+1) before wait(k) hides the H2D of batch k+1 and the D2H of batch
 * k-1 under the kernels of batch k.  Input and output buffers must stay valid until the wait returns and should be pinned
 * (orbx_alloc_pinned).  While a ticket is outstanding only *_submit / orbx_batch_wait may be called on the handle.        */
orbx_status orbx_extract_batch_submit(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                      size_t step, const uint16_t *depth, size_t dstep,
                                      orbx_keypoint *kps, uint8_t *desc, int32_t cap_per_frame, int32_t *counts, int32_t *ticket);
orbx_status orbx_track_batch_submit(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                    size_t step, const uint16_t *depth, size_t dstep,
                                    orbx_keypoint *kps, uint8_t *desc, int32_t cap_per_frame, int32_t *counts,
                                    orbx_dmatch *matches, int32_t *match_counts, float max_dist, int32_t *ticket);
orbx_status orbx_track_batch_boxes_submit(orbx_handle *h, const uint8_t *gray, int32_t nframes, int32_t width, int32_t height,
                                          size_t step, const uint16_t *depth, size_t dstep,
                                          const orbx_box *boxes, const int32_t *box_offsets, uint64_t drop_class_mask,
                                          orbx_keypoint *kps, uint8_t *desc, int32_t cap_per_frame, int32_t *counts,
                                          orbx_dmatch *matches, int32_t *match_counts, float max_dist, int32_t *ticket);
orbx_status orbx_batch_wait(orbx_handle *h, int32_t ticket);

/* ---- matching: cv::BFMatcher(NORM_HAMMING) ----
 * k = 1: BFMatcher::match.   max_dist <= 0: one DMatch per query, query order (exactly match()).
 *                            max_dist  > 0: only matches with distance < max_dist, query order
 *                                           (the frontend's `distance < 50` loop, frontend.cpp:1126-1132).
 * k = 2: BFMatcher::knnMatch(k=2).  ratio <= 0: two DMatch per query (out[2q], out[2q+1]).
 *                            ratio  > 0: Lowe test, keeps the best match iff d0 < ratio*d1 evaluated
 *                                        exactly in integers as d0*den < num*d1 with ratio = num/den
 *                                        rounded to 1/1024; 0.75 is exact.  max_dist also applies if > 0.
 * Ties resolve to the lowest trainIdx, as BFMatcher does.  out capacity must be nq*k.            */
orbx_status orbx_match(orbx_handle *h, const uint8_t *query, int32_t nq, const uint8_t *train, int32_t nt,
                       int32_t k, float max_dist, float ratio, orbx_dmatch *out, int32_t *n_out);
orbx_status orbx_match_device(orbx_handle *h, const uint8_t *d_query, int32_t nq, const uint8_t *d_train, int32_t nt,
                              int32_t k, float max_dist, float ratio, orbx_dmatch *d_out, int32_t *d_n_out);
/* stream matching over a batch: for every pair p in [0,npairs): query = frame q_frame[p], train = frame
 * t_frame[p] of the (d_desc, d_counts, cap_per_frame) arrays produced by orbx_extract_batch_device.
 * Outputs: d_out[p*cap_per_frame*k ...], d_n_out[p].  Host index arrays.                          */
orbx_status orbx_match_pairs_device(orbx_handle *h, const uint8_t *d_desc, const int32_t *d_counts,
                                    int32_t cap_per_frame, const int32_t *q_frame, const int32_t *t_frame,
                                    int32_t npairs, int32_t k, float max_dist, float ratio,
                                    orbx_dmatch *d_out, int32_t *d_n_out);

/* ---- landmark database for Backend::associateObservation (descriptor stage) ----
 * A shard holds `rows` 32-byte descriptors whose GLOBAL landmark indices are first_index ..
 * first_index+rows-1.  Query returns, per query, the two nearest rows of THIS shard as
 * (distance, global index), ties to the lowest index.  Shards on different GPUs are merged with
 * orbx_merge_top2 after an all-gather of the per-shard results.                   */
orbx_status orbx_db_create(orbx_handle *h, int64_t capacity_rows, uint32_t first_index, orbx_db **out);
void        orbx_db_destroy(orbx_db *db);
orbx_status orbx_db_append(orbx_db *db, const uint8_t *rows_host, int64_t nrows);
orbx_status orbx_db_append_device(orbx_db *db, const uint8_t *d_rows, int64_t nrows);
int64_t     orbx_db_rows(const orbx_db *db);
orbx_status orbx_db_query_top2_device(orbx_db *db, const uint8_t *d_query, int32_t nq, orbx_top2 *d_out);
orbx_status orbx_db_query_top2(orbx_db *db, const uint8_t *query, int32_t nq, orbx_top2 *out);
/* merge nshards per-shard results laid out [shard][nq] into out[nq] (lexicographic (dist, idx) min) */
orbx_status orbx_merge_top2_device(orbx_handle *h, const orbx_top2 *d_parts, int32_t nshards, int32_t nq,
                                   orbx_top2 *d_out);
/* reference semantics of the descriptor stage: every row with distance < max_dist is a candidate
 * (backend.cpp:1074-1076).  Writes up to cap (query, global idx, distance) triples as orbx_dmatch
 * (queryIdx, trainIdx = global idx, imgIdx = 0, distance), sorted by (queryIdx, trainIdx).        */
orbx_status orbx_db_query_radius(orbx_db *db, const uint8_t *query, int32_t nq, float max_dist,
                                 orbx_dmatch *out, int32_t cap, int32_t *n_out);

/* ---- Backend::associateObservation as one call (backend.cpp:1064-1120): descriptor candidates (distance < max_desc_dist), then the
 * candidate with the smallest reprojection error, if that error is < max_reproj_err.  Landmark positions (float xyz per row, the
 * reference's cv::Point3f) are attached to the shard's rows with orbx_db_set_positions.  query_px: nq x 2 floats (obs.pixel).
 * All observations of the call see ONE snapshot of the positions (the reference re-triangulates a landmark right after each
 * association, backend.cpp:772; a caller that needs that ordering re-submits the affected observations).  Equal errors: lowest row.
 * Sharded databases: per-shard results -> all-gather -> orbx_merge_assoc_device.                                                  */
orbx_status orbx_db_set_positions(orbx_db *db, int64_t first_row, int64_t nrows, const float *xyz_host);
orbx_status orbx_db_set_positions_device(orbx_db *db, int64_t first_row, int64_t nrows, const float *d_xyz);
orbx_status orbx_db_associate(orbx_db *db, const uint8_t *query, const float *query_px, int32_t nq, const orbx_pose *pose,
                              float max_desc_dist /*50*/, double max_reproj_err /*5*/, orbx_assoc *out);
orbx_status orbx_db_associate_device(orbx_db *db, const uint8_t *d_query, const float *d_query_px, int32_t nq, const orbx_pose *pose,
                                     float max_desc_dist, double max_reproj_err, orbx_assoc *d_out);
orbx_status orbx_merge_assoc_device(orbx_handle *h, const orbx_assoc *d_parts, int32_t nshards, int32_t nq, orbx_assoc *d_out);

/* ---- sharded landmark database over the GPUs of one box: the ONE collective of the path (BASELINE configs[3]) ----
 * One process (or thread) per GPU, each with its own handle and its shard (orbx_db_create with first_index = global index of its row 0).
 * orbx_comm wraps an NCCL communicator owned by the library (NCCL is loaded with dlopen at first use; ORBX_E_UNSUPPORTED if absent):
 * rank 0 calls orbx_comm_get_unique_id and the caller carries the 128 bytes to the other ranks by any means, then EVERY rank calls
 * orbx_comm_create (collective).  The *_sharded_device calls are collective too (same nq on every rank, the queries replicated) and
 * enqueue per-shard kernel -> ncclAllGather (nq x 16 B per rank) -> merge kernel on the handle's stream: no host synchronisation
 * between query and merged result; every rank ends with the global answer (lexicographic (distance, index) resp. (error, index)
 * minimum over the shards = what one unsharded database returns, BFMatcher's lowest-index tie-break included).                    */
typedef struct orbx_comm orbx_comm;
#define ORBX_COMM_ID_BYTES 128
orbx_status orbx_comm_get_unique_id(void *id128);
const char *orbx_comm_last_error(void);
orbx_status orbx_comm_create(orbx_handle *h, int32_t nranks, int32_t rank, const void *id128, orbx_comm **out);
void        orbx_comm_destroy(orbx_comm *c);
int32_t     orbx_comm_ranks(const orbx_comm *c);
int32_t     orbx_comm_rank(const orbx_comm *c);
/* Transport of the exchange step.  0 (default): peer memory — every rank stores its block straight into a mailbox in each peer's HBM
 * over NVLink (CUDA IPC mappings set up by orbx_comm_create) and waits on sequence flags inside one small kernel; used when all ranks
 * could map each other (orbx_comm_peer_memory() == 1) and the block fits the 64 KB slot, else NCCL.  1: always ncclAllGather.  */
orbx_status orbx_comm_set_transport(orbx_comm *c, int32_t transport);
int32_t     orbx_comm_peer_memory(const orbx_comm *c);
orbx_status orbx_db_query_top2_sharded_device(orbx_db *db, orbx_comm *c, const uint8_t *d_query, int32_t nq, orbx_top2 *d_out);
orbx_status orbx_db_associate_sharded_device(orbx_db *db, orbx_comm *c, const uint8_t *d_query, const float *d_query_px, int32_t nq,
                                             const orbx_pose *pose, float max_desc_dist, double max_reproj_err, orbx_assoc *d_out);

/* ---- feature culling for the backend: Frontend::syncCallback, frontend.cpp:1168-1218, SURVEY §8(f) rank 3 ----
 * The set the frontend hands to isKeyframe / publishKeyframe: the keypoint of every (geometrically consistent) match, in match order,
 * then the unmatched keypoints sorted by response — std::sort with `a.first > b.first`, so equal responses stay where libstdc++'s
 * introsort leaves them — while fewer than max_new (reference: 200) were added and response >= min_response (reference: 50).
 * match_query = DMatch::queryIdx of the matches that survived the caller's RANSAC mask (the F-matrix RANSAC stays on the CPU).
 * out_index (optional) receives the index of every output element in the input arrays.  n <= 8192.  The keyframe-criterion match of
 * isKeyframe (frontend.cpp:601-662) is orbx_match(k=1, max_dist=50) against the previous keyframe's culled descriptors.
 * Device variant: asynchronous on the handle's stream; a query index outside [0, n) surfaces as ORBX_E_INVALID at the next sync.   */
orbx_status orbx_cull_keyframe(orbx_handle *h, const orbx_keypoint *kps, const uint8_t *desc, int32_t n,
                               const int32_t *match_query, int32_t n_matches, int32_t max_new /*200*/, float min_response /*50*/,
                               orbx_keypoint *out_kps, uint8_t *out_desc, int32_t *out_index, int32_t cap, int32_t *n_out);
orbx_status orbx_cull_keyframe_device(orbx_handle *h, const orbx_keypoint *d_kps, const uint8_t *d_desc, int32_t n,
                                      const int32_t *d_match_query, int32_t n_matches, int32_t max_new, float min_response,
                                      orbx_keypoint *d_out_kps, uint8_t *d_out_desc, int32_t *d_out_index, int32_t cap, int32_t *d_n_out);

/* ---- geometric validation, hypothesis scoring: SURVEY §8(f) rank 4 ----
 * The K x N evaluation inside cv::findFundamentalMat(prev_pts, curr_pts, mask, cv::FM_RANSAC, 2.0, 0.99) (frontend.cpp:1134-1154, :625-645):
 * for each of nh fundamental-matrix hypotheses F (row-major 3x3 doubles; x2' F x1 = 0 with pts1 = prev_pts, pts2 = curr_pts) the symmetric
 * epipolar error OpenCV's RANSAC uses (double precision, its operation order), inlier <=> (float)err <= (float)(threshold^2).
 * inlier_counts[nh]; *best = hypothesis with the most inliers (ties: lowest index); best_mask[n] = its inlier mask, bit-identical to the
 * mask OpenCV computes for that F.  Sampling and the minimal solves stay with the caller (PnP / F-matrix RANSAC are the reference's,
 * north_star).  inlier_counts, best_mask may be NULL.                                                                          */
orbx_status orbx_fmat_score(orbx_handle *h, const float *pts1, const float *pts2, int32_t n, const double *F, int32_t nh, double threshold,
                            int32_t *inlier_counts, int32_t *best, uint8_t *best_mask);

/* The whole geometric validation on the device: nh minimal samples (8 distinct correspondences each, counter-based generator from `seed`)
 * -> normalised 8-point fundamental matrices (fp64, rank 2 enforced, F(2,2) = 1) -> OpenCV's scoring as above -> the model with the most inliers.
 * NOT a restatement of cv::findFundamentalMat (its 7-point samples come from OpenCV's own RNG and cannot be reproduced); what holds by
 * construction is the contract of its result: best_mask is exactly the inlier set OpenCV's error function gives for F_out at `threshold`.
 * n >= 8, nh >= 1.  F_out: row-major 3x3.  n_inliers may be NULL.                                                                */
orbx_status orbx_fmat_ransac(orbx_handle *h, const float *pts1, const float *pts2, int32_t n, int32_t nh, double threshold, uint32_t seed,
                             double *F_out, uint8_t *best_mask, int32_t *n_inliers);

/* ---- pose estimation, the data-parallel parts of Frontend::estimateCameraPose (frontend.cpp:843-960) ----
 * orbx_pnp_points: the 3D-2D correspondence loop (:858-892).  For every match, in match order: the previous frame's keypoint (trainIdx) is
 * back-projected with the depth at its rounded pixel (std::round), kept iff the pixel is inside the image and 0.3 < d <= 3.0 m;
 * X = (u - cx) * d / fx in float (the reference's intrinsics are float members).  pts3d [nm * 3], pts2d [nm * 2] (current frame's keypoint, queryIdx).
 * A match index outside its keypoint array: ORBX_E_INVALID.
 * orbx_pnp_score: cv::solvePnPRansac's scoring loop (:911-923, PnPRansacCallback::computeError): nh pose hypotheses, each 12 doubles (R row-major =
 * cv::Rodrigues(rvec), then t), against all correspondences — cv::projectPoints in double without distortion, projections stored as float, squared
 * pixel distance in float, inlier <=> err <= (float)(threshold * threshold) (the reference passes 4.0).  inlier_counts [nh] (nullable), *best = the
 * hypothesis with the most inliers (lowest index on ties), best_mask [n] (nullable).  The mask of a given pose equals OpenCV's bit for bit
 * (tests/test_pnp.py pins the arithmetic against cv2.projectPoints).  Minimal-sample solves (P3P / EPnP) and the final refinement stay with the caller. */
orbx_status orbx_pnp_points(orbx_handle *h, const orbx_keypoint *prev_kps, int32_t n_prev, const orbx_keypoint *curr_kps, int32_t n_curr,
                            const orbx_dmatch *matches, int32_t nm, const uint16_t *prev_depth, int32_t width, int32_t height, size_t dstep,
                            float fx, float fy, float cx, float cy, float *pts3d, float *pts2d, int32_t *n_out);
orbx_status orbx_pnp_score(orbx_handle *h, const float *pts3d, const float *pts2d, int32_t n, const double *Rt, int32_t nh,
                           double fx, double fy, double cx, double cy, double threshold, int32_t *inlier_counts, int32_t *best, uint8_t *best_mask);

/* ---- keyframe packing: the landmark / observation loop of Frontend::publishKeyframe (frontend.cpp:731-776), SURVEY §8(f) rank 2 ----
 * For every keypoint with a depth in (0.3, 3.0) m: back-projection with the float intrinsics, world transform R*p + t in double,
 * one 80-byte record; order preserved; landmark_id = index of the keypoint in the input list, as in the reference.
 * Device variant: per-frame lists laid out like the outputs of orbx_extract_batch_device; asynchronous.                  */
orbx_status orbx_pack_keyframe_device(orbx_handle *h, int32_t nframes, const orbx_keypoint *d_kps, const uint8_t *d_desc,
                                      const int32_t *d_counts, int32_t cap_per_frame, const uint16_t *d_depth, int32_t width, int32_t height,
                                      size_t dstep, size_t dframe_stride, const orbx_kfparams *params,
                                      orbx_kfrecord *d_out, int32_t *d_out_counts, int32_t out_cap_per_frame);
orbx_status orbx_pack_keyframe(orbx_handle *h, const orbx_keypoint *kps, const uint8_t *desc, int32_t n, const uint16_t *depth,
                               int32_t width, int32_t height, size_t dstep, const orbx_kfparams *params,
                               orbx_kfrecord *out, int32_t cap, int32_t *n_out);

/* ---- stage access for parity tests (the reference exposes mvImagePyramid publicly, ORBextractor.hpp:84) ----
 * Valid after an extract call, for frame slot `frame` of the last batch.  Host outputs.          */
/* ORBX_OPT_FAST_DENSE runs: the iniThFAST score map of a level of frame slot `frame` of the last batch — cv::FAST's score buffer (S - 1 where
 * the pixel is a corner at iniThFAST, else 0) over the level's detection range: (h - 38) rows x (w - 38) columns, level pixel (19 + c, 19 + r)
 * at out[r][c].  Columns / rows past the cell grid read 0 (ORBextractor.cpp:811-816 skips those cells).                              */
orbx_status orbx_get_fast_scores(orbx_handle *h, int32_t frame, int32_t level, uint8_t *out, size_t out_step);
/* ORBX_OPT_FAST_DENSE runs: the corners the tile kernel left to the cross-tile NMS kernel (tile-edge corners, and every corner of a tile whose
 * pre-test survivors exceeded its queue), as (x, y, score) like orbx_get_candidates.  Test / diagnosis access.                      */
orbx_status orbx_get_fast_edge_corners(orbx_handle *h, int32_t frame, int32_t level, int32_t *out_xys, int32_t cap, int32_t *n_out);
orbx_status orbx_get_pyramid_level(orbx_handle *h, int32_t frame, int32_t level, uint8_t *out, size_t out_step);
orbx_status orbx_get_blurred_level(orbx_handle *h, int32_t frame, int32_t level, uint8_t *out, size_t out_step);
/* Harris corner response (cv::ORB's HarrisResponses, block 7, k 0.04: the HARRIS_SCORE the reference's ORBextractor.hpp:48 names but
 * never computes) of n points, given as (x, y) int32 pairs in the coordinates of pyramid level `level` of frame slot `frame` of the
 * last batch.  Points closer than block/2 + 1 to the level's edge score 0.                                                    */
orbx_status orbx_harris_responses(orbx_handle *h, int32_t frame, int32_t level, const int32_t *xy, int32_t n, int32_t block_size, float k, float *out);
/* FAST candidates of a level before distribution: packed (x, y, score) relative to the border box,
 * unordered.  out_xys: int32 triples.                                                             */
orbx_status orbx_get_candidates(orbx_handle *h, int32_t frame, int32_t level, int32_t *out_xys, int32_t cap, int32_t *n_out);
/* keypoints per level retained by the quadtree, in the reference's list order */
orbx_status orbx_get_level_counts(orbx_handle *h, int32_t frame, int32_t *out /*nlevels*/);

/* ---- seeded synthetic inputs, generated on the device (bench / tests; identical bytes to the oracle's generator) ---- */
orbx_status orbx_synth_gray_device(orbx_handle *h, uint32_t seed, int32_t first_frame, int32_t nframes,
                                   int32_t width, int32_t height, uint8_t *d_out, size_t step, size_t frame_stride);
orbx_status orbx_synth_depth_device(orbx_handle *h, uint32_t seed, int32_t first_frame, int32_t nframes,
                                    int32_t width, int32_t height, uint16_t *d_out, size_t step_bytes, size_t frame_stride_bytes);
orbx_status orbx_synth_descriptors_device(orbx_handle *h, uint32_t seed, uint64_t first_row, int64_t nrows, uint8_t *d_out);

/* ---- options ----
 * ORBX_OPT_SERIAL: 1 = launch every kernel of a step on the handle's one stream, in order (per-kernel event timings are
 * then isolated, as the roofline accounting wants); 0 (default) = the blur runs on a second stream beside FAST + quadtree.
 * ORBX_OPT_FAST_CTAS: resident FAST warps per SM in the overlapped schedule (0 = default); fewer leave room for the blur
 * running beside FAST.
 * ORBX_OPT_FUSED_BLUR: 1 (default) = the 7x7 Gaussian is evaluated inside the descriptor kernel, only at the pixels the descriptors
 * read (same bits); no blurred pyramid is written and orbx_get_blurred_level computes the level on demand.  0 = blur every level
 * with its own kernel first, as the reference does.
 * ORBX_OPT_PDL: 1 (default) = programmatic stream serialization (the next kernel's CTAs become resident while the previous one drains)
 * for the pyramid's chain of launches and, for batches of up to 8 frames (the latency path), for every kernel of the step;
 * 0 = plain stream order. */
#define ORBX_OPT_SERIAL 1
#define ORBX_OPT_FAST_CTAS 2
#define ORBX_OPT_FUSED_BLUR 3
#define ORBX_OPT_PDL 4
/* ORBX_OPT_OVERLAP: 1 = device-buffer batch calls of >= 32 frames are cut into two half-batches that run on two streams, the second one
 * stage behind the first (its pyramid beside the first half's FAST, its FAST beside the first half's quadtree / descriptor / match
 * kernels ...), joined again on the handle's stream before the call returns; same results, same stream semantics.  0 (default) = one
 * chain.  MEASURED SLOWER on B200 (128 x 1280x720: 1.190 ms one chain, 1.206-1.31 ms overlapped, profiles/r02d_overlap.log): every
 * kernel of the step is instruction-issue bound, so co-resident kernels only split the issue slots and the half-batches add tails. */
#define ORBX_OPT_OVERLAP 5
/* ORBX_OPT_MATCH_MMA: the engine of the brute-force matcher, the landmark-database top-2 query and the association's descriptor gate.  All of
 * them compute Hamming distances as an exact integer GEMM on descriptors unpacked to 0/1 bytes (popc(q ^ t) = popc(q) + popc(t) - 2 q.t) or with
 * LOP3 / POPC; same results bit for bit.
 *   1 (default) = the tensor-memory kernels (tcgen05.mma kind::i8, 128 x 128 x 256 tiles, accumulators in TMEM: k_match_umma, k_assoc_umma) when a
 *                 call holds at least 8 M descriptor pairs; a single frame pair stays on the POPC kernel (less fixed latency)
 *   3 = the tensor-memory kernels for every call        2 = the mma.sync m16n8k32.u8 kernels (k_match_mma, k_assoc_mma)        0 = LOP3 / POPC */
#define ORBX_OPT_MATCH_MMA 6
/* ORBX_OPT_FAST_DENSE: the second FAST formulation (k_fast_dense.cu): whole-level tiles of 128 x 16 pixels scored at iniThFAST into a score map
 * and a corner list, a per-corner NMS kernel restricted to the corner's cell, and a retry launch of the warp-per-cell kernel for the cells
 * iniThFAST left empty.  0 (default) = the warp-per-cell kernel for every call; 1 = dense for batches of >= 8 frames; 2 = dense for every
 * call; 3 = dense for every call with the NMS as work items inside the tile kernel (an experiment, slower).  Same keypoint sets (tests/test_gpu_fast_dense.py).  Off by default: measured 0.50 ms against 0.45 ms per 128 frames of 1280 x 720
 * — its NMS kernel re-reads the 420 MB of score maps from HBM.  Switching it on allocates its arenas
 * (0.8 GB at 128 frames of 1280 x 720): ORBX_E_CUDA if they cannot be had.                                                          */
#define ORBX_OPT_FAST_DENSE 7
/* ORBX_OPT_FILTER_FIRST: order of the depth / box filter and the descriptor kernel in the filtered calls.  The filters of the reference
 * (frontend.cpp:503-527 after operator(), backend.cpp:1011-1029) look at a keypoint's position only, which is final once the quadtree has
 * selected it.  1 (default) = the filter runs on the selected positions first and leaves an ordered index list (k_keep_list); the descriptor
 * kernel computes angles and descriptors of the survivors only, straight into their final rows.  0 = the reference's order: describe every
 * selected keypoint, then drop rows (k_filter).  Same output bit for bit (tests/test_gpu_baseline_configs.py runs both); on the RGB-D
 * stream a fifth of the selected keypoints fail the depth test. */
#define ORBX_OPT_FILTER_FIRST 8
orbx_status orbx_set_option(orbx_handle *h, int32_t option, int32_t value);

/* ---- utilities ---- */
void *orbx_alloc_pinned(size_t bytes);
void  orbx_free_pinned(void *p);
void *orbx_alloc_device(orbx_handle *h, size_t bytes);
void  orbx_free_device(orbx_handle *h, void *p);
orbx_status orbx_copy_to_device(orbx_handle *h, void *d_dst, const void *src, size_t bytes);
orbx_status orbx_copy_to_host(orbx_handle *h, void *dst, const void *d_src, size_t bytes);
/* device self-tests used by the parity suite: evaluate the device cosf/sinf/fastAtan2 restatements */
orbx_status orbx_test_trig(orbx_handle *h, const float *in, int32_t n, float *out_cos, float *out_sin);
orbx_status orbx_test_atan2(orbx_handle *h, const float *y, const float *x, int32_t n, float *out);
/* checksum of device cosf/sinf over every float32 angle in [0,360] degrees (exhaustive pin vs glibc) */
orbx_status orbx_test_trig_checksum(orbx_handle *h, uint32_t first_bits, uint32_t last_bits, uint64_t *sum_cos, uint64_t *sum_sin);
orbx_status orbx_test_quadtree(orbx_handle *h, const int32_t *xys, int32_t n, int32_t box_w, int32_t box_h,
                               int32_t wcell, int32_t hcell, int32_t ncols, int32_t N, int32_t *out_xys, int32_t cap, int32_t *n_out);
/* POPC issue-rate microbenchmark (matching roofline denominator): returns popc/s over the whole GPU */
orbx_status orbx_bench_popc(orbx_handle *h, double *popc_per_sec);
/* number of kernels this library has launched on the handle since creation */
int64_t orbx_launch_count(const orbx_handle *h);
/* per-kernel device time, CUDA events recorded on the handle's stream around every launch while enabled.
 * orbx_profile_read synchronises, accumulates and returns total ms and launch count per kernel id
 * (0 <= id < orbx_profile_kernels(); names from orbx_profile_name), then clears the accumulators.   */
void        orbx_profile_enable(orbx_handle *h, int32_t on);
int32_t     orbx_profile_kernels(void);
const char *orbx_profile_name(int32_t id);
orbx_status orbx_profile_read(orbx_handle *h, double *ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H */
