/* orb_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Dependency-free C restatement of the reference's ORB hot path:
 *   ORB_SLAM3::ORBextractor            reference dynamic_visual_slam/src/ORBextractor.cpp
 *   cv::BFMatcher(NORM_HAMMING).match  reference frontend.cpp:1123, :614 ; backend.cpp:1072
 *   Frontend::filterDepth              reference frontend.cpp:457-527
 *   Backend::categorizeObservation     reference backend.cpp:1011-1029
 * plus the OpenCV 4.x primitives those call (resize INTER_LINEAR, FAST-9/16, GaussianBlur 7x7,
 * fastAtan2, cvRound), restated from their published arithmetic and pinned
 * bit-for-bit against python cv2 4.13.0 in tests/test_oracle_vs_cv2.py and tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this library.  The product (liborbx.so) never does.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 16

/* 28-byte cv::KeyPoint layout */
typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } orc_keypoint;
/* 16-byte cv::DMatch layout */
typedef struct { int32_t queryIdx, trainIdx, imgIdx; float distance; } orc_dmatch;
/* yolo_msgs bbox: centre + size, fp64 (reference backend.cpp:1017-1020) */
typedef struct { double cx, cy, w, h; int32_t class_id; int32_t pad; } orc_box;

/* candidate corner inside one level, coordinates relative to the border box (minBorderX,minBorderY) */
typedef struct { int32_t x, y; int32_t score; } orc_cand;

typedef struct {
    int nfeatures, nlevels, iniThFAST, minThFAST;
    double scaleFactor;                       /* double member initialised from a float argument */
    float scale[ORC_MAX_LEVELS], inv_scale[ORC_MAX_LEVELS];
    float sigma2[ORC_MAX_LEVELS], inv_sigma2[ORC_MAX_LEVELS];
    int   nfeat_level[ORC_MAX_LEVELS];
    int   umax[16];
} orc_extractor;

/* optional stage trace for stage-wise parity tests; every pointer may be NULL */
typedef struct {
    uint8_t  *pyramid;        /* levels concatenated, each tightly packed w*h */
    uint8_t  *blurred;        /* same layout */
    orc_cand *cands;          /* per level, capacity cand_cap each: cands + level*cand_cap */
    int32_t   cand_cap;
    int32_t   ncands[ORC_MAX_LEVELS];
    int32_t   nkeys[ORC_MAX_LEVELS];   /* keypoints retained per level */
    int32_t   lw[ORC_MAX_LEVELS], lh[ORC_MAX_LEVELS];
} orc_trace;

int  orc_extractor_init(orc_extractor *ex, int nfeatures, float scaleFactor, int nlevels,
                        int iniThFAST, int minThFAST);
void orc_level_size(const orc_extractor *ex, int w, int h, int level, int *lw, int *lh);
/* 0 = the reference's arithmetic is defined for this frame size; 1/2/3 = it throws or faults (see orb_oracle.c) */
int  orc_geometry_status(const orc_extractor *ex, int w, int h);

void orc_resize_linear(const uint8_t *src, int sw, int sh, size_t sstep,
                       uint8_t *dst, int dw, int dh, size_t dstep);
/* tables used by the resize (exported so the CUDA host code can be checked against them) */
void orc_resize_tables(int ssize, int dsize, int32_t *ofs, int16_t *coef /*2 per i*/, int horizontal);

int  orc_fast_roi(const uint8_t *img, size_t step, int w, int h, int threshold,
                  orc_cand *out, int cap);
int  orc_fast_cells(const orc_extractor *ex, const uint8_t *img, size_t step, int w, int h,
                    orc_cand *out, int cap);
int  orc_distribute_octtree(const orc_cand *cands, int n, int minX, int maxX, int minY, int maxY,
                            int N, orc_cand *out, int cap);
float orc_fast_atan2(float y, float x);
float orc_ic_angle(const uint8_t *img, size_t step, int cx, int cy, const int *umax);
void orc_gaussian_blur7(const uint8_t *src, int w, int h, size_t sstep, uint8_t *dst, size_t dstep);
/* inlier mask / count of one fundamental-matrix hypothesis under OpenCV's RANSAC error (frontend.cpp:1134-1154) */
int  orc_fmat_inliers(const float *p1, const float *p2, int n, const double *F, double thresh, uint8_t *mask);
int  orc_pnp_points(const orc_keypoint *prev_kps, const orc_keypoint *curr_kps, const orc_dmatch *m, int nm, const uint16_t *depth, int w, int h,
                    size_t dstep, float fx, float fy, float cx, float cy, float *p3, float *p2);
int  orc_pnp_inliers(const float *p3, const float *p2, int n, const double *R, const double *t, double fx, double fy, double cx, double cy,
                     double thresh, uint8_t *mask);
/* profile C (cv::ORB) primitives: cv::resize INTER_LINEAR_EXACT and the float-path GaussianBlur of a sub-matrix */
void orc_resize_exact_tables(int ssize, int dsize, int32_t *ofs, int16_t *coef /*2 per i*/);
void orc_resize_linear_exact(const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw, int dh, size_t dstep);
void orc_gaussian_blur7_f32(const uint8_t *src, int w, int h, size_t sstep, uint8_t *dst, size_t dstep);
void orc_descriptor(const uint8_t *blur, size_t step, int cx, int cy, float angle_deg, uint8_t *desc32);

/* ORBextractor::operator() ; returns keypoint count, -1 on empty image, -2 capacity, -3 frame size outside the reference's defined domain */
int  orc_extract(const orc_extractor *ex, const uint8_t *gray, int w, int h, size_t step,
                 orc_keypoint *kps, uint8_t *desc, int cap, orc_trace *trace);
/* frame-parallel (OpenMP) batch, used by the CPU baseline: frames tightly packed */
int  orc_extract_batch(const orc_extractor *ex, const uint8_t *gray, int nframes, int w, int h,
                       orc_keypoint *kps, uint8_t *desc, int cap_per_frame, int32_t *counts, int nthreads);

int  orc_filter_depth(const orc_keypoint *kps, const uint8_t *desc, int n,
                      const uint16_t *depth, int dw, int dh, size_t dstep_elems,
                      float min_depth, float max_depth,
                      orc_keypoint *okps, uint8_t *odesc, int32_t *orig_idx);
/* first box containing the pixel → class id, -1 = "unlabeled" */
int  orc_categorize(float px, float py, const orc_box *boxes, int nboxes);

int  orc_hamming(const uint8_t *a, const uint8_t *b);
/* BFMatcher.match: one DMatch per query (k=1), lowest trainIdx on ties */
int  orc_match(const uint8_t *q, int nq, const uint8_t *t, int nt, orc_dmatch *out, int nthreads);
/* knnMatch k=2: out[2*i], out[2*i+1]; trainIdx=-1 when fewer than k train rows */
int  orc_knn2(const uint8_t *q, int nq, const uint8_t *t, int nt, orc_dmatch *out, int nthreads);

/* Backend::reprojectPoint (backend.cpp:1153-1173); R row-major 3x3, (-1,-1) behind the camera */
void orc_reproject(const float *p, const double *R, const double *t, double fx, double fy, double cx, double cy, float *uv);
/* Backend::associateObservation (backend.cpp:1064-1120) over a batch of observations of one category */
void orc_associate(const uint8_t *q, const float *qpx, int nq, const uint8_t *rows, const float *pos, int nrows,
                   const double *R, const double *t, double fx, double fy, double cx, double cy,
                   double max_desc, double max_reproj, int32_t *out_idx, double *out_err, float *out_dist, int nthreads);

/* one Landmark + Observation pair of Keyframe.msg (reference msg/Landmark.msg, msg/Observation.msg), 80 bytes */
typedef struct { uint64_t landmark_id; double position[3]; double pixel_x, pixel_y; uint8_t descriptor[32]; } orc_kfrecord;
/* Frontend::publishKeyframe packing loop (frontend.cpp:731-776); returns the number of records */
int orc_pack_keyframe(const orc_keypoint *kps, const uint8_t *desc, int n, const uint16_t *depth, int dw, int dh, size_t dstep_elems,
                      float fx, float fy, float cx, float cy, const double *R, const double *t, orc_kfrecord *out);

/* Frontend feature culling for the backend (frontend.cpp:1168-1218): match query indices in match order, then the best unmatched
 * keypoints by response (std::sort tie order of libstdc++); returns the number of indices written (<= n_matches + max_new), -1 on a bad index */
int orc_cull_keyframe(const float *response, int n, const int32_t *match_query, int n_matches, int max_new, float min_response, int32_t *out_index);

/* cv::ORB's HarrisResponses for one point (HARRIS_SCORE, ORBextractor.hpp:48) */
float orc_harris_response(const uint8_t *img, size_t step, int x0, int y0, int blockSize, float k);

/* cv::cvtColor(BGR2GRAY), reference frontend.cpp:1084 */
void orc_bgr2gray(const uint8_t *bgr, int w, int h, size_t sstep, uint8_t *gray, size_t dstep);

/* seeded integer-only synthetic inputs (identical bytes on host and device) */
void orc_synth_gray(uint32_t seed, int frame, int w, int h, uint8_t *out, size_t step);
void orc_synth_depth(uint32_t seed, int frame, int w, int h, uint16_t *out, size_t step_elems);
void orc_synth_descriptors(uint32_t seed, uint64_t first_row, int nrows, uint8_t *out);
int  orc_synth_boxes(uint32_t seed, int frame, int w, int h, orc_box *out /*4*/);
int  orc_filter_boxes(const orc_keypoint *kps, const uint8_t *desc, int n, const orc_box *boxes, int nboxes, uint64_t drop_mask,
                      orc_keypoint *okps, uint8_t *odesc);

/* libstdc++ std::sort emulation on (count, ULx) keys, exported for its own test */
void orc_introsort_pairs(int32_t *cnt, int32_t *ulx, int32_t *payload, int n);

float orc_cosf(float x);
float orc_sinf(float x);
/* wrapping sums of the bit patterns of cosf / sinf(angle * factorPI) over the float bit patterns [first, last] of `angle` (libm) */
void orc_trig_checksum(uint32_t first, uint32_t last, uint64_t *sum_cos, uint64_t *sum_sin, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
