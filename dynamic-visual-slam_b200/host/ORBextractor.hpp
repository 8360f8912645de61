// ORBextractor.hpp — header-only C++ adapter over the C ABI (include/orbx.h).
//
// Reproduces the two call shapes the reference's nodes use today, so that switching the frontend / backend
// to the B200 path is a change of #include and namespace, not of call sites:
//
//   seam 1  ORB_SLAM3::ORBextractor                      reference include/dynamic_visual_slam/ORBextractor.hpp:44-111
//           ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)                          :50-51
//           int operator()(InputArray image, InputArray mask, vector<KeyPoint>&, OutputArray desc,
//                          vector<int>& vLappingArea)                                                    :58-60
//           GetLevels / GetScaleFactor / GetScaleFactors / GetInverseScaleFactors /
//           GetScaleSigmaSquares / GetInverseScaleSigmaSquares                                           :62-82
//           public mvImagePyramid                                                                        :84
//   seam 2  cv::BFMatcher(cv::NORM_HAMMING).match(query, train, matches)      reference frontend.cpp:220, :614, :1123;
//                                                                             backend.cpp:222, :1072
//   plus    Frontend::filterDepth (frontend.cpp:503-527) and the `distance < 50` loop (:1126-1132) as optional
//           fused post-filters (extractFiltered / matchBelow), results identical to running them on the host.
//
// Error behaviour mirrors the reference: operator() returns -1 for an empty image (ORBextractor.cpp:1090-1091),
// a non-CV_8UC1 image is a programming error (the reference asserts, :1094), everything else throws
// (cv::Exception when OpenCV is present, std::runtime_error otherwise) and is caught by the nodes' existing
// try/catch (frontend.cpp:1319-1323).  Like the reference's extractor an instance is NOT re-entrant.
//
// Build modes:
//   -DORBX_WITH_OPENCV   cv::Mat / cv::KeyPoint / cv::DMatch are used directly (what the ROS nodes compile with)
//   (default)            layout-identical stand-ins in orbx::compat, so the adapter builds and is tested where
//                        OpenCV headers are absent (this repository's containers).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/orbx.h"

#ifdef ORBX_WITH_OPENCV
#include <opencv2/core.hpp>
#include <opencv2/features2d.hpp>
#endif

namespace orbx {

#ifdef ORBX_WITH_OPENCV
using Mat = cv::Mat;
using KeyPoint = cv::KeyPoint;
using DMatch = cv::DMatch;
static inline int type_8uc1() { return CV_8UC1; }
static inline int type_16uc1() { return CV_16UC1; }
[[noreturn]] static inline void raise(const std::string &what) { CV_Error(cv::Error::StsError, what); }
#else
namespace compat {
// cv::KeyPoint is { Point2f pt; float size, angle, response; int octave, class_id; } = 28 bytes
struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
// cv::DMatch is { int queryIdx, trainIdx, imgIdx; float distance; } = 16 bytes
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };
// the subset of cv::Mat the path touches: a dense 2-D array with a row step, optionally owning its bytes
struct Mat {
    int rows = 0, cols = 0, type_ = 0;       // type_: 0 = CV_8UC1, 2 = CV_16UC1 (OpenCV's numeric values)
    size_t step = 0;
    uint8_t *data = nullptr;
    std::vector<uint8_t> own;
    Mat() {}
    Mat(int r, int c, int t, void *d, size_t s) : rows(r), cols(c), type_(t), step(s), data((uint8_t *)d) {}
    void create(int r, int c, int t) {
        rows = r; cols = c; type_ = t; step = (size_t)c * (t == 2 ? 2 : 1);
        own.assign(step * (size_t)(r > 0 ? r : 0), 0); data = own.data();
    }
    bool empty() const { return rows <= 0 || cols <= 0 || !data; }
    int type() const { return type_; }
    template <class T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <class T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
};
}  // namespace compat
using Mat = compat::Mat;
using KeyPoint = compat::KeyPoint;
using DMatch = compat::DMatch;
static inline int type_8uc1() { return 0; }
static inline int type_16uc1() { return 2; }
[[noreturn]] static inline void raise(const std::string &what) { throw std::runtime_error(what); }
#endif

static_assert(sizeof(KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint layout");
static_assert(sizeof(DMatch) == sizeof(orbx_dmatch), "cv::DMatch layout");

// Drop-in for ORB_SLAM3::ORBextractor.
class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    // max_width / max_height size the device arenas once (README default stream: 1280x720); device = CUDA ordinal.
    // profile: ORBX_PROFILE_SLAM = this class in the reference (the frontend's extractor); ORBX_PROFILE_CVORB = cv::ORB::create(nfeatures,
    // scaleFactor, nlevels, 31, 0, 2, HARRIS_SCORE, 31, iniThFAST), the extractor of the reference's gtest (test/test_dbow2_integration.cpp:19)
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                 int max_width = 1280, int max_height = 720, int device = 0, int profile = ORBX_PROFILE_SLAM)
    {
        orbx_params p;
        orbx_default_params(&p);
        p.profile = profile;
        p.nfeatures = nfeatures; p.scale_factor = scaleFactor; p.nlevels = nlevels;
        p.ini_th_fast = iniThFAST; p.min_th_fast = minThFAST;
        p.max_width = max_width; p.max_height = max_height; p.device = device;
        const orbx_status st = orbx_create(&p, &h_);
        if (st != ORBX_OK) raise(std::string("orbx_create: ") + orbx_last_error(nullptr));
        nlevels_ = nlevels;
    }
    ~ORBextractor() { orbx_destroy(h_); }
    ORBextractor(const ORBextractor &) = delete;
    ORBextractor &operator=(const ORBextractor &) = delete;

    // Mask is ignored, as in the reference (ORBextractor.hpp:55-57).  vLappingArea reproduces the reference's mono / stereo split
    // (ORBextractor.cpp:1139-1166): keypoints with vLappingArea[0] <= pt.x <= vLappingArea[1] (scaled coordinates) are written from the
    // BACK of the arrays in the order they are met, all others from the front; the return value is the number at the front (monoIndex).
    // The frontend passes {0, 0} (frontend.cpp:290), for which everything with pt.x != 0 is "mono" and the result is the plain list.
#ifdef ORBX_WITH_OPENCV
    int operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<KeyPoint> &keypoints, cv::OutputArray _descriptors,
                   std::vector<int> &vLappingArea)
    {
        if (_image.empty()) return -1;                                             // ORBextractor.cpp:1090-1091
        Mat image = _image.getMat(), desc;
        const int n = extractFiltered(image, Mat(), keypoints, desc);
        const int mono = applyLappingArea(keypoints, desc, vLappingArea);
        if (n == 0) _descriptors.release();                                        // :1108
        else desc.copyTo(_descriptors);                                            // _descriptors.create(n, 32, CV_8U), :1112
        return mono;
    }
#else
    int operator()(const Mat &image, const Mat & /*mask*/, std::vector<KeyPoint> &keypoints, Mat &descriptors,
                   std::vector<int> &vLappingArea)
    {
        const int n = extractFiltered(image, Mat(), keypoints, descriptors);
        return n < 0 ? n : applyLappingArea(keypoints, descriptors, vLappingArea);
    }
#endif

    // operator() followed by Frontend::filterDepth (frontend.cpp:503-527) when `depth` (CV_16UC1, millimetres) is given, and by
    // Backend::categorizeObservation + the filtered_objects_ test (backend.cpp:1011-1029, 746-751) when YOLO boxes are given: a keypoint
    // whose FIRST containing box has a class with its bit set in drop_class_mask is dropped.  Both run after selection, as in the reference.
    int extractFiltered(const Mat &image, const Mat &depth, std::vector<KeyPoint> &keypoints, Mat &descriptors,
                        const std::vector<orbx_box> &boxes = std::vector<orbx_box>(), uint64_t drop_class_mask = 0)
    {
        if (image.empty()) return -1;                                              // ORBextractor.cpp:1090-1091
        if (image.type() != type_8uc1()) raise("ORBextractor: image must be CV_8UC1");   // reference: assert, :1094
        if (!depth.empty() && (depth.type() != type_16uc1() || depth.rows != image.rows || depth.cols != image.cols))
            raise("ORBextractor: depth must be CV_16UC1 of the image size");
        kbuf_.resize(cap_);
        dbuf_.resize((size_t)cap_ * ORBX_DESC_BYTES);
        int32_t n = 0;
        orbx_status st = ORBX_OK;
        for (int attempt = 0; attempt < 2; attempt++) {                            // grow once and retry on ORBX_E_CAPACITY
            st = orbx_extract_filtered(h_, image.data, image.cols, image.rows, (size_t)image.step,
                                       depth.empty() ? nullptr : (const uint16_t *)depth.data, depth.empty() ? 0 : (size_t)depth.step,
                                       boxes.empty() ? nullptr : boxes.data(), (int32_t)boxes.size(), drop_class_mask, kbuf_.data(), dbuf_.data(), cap_, &n);
            if (st != ORBX_E_CAPACITY || n <= cap_) break;
            cap_ = n + 64; kbuf_.resize(cap_); dbuf_.resize((size_t)cap_ * ORBX_DESC_BYTES);
        }
        if (st != ORBX_OK) raise(std::string("orbx_extract: ") + orbx_last_error(h_));
        keypoints.resize((size_t)n);
        if (n > 0) std::memcpy((void *)keypoints.data(), kbuf_.data(), (size_t)n * sizeof(orbx_keypoint));
        if (n == 0) descriptors = Mat();                                           // _descriptors.release(), :1108
        else {
            descriptors.create(n, 32, type_8uc1());                                // :1112
            for (int r = 0; r < n; r++) std::memcpy(descriptors.ptr<uint8_t>(r), dbuf_.data() + (size_t)r * ORBX_DESC_BYTES, ORBX_DESC_BYTES);
        }
        last_w_ = image.cols; last_h_ = image.rows;
        pyramid_valid_ = false;
        return n;                                                                  // monoIndex for vLappingArea = {0, 0}, :1166
    }

    // The "FEATURE CULLING FOR BACKEND" block of Frontend::syncCallback (frontend.cpp:1168-1218): backend set = the query keypoint of
    // every geometrically consistent match, then the best unmatched features by response (at most max_new, response >= min_response),
    // equal responses in the order the reference's std::sort leaves them.
    void cullForBackend(const std::vector<KeyPoint> &filtered_keypoints, const Mat &filtered_descriptors, const std::vector<DMatch> &matches,
                        std::vector<KeyPoint> &backend_keypoints, Mat &backend_descriptors, int max_new = 200, float min_response = 50.0f)
    {
        const int n = (int)filtered_keypoints.size();
        if (n > 0 && (filtered_descriptors.rows != n || filtered_descriptors.cols != 32 || filtered_descriptors.type() != type_8uc1()))
            raise("cullForBackend: descriptors must be n x 32 CV_8UC1");
        std::vector<int32_t> q(matches.size());
        for (size_t i = 0; i < matches.size(); i++) q[i] = matches[i].queryIdx;
        std::vector<uint8_t> din((size_t)std::max(n, 1) * ORBX_DESC_BYTES);
        for (int r = 0; r < n; r++) std::memcpy(din.data() + (size_t)r * ORBX_DESC_BYTES, filtered_descriptors.ptr<uint8_t>(r), ORBX_DESC_BYTES);
        const int cap = (int)q.size() + std::min(max_new, n);
        std::vector<orbx_keypoint> ok((size_t)std::max(cap, 1));
        std::vector<uint8_t> od((size_t)std::max(cap, 1) * ORBX_DESC_BYTES);
        int32_t m = 0;
        const orbx_status st = orbx_cull_keyframe(h_, (const orbx_keypoint *)filtered_keypoints.data(), din.data(), n, q.data(), (int32_t)q.size(),
                                                  max_new, min_response, ok.data(), od.data(), nullptr, cap, &m);
        if (st != ORBX_OK) raise(std::string("orbx_cull_keyframe: ") + orbx_last_error(h_));
        backend_keypoints.resize((size_t)m);
        if (m > 0) std::memcpy((void *)backend_keypoints.data(), ok.data(), (size_t)m * sizeof(orbx_keypoint));
        if (m == 0) backend_descriptors = Mat();
        else {
            backend_descriptors.create(m, 32, type_8uc1());
            for (int r = 0; r < m; r++) std::memcpy(backend_descriptors.ptr<uint8_t>(r), od.data() + (size_t)r * ORBX_DESC_BYTES, ORBX_DESC_BYTES);
        }
    }

    int GetLevels() { return orbx_get_levels(h_); }
    float GetScaleFactor() { return orbx_get_scale_factor(h_); }
    std::vector<float> GetScaleFactors() { return vec(orbx_get_scale_factors); }
    std::vector<float> GetInverseScaleFactors() { return vec(orbx_get_inverse_scale_factors); }
    std::vector<float> GetScaleSigmaSquares() { return vec(orbx_get_scale_sigma_squares); }
    std::vector<float> GetInverseScaleSigmaSquares() { return vec(orbx_get_inverse_scale_sigma_squares); }

    // The reference exposes its pyramid as a public member (ORBextractor.hpp:84).  Here the levels live in HBM and
    // are fetched on demand (nothing in the two nodes reads them; they exist for parity checks).
    const std::vector<Mat> &imagePyramid()
    {
        if (!pyramid_valid_ && last_w_ > 0) {
            mvImagePyramid.resize((size_t)nlevels_);
            for (int l = 0; l < nlevels_; l++) {
                int32_t lw = 0, lh = 0;
                orbx_level_size(h_, last_w_, last_h_, l, &lw, &lh);
                mvImagePyramid[l].create(lh, lw, type_8uc1());
                if (orbx_get_pyramid_level(h_, 0, l, mvImagePyramid[l].data, (size_t)mvImagePyramid[l].step) != ORBX_OK)
                    raise(std::string("orbx_get_pyramid_level: ") + orbx_last_error(h_));
            }
            pyramid_valid_ = true;
        }
        return mvImagePyramid;
    }
    std::vector<Mat> mvImagePyramid;

    orbx_handle *handle() { return h_; }

private:
    // the reference's two-ended fill (ORBextractor.cpp:1152-1161), applied to the finished list; returns monoIndex
    int applyLappingArea(std::vector<KeyPoint> &keypoints, Mat &descriptors, const std::vector<int> &lap)
    {
        const int n = (int)keypoints.size();
        if (lap.size() < 2) raise("ORBextractor: vLappingArea needs two entries");     // the reference indexes [0] and [1] unchecked
        std::vector<int> src((size_t)n);
        int mono = 0, stereo = n - 1;
        for (int i = 0; i < n; i++) {
            const float x = keypoints[(size_t)i].pt.x;
            if (x >= lap[0] && x <= lap[1]) src[(size_t)stereo--] = i; else src[(size_t)mono++] = i;
        }
        if (mono == n) return mono;                                                // nothing in the lapping area: order unchanged
        std::vector<KeyPoint> k((size_t)n);
        std::vector<uint8_t> d((size_t)n * ORBX_DESC_BYTES);
        for (int j = 0; j < n; j++) {
            k[(size_t)j] = keypoints[(size_t)src[(size_t)j]];
            std::memcpy(d.data() + (size_t)j * ORBX_DESC_BYTES, descriptors.ptr<uint8_t>(src[(size_t)j]), ORBX_DESC_BYTES);
        }
        keypoints.swap(k);
        for (int j = 0; j < n; j++) std::memcpy(descriptors.ptr<uint8_t>(j), d.data() + (size_t)j * ORBX_DESC_BYTES, ORBX_DESC_BYTES);
        return mono;
    }
    std::vector<float> vec(void (*fn)(const orbx_handle *, float *))
    {
        std::vector<float> v((size_t)nlevels_);
        fn(h_, v.data());
        return v;
    }
    orbx_handle *h_ = nullptr;
    int nlevels_ = 0, cap_ = 2048, last_w_ = 0, last_h_ = 0;
    bool pyramid_valid_ = false;
    std::vector<orbx_keypoint> kbuf_;
    std::vector<uint8_t> dbuf_;
};

// Drop-in for the `cv::BFMatcher matcher_(cv::NORM_HAMMING)` members (frontend.cpp:294, backend.cpp:628).
class BFMatcher {
public:
    explicit BFMatcher(ORBextractor &ex) : h_(ex.handle()) {}
    explicit BFMatcher(orbx_handle *h) : h_(h) {}

    // cv::DescriptorMatcher::match: one DMatch per query row, lowest trainIdx on ties
    void match(const Mat &query, const Mat &train, std::vector<DMatch> &matches) { run(query, train, 1, 0.f, 0.f, matches); }
    // match() followed by the nodes' `if (m.distance < max_dist) good.push_back(m)` loop (frontend.cpp:1126-1132)
    void matchBelow(const Mat &query, const Mat &train, float max_dist, std::vector<DMatch> &good) { run(query, train, 1, max_dist, 0.f, good); }
    // cv::DescriptorMatcher::knnMatch(k = 2)
    void knnMatch(const Mat &query, const Mat &train, std::vector<std::vector<DMatch>> &matches, int k)
    {
        if (k != 2) raise("BFMatcher::knnMatch: only k = 2 is implemented");
        std::vector<DMatch> flat;
        run(query, train, 2, 0.f, 0.f, flat);
        matches.assign(flat.size() / 2, std::vector<DMatch>());
        for (size_t i = 0; i < matches.size(); i++)
            for (int j = 0; j < 2; j++) if (flat[2 * i + j].trainIdx >= 0) matches[i].push_back(flat[2 * i + j]);
    }

private:
    void run(const Mat &query, const Mat &train, int k, float max_dist, float ratio, std::vector<DMatch> &out)
    {
        out.clear();
        if (query.empty() || train.empty()) return;                 // callers pre-guard this case (frontend.cpp:1107-1117)
        if (query.cols != 32 || train.cols != 32 || query.type() != type_8uc1() || train.type() != type_8uc1())
            raise("BFMatcher: descriptors must be N x 32 CV_8U");
        std::vector<uint8_t> q = pack(query), t = pack(train);
        std::vector<orbx_dmatch> buf((size_t)query.rows * k);
        int32_t n = 0;
        if (orbx_match(h_, q.data(), query.rows, t.data(), train.rows, k, max_dist, ratio, buf.data(), &n) != ORBX_OK)
            raise(std::string("orbx_match: ") + orbx_last_error(h_));
        out.resize((size_t)n);
        if (n > 0) std::memcpy((void *)out.data(), buf.data(), (size_t)n * sizeof(orbx_dmatch));
    }
    static std::vector<uint8_t> pack(const Mat &m)
    {
        std::vector<uint8_t> v((size_t)m.rows * 32);
        for (int r = 0; r < m.rows; r++) std::memcpy(v.data() + (size_t)r * 32, m.data + (size_t)r * m.step, 32);
        return v;
    }
    orbx_handle *h_;
};

// Backend side: the landmark loop of Backend::associateObservation (backend.cpp:1064-1120) against descriptors resident in HBM.
// One instance per category map (the reference keeps category -> id -> LandmarkInfo, backend.cpp:306-323); rows are appended in the
// order landmarks are created, `first_index` is the global index of row 0 when the map is sharded over several GPUs.
class LandmarkDB {
public:
    LandmarkDB(ORBextractor &ex, int64_t capacity_rows, uint32_t first_index = 0) : h_(ex.handle())
    {
        if (orbx_db_create(h_, capacity_rows, first_index, &db_) != ORBX_OK) raise(std::string("orbx_db_create: ") + orbx_last_error(h_));
    }
    ~LandmarkDB() { orbx_db_destroy(db_); }
    LandmarkDB(const LandmarkDB &) = delete;
    LandmarkDB &operator=(const LandmarkDB &) = delete;

    // append landmark descriptors (N x 32 CV_8U) with their positions (the reference's cv::Point3f, xyz floats)
    void append(const Mat &descriptors, const float *xyz)
    {
        if (descriptors.empty()) return;
        if (descriptors.cols != 32 || descriptors.type() != type_8uc1()) raise("LandmarkDB: descriptors must be N x 32 CV_8U");
        std::vector<uint8_t> rows((size_t)descriptors.rows * 32);
        for (int r = 0; r < descriptors.rows; r++) std::memcpy(rows.data() + (size_t)r * 32, descriptors.ptr<uint8_t>(r), 32);
        const int64_t first = orbx_db_rows(db_);
        if (orbx_db_append(db_, rows.data(), descriptors.rows) != ORBX_OK) raise(std::string("orbx_db_append: ") + orbx_last_error(h_));
        if (xyz && orbx_db_set_positions(db_, first, descriptors.rows, xyz) != ORBX_OK) raise(std::string("orbx_db_set_positions: ") + orbx_last_error(h_));
    }
    void setPositions(int64_t first_row, int64_t nrows, const float *xyz)
    {
        if (orbx_db_set_positions(db_, first_row, nrows, xyz) != ORBX_OK) raise(std::string("orbx_db_set_positions: ") + orbx_last_error(h_));
    }
    int64_t rows() const { return orbx_db_rows(db_); }

    // associateObservation for a batch of observations of this category: per observation the landmark (global row, -1 = none) with the
    // smallest reprojection error among the rows at Hamming distance < max_desc_dist, if that error is < max_reproj_err
    // (reference literals 50.0 and 5.0, backend.cpp:225-226).  pixels: nq x 2 floats (obs.pixel).
    void associate(const Mat &descriptors, const float *pixels, const orbx_pose &pose, std::vector<orbx_assoc> &out,
                   float max_desc_dist = 50.0f, double max_reproj_err = 5.0)
    {
        out.assign((size_t)(descriptors.empty() ? 0 : descriptors.rows), orbx_assoc());
        if (descriptors.empty()) return;
        if (descriptors.cols != 32 || descriptors.type() != type_8uc1()) raise("LandmarkDB: descriptors must be N x 32 CV_8U");
        std::vector<uint8_t> q((size_t)descriptors.rows * 32);
        for (int r = 0; r < descriptors.rows; r++) std::memcpy(q.data() + (size_t)r * 32, descriptors.ptr<uint8_t>(r), 32);
        if (orbx_db_associate(db_, q.data(), pixels, descriptors.rows, &pose, max_desc_dist, max_reproj_err, out.data()) != ORBX_OK)
            raise(std::string("orbx_db_associate: ") + orbx_last_error(h_));
    }
    // descriptor stage alone: the two nearest rows per query (BFMatcher tie-break), e.g. for an all-gather + orbx_merge_top2 over shards
    void queryTop2(const Mat &descriptors, std::vector<orbx_top2> &out)
    {
        out.assign((size_t)(descriptors.empty() ? 0 : descriptors.rows), orbx_top2());
        if (descriptors.empty()) return;
        std::vector<uint8_t> q((size_t)descriptors.rows * 32);
        for (int r = 0; r < descriptors.rows; r++) std::memcpy(q.data() + (size_t)r * 32, descriptors.ptr<uint8_t>(r), 32);
        if (orbx_db_query_top2(db_, q.data(), descriptors.rows, out.data()) != ORBX_OK) raise(std::string("orbx_db_query_top2: ") + orbx_last_error(h_));
    }
    orbx_db *db() { return db_; }

private:
    orbx_handle *h_;
    orbx_db *db_ = nullptr;
};

}  // namespace orbx
