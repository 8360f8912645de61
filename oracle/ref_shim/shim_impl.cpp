// shim_impl.cpp — ORACLE build shim (test infrastructure): the five OpenCV primitives that the reference's
// ORBextractor.cpp calls, forwarded to the C primitives of oracle/orb_oracle.c.  Those are restatements of
// OpenCV 4.x's published arithmetic and are pinned bit for bit against python cv2 4.13.0
// in tests/test_oracle_vs_cv2.py; this file adds only argument checking and cv::Mat plumbing.
#include "opencv2/opencv.hpp"

#include <algorithm>

#include "../orb_oracle.h"

namespace cv {

static ShimFastHook g_fast_hook = nullptr;
static void* g_fast_user = nullptr;
void shim_set_fast_hook(ShimFastHook hook, void* user) { g_fast_hook = hook; g_fast_user = user; }

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

void resize(InputArray src_, OutputArray dst_, Size dsize, double fx, double fy, int interpolation)
{
    Mat src = src_.getMat();
    if (src.type() != CV_8UC1) shim_fail("resize: only CV_8UC1");
    if (interpolation != INTER_LINEAR) shim_fail("resize: only INTER_LINEAR (ORBextractor.cpp:1182)");
    if (fx != 0 || fy != 0) shim_fail("resize: explicit dsize only");
    if (src.empty() || dsize.width <= 0 || dsize.height <= 0) throw Exception("resize: (-215:Assertion failed) !ssize.empty() / !dsize.empty()");
    dst_.create(dsize, src.type());                 // keeps the pyramid ROI when the size already matches, as cv::Mat::create does
    Mat dst = dst_.getMat();
    if (dst.data == src.data) shim_fail("resize: in place");
    orc_resize_linear(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

// BORDER_REFLECT_101 with or without BORDER_ISOLATED.  Without ISOLATED OpenCV would read real pixels of the parent
// matrix when src is a view; the reference's non-isolated call (:1189) passes the caller's whole image, for which the
// two are the same — a non-isolated view is refused rather than approximated.
void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType)
{
    Mat src = src_.getMat();
    if (src.type() != CV_8UC1) shim_fail("copyMakeBorder: only CV_8UC1");
    const bool isolated = (borderType & BORDER_ISOLATED) != 0;
    if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) shim_fail("copyMakeBorder: only BORDER_REFLECT_101");
    if (!isolated && src.data != src.datastart) shim_fail("copyMakeBorder: non-isolated border of a view");
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = dst_.getMat();
    const int w = src.cols, h = src.rows;
    // interior first (memmove: the reference passes src = the interior view of dst), then the ring from the interior
    uchar* inner = dst.data + (size_t)top * dst.step + left;
    if (inner != src.data)
        for (int y = 0; y < h; y++) std::memmove(inner + (size_t)y * dst.step, src.data + (size_t)y * src.step, (size_t)w);
    for (int y = 0; y < h; y++) {
        uchar* row = inner + (size_t)y * dst.step;
        for (int x = -left; x < 0; x++) row[x] = row[reflect101(x, w)];
        for (int x = w; x < w + right; x++) row[x] = row[reflect101(x, w)];
    }
    for (int y = -top; y < h + bottom; y++) {
        if (y >= 0 && y < h) continue;
        std::memcpy(dst.data + (size_t)(y + top) * dst.step, dst.data + (size_t)(reflect101(y, h) + top) * dst.step, (size_t)(w + left + right));
    }
}

// cv::GaussianBlur on a continuous, non-view CV_8UC1 with a 7x7 kernel and sigma 2 takes OpenCV's fixed-point path
//  — that is the only shape the reference produces (a .clone(), ORBextractor.cpp:1132-1133).
void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int borderType)
{
    Mat src = src_.getMat();
    if (src.type() != CV_8UC1) shim_fail("GaussianBlur: only CV_8UC1");
    if (ksize.width != 7 || ksize.height != 7 || sigmaX != 2.0 || (sigmaY != 2.0 && sigmaY != 0.0)) shim_fail("GaussianBlur: only 7x7, sigma 2");
    if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) shim_fail("GaussianBlur: only BORDER_REFLECT_101");
    if (src.data != src.datastart || !src.isContinuous()) shim_fail("GaussianBlur: a view takes OpenCV's float path, not emulated");
    Mat tmp(src.rows, src.cols, CV_8UC1);
    orc_gaussian_blur7(src.data, src.cols, src.rows, src.step, tmp.data, tmp.step);
    tmp.copyTo(dst_);
}

// cv::FAST, TYPE_9_16, with non-maximum suppression: KeyPoint(x, y, 7.f, -1, score) in raster order
void FAST(InputArray image_, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression)
{
    Mat img = image_.getMat();
    if (img.type() != CV_8UC1) shim_fail("FAST: only CV_8UC1");
    if (!nonmaxSuppression) shim_fail("FAST: only nonmaxSuppression = true");
    keypoints.clear();
    if (!img.empty()) {
        const int cap = img.rows * img.cols;
        std::vector<orc_cand> c((size_t)cap);
        int n = orc_fast_roi(img.data, img.step, img.cols, img.rows, threshold, c.data(), cap);
        keypoints.reserve((size_t)n);
        for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)c[i].x, (float)c[i].y, 7.f, -1, (float)c[i].score));
    }
    if (g_fast_hook) {
        ShimFastCall call;
        call.datastart = img.datastart;
        size_t off = img.data ? (size_t)(img.data - img.datastart) : 0;
        call.y0 = img.step ? (int)(off / img.step) : 0;
        call.x0 = img.step ? (int)(off % img.step) : 0;
        call.w = img.cols; call.h = img.rows; call.threshold = threshold;
        call.out = keypoints;
        g_fast_hook(call, g_fast_user);
    }
}

// cv::KeyPointsFilter::retainBest as OpenCV documents it: keep the n strongest responses and every further keypoint
// whose response equals the n-th one (used only by the reference's dead ComputeKeyPointsOld, :1049/:1067)
void KeyPointsFilter::retainBest(std::vector<KeyPoint>& keypoints, int npoints)
{
    if (npoints < 0 || keypoints.size() <= (size_t)npoints) return;
    if (npoints == 0) { keypoints.clear(); return; }
    std::nth_element(keypoints.begin(), keypoints.begin() + npoints - 1, keypoints.end(),
                     [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
    const float edge = keypoints[(size_t)npoints - 1].response;
    auto last = std::partition(keypoints.begin() + npoints, keypoints.end(), [edge](const KeyPoint& k) { return k.response >= edge; });
    keypoints.resize((size_t)(last - keypoints.begin()));
}

}  // namespace cv
