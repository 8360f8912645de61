// opencv2/opencv.hpp — ORACLE build shim (test infrastructure, NOT product code, NOT OpenCV).
//
// Purpose: let the reference's OWN translation unit
//     /root/reference/dynamic_visual_slam/src/ORBextractor.cpp  (+ include/dynamic_visual_slam/ORBextractor.hpp)
// compile UNMODIFIED in a container that has no OpenCV C++ headers or libraries, so that its real
// control logic — std::list / std::sort / DivideNode / the per-cell FAST loop / operator() assembly
// (ORBextractor.cpp:409-896, 1077-1194) — runs here and pins the C restatement (oracle/orb_oracle.c)
// and, through it, the CUDA path.  See oracle/Makefile target `_ref` and oracle/ref_glue.cpp.
//
// What is here: exactly the slice of the cv:: API that file touches — Mat (reference-counted buffer,
// ROI views, step), Point_/Size_/Rect_, KeyPoint (28-byte POD layout), Input/OutputArray proxies,
// cvRound/cvFloor/cvCeil, and declarations of the five OpenCV primitives it calls (resize, copyMakeBorder,
// FAST, GaussianBlur, fastAtan2) + KeyPointsFilter::retainBest.  The primitives are implemented in
// ref_shim/shim_impl.cpp by forwarding to the C primitives of orb_oracle.c, which are pinned bit for bit
// against python cv2 4.13.0 (tests/test_oracle_vs_cv2.py).  Nothing else of OpenCV is emulated; an
// unsupported argument (other interpolation, kernel size, type) aborts loudly instead of approximating.
#pragma once
// the real opencv2/core pulls these standard headers in; the reference relies on that (std::sort, std::back_inserter)
#include <algorithm>
#include <iterator>
#include <string>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

// OpenCV's rounding helpers: cvRound is round-half-to-even (lrint under the default rounding mode)
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

[[noreturn]] inline void shim_fail(const char* what)
{
    std::fprintf(stderr, "opencv shim (oracle/ref_shim): unsupported use: %s\n", what);
    std::abort();
}

// cv::Exception: what OpenCV's CV_Assert failures throw (the reference's callers catch std::exception, frontend.cpp:1319-1323)
class Exception : public std::exception {
public:
    explicit Exception(const std::string& m) : msg(m) {}
    const char* what() const noexcept override { return msg.c_str(); }
    std::string msg;
};

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_& operator*=(U s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

// cv::KeyPoint: 28-byte POD, field order as in OpenCV's types.hpp
class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

// cv::DMatch: 16-byte POD (used by the C++ adapter's OpenCV build mode, dynamic-visual-slam_b200/host/ORBextractor.hpp)
struct DMatch {
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    int queryIdx, trainIdx, imgIdx;
    float distance;
};
static_assert(sizeof(DMatch) == 16, "cv::DMatch layout");
namespace Error { enum { StsError = -2, StsBadArg = -5 }; }

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3, INTER_LINEAR_EXACT = 5 };

class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;           // bytes per row (cv::MatStep converts to size_t; the reference only casts it)
    const uchar* datastart; // first byte of the owning buffer (used by the stage trace to name a view's parent)

    Mat() : rows(0), cols(0), data(nullptr), step(0), datastart(nullptr), type_(CV_8UC1) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
    // view over caller-owned memory (no ownership), as cv::Mat(rows, cols, type, data, step)
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : rows(r), cols(c), data((uchar*)ext), datastart((const uchar*)ext), type_(type)
    {
        step = step_ ? step_ : (size_t)c * elemSize();
    }
    Mat(const Mat& m, const Rect& roi) : rows(roi.height), cols(roi.width), data(m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.elemSize()),
                                         step(m.step), datastart(m.datastart), type_(m.type_), buf_(m.buf_)
    {
        if (roi.x < 0 || roi.y < 0 || roi.width < 0 || roi.height < 0 || roi.x + roi.width > m.cols || roi.y + roi.height > m.rows)
            shim_fail("Mat ROI outside the matrix");
    }

    void create(int r, int c, int type)
    {
        if (data && r == rows && c == cols && type == type_) return;       // cv::Mat::create keeps a matching buffer
        type_ = type; rows = r; cols = c;
        step = (size_t)c * elemSize();
        size_t bytes = step * (size_t)r;
        buf_ = std::shared_ptr<uchar>(bytes ? (uchar*)std::malloc(bytes) : nullptr, std::free);
        data = buf_.get(); datastart = data;
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf_.reset(); data = nullptr; datastart = nullptr; rows = cols = 0; step = 0; }

    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize1() const { static const int s[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return (size_t)s[depth()]; }
    size_t elemSize() const { return elemSize1() * (size_t)channels(); }
    size_t step1() const { return step / elemSize1(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }

    Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
    Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
    Mat row(int y) const { return Mat(*this, Rect(0, y, cols, 1)); }

    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int y = 0; y < rows; y++) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
        return m;
    }
    inline void copyTo(OutputArray dst) const;

    static Mat zeros(int r, int c, int type)
    {
        Mat m(r, c, type);
        if (m.data) std::memset(m.data, 0, m.step * (size_t)r);
        return m;
    }

    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }

private:
    int type_;
    std::shared_ptr<uchar> buf_;
};

// proxies: the reference only passes cv::Mat (lvalues, temporaries) and reads them back with getMat()
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    _OutputArray(const Mat& m) : _InputArray(m) {}      // temporaries such as m.row(i): the header is fixed, the pixels are written
    void create(int r, int c, int type) const { if (!m_) shim_fail("create() on noArray()"); m_->create(r, c, type); }
    void create(Size sz, int type) const { create(sz.height, sz.width, type); }
    void release() const { if (m_) m_->release(); }
    Mat& getMatRef() const { if (!m_) shim_fail("getMatRef() on noArray()"); return *m_; }
};
inline InputArray noArray() { static const _OutputArray none; return none; }

inline void Mat::copyTo(OutputArray dst) const
{
    dst.create(rows, cols, type_);
    Mat d = dst.getMat();
    for (int y = 0; y < rows; y++) std::memmove(d.data + (size_t)y * d.step, data + (size_t)y * step, (size_t)cols * elemSize());
}

// ---- the OpenCV primitives the reference calls (implemented in ref_shim/shim_impl.cpp) ------------------
float fastAtan2(float y, float x);                                                     // ORBextractor.cpp:102
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);   // :1182
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);            // :1184, :1189
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);  // :1133
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);               // :826, :845
class KeyPointsFilter {
public:
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);                                               // :1049, :1067 (dead path)
};

// ---- stage trace (shim extension, absent from OpenCV): every FAST call can be observed by the test glue ----
struct ShimFastCall { const uchar* datastart; int x0, y0, w, h, threshold; std::vector<KeyPoint> out; };
typedef void (*ShimFastHook)(const ShimFastCall&, void* user);
void shim_set_fast_hook(ShimFastHook hook, void* user);

}  // namespace cv

#define CV_Error(code, msg) throw cv::Exception(std::string(msg))
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(std::string("Assertion failed: ") + #expr); } while (0)
