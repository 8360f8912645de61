// ref_glue.cpp — ORACLE (test infrastructure): C entry points around the REFERENCE's own ORB_SLAM3::ORBextractor,
// compiled unmodified from /root/reference/dynamic_visual_slam/src/ORBextractor.cpp by `make -C oracle _ref`
// (output oracle/_ref/libref_orbextractor.so, git-ignored).  Used by tests/ to pin oracle/orb_oracle.c — and with it the
// golden vectors and the CUDA path — to the reference's real std::list / std::sort / DivideNode code, and by bench.py's
// CPU legs (cpu_baseline.kind = "reference").  The product never links or loads it.
//
// The protected stages are reached through a subclass (no edit of the reference): DistributeOctTree (:555-779),
// ComputePyramid (:1169-1194), ComputeKeyPointsOctTree (:781-896), ComputeKeyPointsOld (:898-1075, dead in the reference).
#include <cstdint>
#include <cstring>
#include <new>
#include <stdexcept>
#include <vector>

#include "dynamic_visual_slam/ORBextractor.hpp"
#include "orb_oracle.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct RefExtractor : public ORB_SLAM3::ORBextractor {
    using ORB_SLAM3::ORBextractor::ORBextractor;
    using ORB_SLAM3::ORBextractor::ComputeKeyPointsOctTree;
    using ORB_SLAM3::ORBextractor::ComputeKeyPointsOld;
    using ORB_SLAM3::ORBextractor::ComputePyramid;
    using ORB_SLAM3::ORBextractor::DistributeOctTree;
    using ORB_SLAM3::ORBextractor::mnFeaturesPerLevel;
    using ORB_SLAM3::ORBextractor::umax;
    int nf, nl, ini, mn;
    float sf;
};

struct FastTrace {
    RefExtractor* ex;
    std::vector<std::vector<orc_cand>> per_level;     // what each cell's LAST FAST call returned, in call order
    std::vector<int> calls_th_ini, calls_th_min;
    // the reference calls FAST(th=ini) and, iff that returned nothing, FAST(th=min) on the same view (:826-846); the retry
    // replaces an empty list, so appending every call's output gives the candidate list it builds
};

void fast_hook(const cv::ShimFastCall& c, void* user)
{
    FastTrace* t = (FastTrace*)user;
    for (size_t l = 0; l < t->ex->mvImagePyramid.size(); l++) {
        const cv::Mat& lv = t->ex->mvImagePyramid[l];
        if (lv.datastart != c.datastart) continue;
        size_t off = (size_t)(lv.data - lv.datastart);
        const int ly = (int)(off / lv.step), lx = (int)(off % lv.step);     // the level view's origin inside its padded buffer
        const int bx = c.x0 - lx - 16, by = c.y0 - ly - 16;                 // relative to (minBorderX, minBorderY) = (16, 16)
        for (const cv::KeyPoint& k : c.out) {
            orc_cand oc;
            oc.x = (int)k.pt.x + bx; oc.y = (int)k.pt.y + by; oc.score = (int)k.response;
            t->per_level[l].push_back(oc);
        }
        (c.threshold == t->ex->ini ? t->calls_th_ini : t->calls_th_min)[l]++;
        return;
    }
}

void copy_kp(const cv::KeyPoint& k, orc_keypoint* o)
{
    static_assert(sizeof(cv::KeyPoint) == sizeof(orc_keypoint), "KeyPoint layout");
    std::memcpy(o, &k, sizeof(orc_keypoint));
}

}  // namespace

extern "C" {

void* ref_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
{
    RefExtractor* e = new RefExtractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
    e->nf = nfeatures; e->nl = nlevels; e->ini = iniThFAST; e->mn = minThFAST; e->sf = scaleFactor;
    return e;
}
void ref_destroy(void* h) { delete (RefExtractor*)h; }

// ctor tables (ORBextractor.cpp:409-469) through the reference's getters
void ref_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* nfeat_level, int* umax16)
{
    RefExtractor* e = (RefExtractor*)h;
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(), d = e->GetInverseScaleSigmaSquares();
    for (int i = 0; i < e->GetLevels(); i++) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; nfeat_level[i] = e->mnFeaturesPerLevel[i]; }
    for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}

// ORBextractor::operator() (:1086-1167).  Returns its return value (monoIndex; -1 on an empty image); *n_total = keypoints written.
int ref_extract(void* h, const uint8_t* gray, int w, int hgt, size_t step, int lap0, int lap1,
                orc_keypoint* kps, uint8_t* desc, int cap, int* n_total)
{
    RefExtractor* e = (RefExtractor*)h;
    cv::Mat img = (gray && w > 0 && hgt > 0) ? cv::Mat(hgt, w, CV_8UC1, (void*)gray, step) : cv::Mat();
    std::vector<cv::KeyPoint> keys;
    cv::Mat d;
    std::vector<int> lap = {lap0, lap1};
    int r;
    // degenerate pyramids (a level of 32 rows, or thinner than the 16-px border on one axis only) make the reference compute a negative
    // or infinite root count (:559) and throw from std::vector::resize (:566); the frontend catches std::exception per frame
    // (frontend.cpp:1319-1323); a level that rounds to zero pixels makes cv::resize throw.  Reported as -3 so a test can pin exactly where.
    try { r = (*e)(img, cv::noArray(), keys, d, lap); }
    catch (const std::exception&) { if (n_total) *n_total = 0; return -3; }
    if (n_total) *n_total = (int)keys.size();
    if (r < 0) return r;
    if ((int)keys.size() > cap) return -2;
    for (size_t i = 0; i < keys.size(); i++) copy_kp(keys[i], kps + i);
    for (int i = 0; i < d.rows; i++) std::memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
    return r;
}

// frame-parallel batch for the CPU baseline: one reference extractor per thread (the class is stateful)
int ref_extract_batch(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                      const uint8_t* gray, int nframes, int w, int hgt,
                      orc_keypoint* kps, uint8_t* desc, int cap_per_frame, int32_t* counts, int nthreads)
{
    int bad = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        void* h = ref_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int f = 0; f < nframes; f++) {
            int n = 0;
            int r = ref_extract(h, gray + (size_t)f * w * hgt, w, hgt, (size_t)w, 0, 0,
                                kps + (size_t)f * cap_per_frame, desc + (size_t)f * cap_per_frame * 32, cap_per_frame, &n);
            counts[f] = r < 0 ? r : n;
            if (r < 0) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
                bad = 1;
            }
        }
        ref_destroy(h);
    }
    return bad ? -1 : 0;
}

// public member mvImagePyramid after the last ref_extract / ref_stage_trace: copies level `level` tightly packed
int ref_get_level(void* h, int level, uint8_t* out, int* w, int* hgt)
{
    RefExtractor* e = (RefExtractor*)h;
    if (level < 0 || level >= (int)e->mvImagePyramid.size() || e->mvImagePyramid[level].empty()) return -1;
    const cv::Mat& m = e->mvImagePyramid[level];
    *w = m.cols; *hgt = m.rows;
    if (out) for (int y = 0; y < m.rows; y++) std::memcpy(out + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
    return 0;
}
// the 19-px REFLECT_101 ring ComputePyramid writes around each level (:1184-1190): padded buffer, tightly packed
int ref_get_level_padded(void* h, int level, uint8_t* out, int* w, int* hgt)
{
    RefExtractor* e = (RefExtractor*)h;
    if (level < 0 || level >= (int)e->mvImagePyramid.size() || e->mvImagePyramid[level].empty()) return -1;
    const cv::Mat& m = e->mvImagePyramid[level];
    const int pw = m.cols + 38, ph = m.rows + 38;
    *w = pw; *hgt = ph;
    if (out) for (int y = 0; y < ph; y++) std::memcpy(out + (size_t)y * pw, m.datastart + (size_t)y * m.step, (size_t)pw);
    return 0;
}

// ComputePyramid + ComputeKeyPointsOctTree with every cv::FAST call observed: per level the candidate list handed to
// DistributeOctTree (relative to the border box, the reference's push order) and the retained keypoints WITH orientation
// (level coordinates, before operator()'s `pt *= scale`).  cands/keys: per level `cap` entries.
int ref_stage_trace(void* h, const uint8_t* gray, int w, int hgt, size_t step,
                    orc_cand* cands, int32_t* ncands, orc_keypoint* keys, int32_t* nkeys, int cap,
                    int32_t* calls_ini, int32_t* calls_min)
{
    RefExtractor* e = (RefExtractor*)h;
    cv::Mat img(hgt, w, CV_8UC1, (void*)gray, step);
    e->ComputePyramid(img);
    FastTrace t;
    t.ex = e;
    t.per_level.resize(e->nl); t.calls_th_ini.assign(e->nl, 0); t.calls_th_min.assign(e->nl, 0);
    cv::shim_set_fast_hook(fast_hook, &t);
    std::vector<std::vector<cv::KeyPoint>> all;
    e->ComputeKeyPointsOctTree(all);
    cv::shim_set_fast_hook(nullptr, nullptr);
    int rc = 0;
    for (int l = 0; l < e->nl; l++) {
        ncands[l] = (int32_t)t.per_level[l].size();
        nkeys[l] = (int32_t)all[l].size();
        if (calls_ini) calls_ini[l] = t.calls_th_ini[l];
        if (calls_min) calls_min[l] = t.calls_th_min[l];
        if (ncands[l] > cap || nkeys[l] > cap) { rc = -2; continue; }
        if (cands) std::memcpy(cands + (size_t)l * cap, t.per_level[l].data(), sizeof(orc_cand) * t.per_level[l].size());
        if (keys) for (size_t i = 0; i < all[l].size(); i++) copy_kp(all[l][i], keys + (size_t)l * cap + i);
    }
    return rc;
}

// DistributeOctTree (:555-779) on a caller-made candidate list (x, y relative to the box, score = response)
int ref_distribute_octtree(void* h, const orc_cand* cands, int n, int minX, int maxX, int minY, int maxY, int N, orc_cand* out, int cap)
{
    RefExtractor* e = (RefExtractor*)h;
    std::vector<cv::KeyPoint> v((size_t)n);
    for (int i = 0; i < n; i++) v[i] = cv::KeyPoint((float)cands[i].x, (float)cands[i].y, 7.f, -1, (float)cands[i].score);
    std::vector<cv::KeyPoint> r = e->DistributeOctTree(v, minX, maxX, minY, maxY, N, 0);
    if ((int)r.size() > cap) return -2;
    for (size_t i = 0; i < r.size(); i++) { out[i].x = (int)r[i].pt.x; out[i].y = (int)r[i].pt.y; out[i].score = (int)r[i].response; }
    return (int)r.size();
}

// ComputePyramid + the reference's dead per-cell top-N variant ComputeKeyPointsOld (:898-1075; retainBest :1049, :1067)
int ref_keypoints_old(void* h, const uint8_t* gray, int w, int hgt, size_t step, orc_keypoint* keys, int32_t* nkeys, int cap)
{
    RefExtractor* e = (RefExtractor*)h;
    cv::Mat img(hgt, w, CV_8UC1, (void*)gray, step);
    e->ComputePyramid(img);
    std::vector<std::vector<cv::KeyPoint>> all;
    e->ComputeKeyPointsOld(all);
    int rc = 0;
    for (int l = 0; l < e->nl; l++) {
        nkeys[l] = (int32_t)all[l].size();
        if (nkeys[l] > cap) { rc = -2; continue; }
        for (size_t i = 0; i < all[l].size(); i++) copy_kp(all[l][i], keys + (size_t)l * cap + i);
    }
    return rc;
}

const char* ref_source_path(void) { return REF_SOURCE_PATH; }

}  // extern "C"
