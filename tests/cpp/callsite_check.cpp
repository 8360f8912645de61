// callsite_check.cpp — the SAME call-site source, instantiated once on the reference's own class and once on the drop-in.
//
//   template arm A: ORB_SLAM3::ORBextractor — the reference's ORBextractor.cpp, compiled unmodified (oracle/_ref/ORBextractor.o)
//   template arm B: orbx::ORBextractor      — dynamic-visual-slam_b200/host/ORBextractor.hpp in its OpenCV build mode
// Both see the cv:: types of the header shim (oracle/ref_shim; this image has no OpenCV C++ headers), so the call below is the
// frontend's call (reference frontend.cpp:1094-1095: `(*orb_extractor_)(gray, cv::noArray(), keypoints, descriptors, vLappingArea)`)
// character for character.  Built by `make -C oracle _ref` into oracle/_ref/callsite_check (it needs the reference header, which
// only exists in the build container); tests/test_host_adapter.py runs it on the GPU box.
//   callsite_check W H gray.raw        exit 0 = every comparison equal, 1 = mismatch (printed), 3 = no CUDA device
#define ORBX_WITH_OPENCV
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dynamic_visual_slam/ORBextractor.hpp"
#include "../../dynamic-visual-slam_b200/host/ORBextractor.hpp"
#include "../../oracle/orb_oracle.h"

struct Result { int ret; std::vector<cv::KeyPoint> kps; cv::Mat desc; };

template <class Extractor>
static Result call_like_the_frontend(Extractor &orb_extractor_, const cv::Mat &gray, std::vector<int> vLappingArea)
{
    Result r;
    r.ret = orb_extractor_(gray, cv::noArray(), r.kps, r.desc, vLappingArea);     // frontend.cpp:1094-1095
    return r;
}

static bool same(const Result &a, const Result &b, const char *what)
{
    bool ok = a.ret == b.ret && a.kps.size() == b.kps.size() && a.desc.rows == b.desc.rows;
    if (ok && !a.kps.empty()) ok = std::memcmp(a.kps.data(), b.kps.data(), a.kps.size() * sizeof(cv::KeyPoint)) == 0;
    for (int r = 0; ok && r < a.desc.rows; r++) ok = std::memcmp(a.desc.ptr(r), b.desc.ptr(r), 32) == 0;
    std::printf("%-34s ret %d / %d, %zu / %zu keypoints: %s\n", what, a.ret, b.ret, a.kps.size(), b.kps.size(), ok ? "equal" : "DIFFERENT");
    return ok;
}

int main(int argc, char **argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: callsite_check W H gray.raw\n"); return 2; }
    const int W = std::atoi(argv[1]), H = std::atoi(argv[2]);
    std::vector<uint8_t> px((size_t)W * H);
    FILE *f = std::fopen(argv[3], "rb");
    if (!f || std::fread(px.data(), 1, px.size(), f) != px.size()) { std::fprintf(stderr, "cannot read %s\n", argv[3]); return 2; }
    std::fclose(f);
    try {
        ORB_SLAM3::ORBextractor reference(1000, 1.2f, 8, 20, 7);                       // frontend.cpp:205-211
        orbx::ORBextractor dropin(1000, 1.2f, 8, 20, 7, W, H);
        cv::Mat gray(H, W, CV_8UC1, px.data(), (size_t)W);
        bool ok = true;
        const int laps[][2] = { {0, 0}, {W / 4, W / 2}, {0, W}, {W, 2 * W} };
        for (const auto &lap : laps) {
            char what[64];
            std::snprintf(what, sizeof(what), "operator() vLappingArea {%d,%d}", lap[0], lap[1]);
            ok &= same(call_like_the_frontend(reference, gray, {lap[0], lap[1]}), call_like_the_frontend(dropin, gray, {lap[0], lap[1]}), what);
        }
        ok &= same(call_like_the_frontend(reference, cv::Mat(), {0, 0}), call_like_the_frontend(dropin, cv::Mat(), {0, 0}), "empty image");
        // getters (ORBextractor.hpp:62-82)
        ok &= reference.GetLevels() == dropin.GetLevels() && reference.GetScaleFactor() == dropin.GetScaleFactor() &&
              reference.GetScaleFactors() == dropin.GetScaleFactors() && reference.GetInverseScaleFactors() == dropin.GetInverseScaleFactors() &&
              reference.GetScaleSigmaSquares() == dropin.GetScaleSigmaSquares() && reference.GetInverseScaleSigmaSquares() == dropin.GetInverseScaleSigmaSquares();
        std::printf("getters: %s\n", ok ? "equal" : "DIFFERENT");
        // matcher_.match(desc1, desc0, matches) (frontend.cpp:1123) — the reference side is cv::BFMatcher, restated in the C oracle
        Result a = call_like_the_frontend(dropin, gray, {0, 0});
        std::vector<cv::DMatch> got;
        orbx::BFMatcher matcher_(dropin);
        matcher_.match(a.desc, a.desc, got);
        std::vector<uint8_t> rows((size_t)a.desc.rows * 32);
        for (int r = 0; r < a.desc.rows; r++) std::memcpy(rows.data() + (size_t)r * 32, a.desc.ptr(r), 32);
        std::vector<orc_dmatch> want((size_t)a.desc.rows);
        orc_match(rows.data(), a.desc.rows, rows.data(), a.desc.rows, want.data(), 1);
        bool mok = got.size() == want.size() && std::memcmp(got.data(), want.data(), got.size() * sizeof(cv::DMatch)) == 0;
        std::printf("matcher_.match: %zu matches: %s\n", got.size(), mok ? "equal" : "DIFFERENT");
        return ok && mok ? 0 : 1;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "callsite error: %s\n", e.what());
        return 3;
    }
}
