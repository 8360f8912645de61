// adapter_check.cpp — drives the reference-shaped C++ adapter (host/ORBextractor.hpp) exactly as the frontend does
// (reference frontend.cpp:205-211 ctor, :1094-1095 operator(), :1100 filterDepth, :1123-1132 match + distance < 50)
// on two raw frames and dumps the results for the Python test to compare with the oracle.
//   adapter_check W H gray0.raw gray1.raw depth1.raw out.bin      exit 0 ok, 3 = no CUDA device (loud failure)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../dynamic-visual-slam_b200/host/ORBextractor.hpp"

static std::vector<uint8_t> slurp(const char *path, size_t n)
{
    std::vector<uint8_t> v(n);
    FILE *f = fopen(path, "rb");
    if (!f || fread(v.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: adapter_check W H gray0 gray1 depth1 out\n"); return 2; }
    const int W = atoi(argv[1]), H = atoi(argv[2]);
    try {
        orbx::ORBextractor extractor(1000, 1.2f, 8, 20, 7, W, H);              // frontend.cpp:205-211
        orbx::BFMatcher matcher(extractor);                                     // frontend.cpp:220
        std::vector<uint8_t> g0 = slurp(argv[3], (size_t)W * H), g1 = slurp(argv[4], (size_t)W * H), d1 = slurp(argv[5], (size_t)W * H * 2);
        orbx::Mat gray0(H, W, orbx::type_8uc1(), g0.data(), (size_t)W), gray1(H, W, orbx::type_8uc1(), g1.data(), (size_t)W);
        orbx::Mat depth1(H, W, orbx::type_16uc1(), d1.data(), (size_t)W * 2);
        std::vector<orbx::KeyPoint> k0, k1, k1f;
        orbx::Mat desc0, desc1, desc1f;
        std::vector<int> vLappingArea = { 0, 0 };                               // frontend.cpp:290
        const int n0 = extractor(gray0, orbx::Mat(), k0, desc0, vLappingArea);
        const int n1 = extractor(gray1, orbx::Mat(), k1, desc1, vLappingArea);
        const int n1f = extractor.extractFiltered(gray1, depth1, k1f, desc1f);
        std::vector<orbx::DMatch> matches, good;
        matcher.match(desc1, desc0, matches);
        matcher.matchBelow(desc1, desc0, 50.0f, good);
        const int empty_rc = extractor(orbx::Mat(), orbx::Mat(), k1f, desc1f, vLappingArea);     // must be -1
        const std::vector<orbx::Mat> &pyr = extractor.imagePyramid();
        FILE *f = fopen(argv[6], "wb");
        if (!f) return 2;
        int32_t hdr[8] = { n0, n1, n1f, (int32_t)matches.size(), (int32_t)good.size(), empty_rc, extractor.GetLevels(), (int32_t)pyr.size() };
        fwrite(hdr, sizeof(hdr), 1, f);
        fwrite(k0.data(), sizeof(orbx::KeyPoint), k0.size(), f);
        for (int r = 0; r < desc0.rows; r++) fwrite(desc0.ptr<uint8_t>(r), 1, 32, f);
        fwrite(k1.data(), sizeof(orbx::KeyPoint), k1.size(), f);
        for (int r = 0; r < desc1.rows; r++) fwrite(desc1.ptr<uint8_t>(r), 1, 32, f);
        fwrite(matches.data(), sizeof(orbx::DMatch), matches.size(), f);
        fwrite(good.data(), sizeof(orbx::DMatch), good.size(), f);
        std::vector<float> sf = extractor.GetScaleFactors();
        fwrite(sf.data(), sizeof(float), sf.size(), f);
        // feature culling for the backend (frontend.cpp:1168-1218) on frame 1 with the distance-filtered matches standing in for the RANSAC inliers
        std::vector<orbx::KeyPoint> bk;
        orbx::Mat bd;
        extractor.cullForBackend(k1, desc1, good, bk, bd);
        const int32_t nb = (int32_t)bk.size();
        fwrite(&nb, sizeof(nb), 1, f);
        fwrite(bk.data(), sizeof(orbx::KeyPoint), bk.size(), f);
        for (int r = 0; r < bd.rows; r++) fwrite(bd.ptr<uint8_t>(r), 1, 32, f);
        fclose(f);
        printf("adapter ok: %d %d %d keypoints, %zu matches, %zu good\n", n0, n1, n1f, matches.size(), good.size());
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "adapter error: %s\n", e.what());
        return 3;
    }
}
