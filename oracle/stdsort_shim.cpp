// stdsort_shim.cpp — ORACLE support (test infrastructure).  Calls the REAL libstdc++ std::sort with
// a comparator of the same shape as the reference's compareNodes (ORBextractor.cpp:538-553) on
// vector<pair<int, Node*>>, so the emulation in orb_oracle.c / the CUDA kernel can be pinned to the
// tie order the reference's own sort call (ORBextractor.cpp:700) produces.
#include <algorithm>
#include <utility>
#include <vector>
namespace {
struct Node { int ULx; int payload; };
bool cmp(std::pair<int, Node*>& e1, std::pair<int, Node*>& e2)
{
    if (e1.first < e2.first) return true;
    if (e1.first > e2.first) return false;
    return e1.second->ULx < e2.second->ULx;
}
}
extern "C" void real_std_sort(int* cnt, int* ulx, int* payload, int n)
{
    std::vector<Node> nodes(n);
    std::vector<std::pair<int, Node*>> v;
    v.reserve(n);
    for (int i = 0; i < n; i++) { nodes[i].ULx = ulx[i]; nodes[i].payload = payload[i]; v.push_back(std::make_pair(cnt[i], &nodes[i])); }
    std::sort(v.begin(), v.end(), cmp);
    std::vector<int> c(n), u(n), p(n);
    for (int i = 0; i < n; i++) { c[i] = v[i].first; u[i] = v[i].second->ULx; p[i] = v[i].second->payload; }
    for (int i = 0; i < n; i++) { cnt[i] = c[i]; ulx[i] = u[i]; payload[i] = p[i]; }
}

// the frontend's sort of unmatched features (reference frontend.cpp:1193-1202): vector<pair<float, int>>, comparator a.first > b.first
extern "C" void real_std_sort_response_desc(float* response, int* index, int n)
{
    std::vector<std::pair<float, int>> v;
    v.reserve(n);
    for (int i = 0; i < n; i++) v.push_back({response[i], index[i]});
    std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    for (int i = 0; i < n; i++) { response[i] = v[i].first; index[i] = v[i].second; }
}
