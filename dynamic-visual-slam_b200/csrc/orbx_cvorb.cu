// orbx_cvorb.cu — "profile C": cv::ORB's pipeline (north_star stage list; the extractor of the reference's gtest,
// test/test_dbow2_integration.cpp:19,38: cv::ORB::create(n)->detectAndCompute; BASELINE configs[0]) behind the same C ABI as the frontend's
// ORB_SLAM3::ORBextractor ("profile S").  orbx_params.profile = ORBX_PROFILE_CVORB selects it; depth / box post-filters and the matcher are
// shared.  The arithmetic is OpenCV 4.x's (features2d/src/orb.cpp), restated in oracle/cvorb_oracle.py and pinned against cv2 4.13.0:
//   pyramid    level sizes cvRound(cols / scale_l), chained cv::resize(INTER_LINEAR_EXACT): 8.8 fixed-point coefficients from the double
//              source coordinate, exact 16-bit horizontal pass, one rounding after the vertical pass            (k_resize_exact)
//   detection  whole-level FAST-9/16 (threshold, 3x3 NMS over the level, not per cell), points within 31 px of the border dropped,
//              KeyPointsFilter::retainBest(2 N_l) by FAST score with every tie at the cut-off kept              (k_cfast_score, k_cfast_nms, k_cretain)
//   scoring    HarrisResponses: 7x7 block of Sobel-like derivatives, k = 0.04, fp32 expression in OpenCV's order, then retainBest(N_l)
//              by Harris, ties kept                                                                              (k_cretain)
//   blur       cv::GaussianBlur(7x7, sigma 2) of a SUB-MATRIX = OpenCV's float path: fp32 FMAs left to right, then symmetric pairs (k_cblur_f32)
//   describe   IC_Angle + rBRIEF-256 on the blurred level (k_describe_c in k_describe.cu), pt *= scale_l, size = 31 * scale_l
// Output: levels in order, inside a level sorted by (Harris response descending, y, x) — OpenCV leaves std::nth_element's order there,
// so parity with cv2 is on SETS (north_star: "identical except for ties at the retention cutoff"), parity with the oracle on arrays.
// One frame per pipeline run (the batch entry points loop over the frames on the handle's stream).
#include "orbx_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>

#define CV_EDGE 31                    // edgeThreshold
#define CV_LIST_CAP 4096              // per level: points kept by retainBest(2 N) incl. ties (sorted in shared memory)
#define CV_TILE_W 32
#define CV_TILE_H 8

struct CvLevels {
    int nlevels;
    int w[ORBX_MAX_LEVELS], h[ORBX_MAX_LEVELS], pitch[ORBX_MAX_LEVELS];
    size_t off[ORBX_MAX_LEVELS];          // level l >= 1 inside d_pyr; all levels inside d_score / d_blur
    size_t boff[ORBX_MAX_LEVELS];
    int N[ORBX_MAX_LEVELS];               // features per level
    int cand_cap[ORBX_MAX_LEVELS]; size_t cand_off[ORBX_MAX_LEVELS];
    int tile_first[ORBX_MAX_LEVELS + 1];  // 32 x 8 tiles of every level, flattened
    float scale[ORBX_MAX_LEVELS];
    int xtab[ORBX_MAX_LEVELS], ytab[ORBX_MAX_LEVELS];
};

struct orbx_cvorb {
    CvLevels L; int width, height;
    uint8_t *d_pyr, *d_score, *d_blur; size_t pyr_bytes, lvl_bytes;
    ResizeTab *d_xtab, *d_ytab; int tab_cap;
    uint32_t *d_cand; size_t cand_entries; int32_t *d_ncand;
    uint32_t *d_fxy; float *d_fresp; int32_t *d_fcount;      // final per-level lists [level][CV_LIST_CAP]
    size_t pyr_cap, lvl_cap, cand_cap_total;
};

// ---- cv::resize INTER_LINEAR_EXACT ----
static void exact_table(int ssize, int dsize, ResizeTab *out)
{
    const double scale = 1.0 / ((double)dsize / (double)ssize);
    for (int x = 0; x < dsize; x++) {
        const double fval = scale * ((double)x + 0.5) - 0.5;
        int ival = (int)floor(fval), c0, c1;
        if (ival >= 0 && ssize > 1) {
            if (ival < ssize - 1) { c1 = (int)lrint((fval - (double)ival) * 256.0); c0 = 256 - c1; }
            else { ival = ssize - 2; c0 = 0; c1 = 256; }
        } else { ival = 0; c0 = 256; c1 = 0; }
        out[x].ofs = ival; out[x].a0 = (int16_t)c0; out[x].a1 = (int16_t)c1;
    }
}

__global__ void __launch_bounds__(256) k_resize_exact(const uint8_t *__restrict__ src, int sstep, int sw, int sh, uint8_t *__restrict__ dst, int dstep,
                                                      int dw, int dh, const ResizeTab *__restrict__ xt, const ResizeTab *__restrict__ yt)
{
    ORBX_PDL_ENTRY();
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= dw || y >= dh) return;
    const ResizeTab tx = xt[x], ty = yt[y];
    const int x1 = min(tx.ofs + 1, sw - 1), y1 = min(ty.ofs + 1, sh - 1);
    const uint8_t *S0 = src + (size_t)ty.ofs * sstep, *S1 = src + (size_t)y1 * sstep;
    const uint32_t h0 = (uint32_t)__ldg(S0 + tx.ofs) * (uint32_t)tx.a0 + (uint32_t)__ldg(S0 + x1) * (uint32_t)tx.a1;
    const uint32_t h1 = (uint32_t)__ldg(S1 + tx.ofs) * (uint32_t)tx.a0 + (uint32_t)__ldg(S1 + x1) * (uint32_t)tx.a1;
    const uint32_t v = (h0 * (uint32_t)ty.a0 + h1 * (uint32_t)ty.a1 + 32768u) >> 16;
    dst[(size_t)y * dstep + x] = (uint8_t)min(v, 255u);
}

// ---- whole-level FAST-9/16: score map (score - 1 for corners, 0 elsewhere) ----
struct CvImgs { const uint8_t *img[ORBX_MAX_LEVELS]; int step[ORBX_MAX_LEVELS]; };

__device__ __forceinline__ int cv_tile_level(const CvLevels &L, int t)
{
    int l = 0;
    while (l + 1 < L.nlevels && t >= L.tile_first[l + 1]) l++;
    return l;
}

// S = max over the 16 arcs of 9 contiguous ring pixels of min(I(p) - ring) resp. min(ring - I(p)), on packed u16x2 lanes (lo = r, hi = 255 - r)
__device__ __forceinline__ int cv_fast_score(const uint8_t *p, int st)
{
    const int v = p[0];
    uint32_t r[16];
#define PK(x) ((uint32_t)(x) * 0xFFFF0001u + 0x00FF0000u)
    r[0] = PK(p[3 * st]);       r[1] = PK(p[3 * st + 1]);   r[2] = PK(p[2 * st + 2]);   r[3] = PK(p[st + 3]);
    r[4] = PK(p[3]);            r[5] = PK(p[-st + 3]);      r[6] = PK(p[-2 * st + 2]);  r[7] = PK(p[-3 * st + 1]);
    r[8] = PK(p[-3 * st]);      r[9] = PK(p[-3 * st - 1]);  r[10] = PK(p[-2 * st - 2]); r[11] = PK(p[-st - 3]);
    r[12] = PK(p[-3]);          r[13] = PK(p[st - 3]);      r[14] = PK(p[2 * st - 2]);  r[15] = PK(p[3 * st - 1]);
#undef PK
    uint32_t m3[16], best = 0u;
#pragma unroll
    for (int k = 0; k < 16; k++) m3[k] = __vminu2(__vminu2(r[k], r[(k + 1) & 15]), r[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) best = __vmaxu2(best, __vminu2(__vminu2(m3[k], m3[(k + 3) & 15]), m3[(k + 6) & 15]));
    const int s_bright = (int)(best & 0xFFFFu) - v, s_dark = v - (255 - (int)(best >> 16));
    return max(s_dark, s_bright);
}

__global__ void __launch_bounds__(CV_TILE_W * CV_TILE_H) k_cfast_score(CvLevels L, CvImgs I, uint8_t *__restrict__ score, int th)
{
    ORBX_PDL_ENTRY();
    const int l = cv_tile_level(L, blockIdx.x);
    const int t = blockIdx.x - L.tile_first[l], tx = (L.w[l] + CV_TILE_W - 1) / CV_TILE_W;
    const int x = (t % tx) * CV_TILE_W + (threadIdx.x & 31), y = (t / tx) * CV_TILE_H + (threadIdx.x >> 5);
    const int w = L.w[l], h = L.h[l];
    if (x >= w || y >= h) return;
    uint8_t out = 0;
    if (x >= 3 && y >= 3 && x < w - 3 && y < h - 3) {
        const int st = I.step[l];
        const uint8_t *p = I.img[l] + (size_t)y * st + x;
        const int v = p[0], hi = v + th, lo = v - th;
        // exact necessary condition: each opposite ring pair holds a member of any 9-arc
        int a = p[3 * st], b = p[-3 * st];
        bool br = (a > hi) | (b > hi), dk = (a < lo) | (b < lo);
        a = p[3]; b = p[-3];
        br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
        if (br | dk) {
            a = p[2 * st + 2]; b = p[-2 * st - 2];
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            a = p[-2 * st + 2]; b = p[2 * st - 2];
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            if (br | dk) {
                const int s = cv_fast_score(p, st);
                if (s > th) out = (uint8_t)(s - 1);
            }
        }
    }
    score[L.boff[l] + (size_t)y * L.pitch[l] + x] = out;
}

// strict 3x3 NMS over the whole level + runByImageBorder(31) -> unordered candidate list per level
__global__ void __launch_bounds__(CV_TILE_W * CV_TILE_H) k_cfast_nms(CvLevels L, const uint8_t *__restrict__ score, uint32_t *__restrict__ cand, int32_t *__restrict__ ncand,
                                                                      int32_t *__restrict__ status)
{
    ORBX_PDL_ENTRY();
    const int l = cv_tile_level(L, blockIdx.x);
    const int t = blockIdx.x - L.tile_first[l], tx = (L.w[l] + CV_TILE_W - 1) / CV_TILE_W;
    const int x = (t % tx) * CV_TILE_W + (threadIdx.x & 31), y = (t / tx) * CV_TILE_H + (threadIdx.x >> 5);
    const int w = L.w[l], h = L.h[l], sp = L.pitch[l];
    if (x < CV_EDGE || y < CV_EDGE || x >= w - CV_EDGE || y >= h - CV_EDGE) return;
    const uint8_t *q = score + L.boff[l] + (size_t)y * sp + x;
    const int s = q[0];
    if (!s) return;
    const int m = max(max(max(q[-sp - 1], q[-sp]), max(q[-sp + 1], q[-1])), max(max(q[1], q[sp - 1]), max(q[sp], q[sp + 1])));
    if (s <= m) return;
    const int i = atomicAdd(&ncand[l], 1);
    if (i < L.cand_cap[l]) cand[L.cand_off[l] + i] = orbx_pack(x, y, s);
    else atomicOr(status, ORBX_DS_CAND_OVERFLOW);
}

// ---- retainBest(2N) by FAST score -> HarrisResponses -> retainBest(N) by Harris; one CTA per level ----
__device__ __forceinline__ float cv_harris(const uint8_t *img, int st, int x0, int y0)
{
    int a = 0, b = 0, c = 0;
    for (int u = -3; u <= 3; u++) for (int v = -3; v <= 3; v++) {
        const uint8_t *p = img + (size_t)(y0 + u) * st + (x0 + v);
        const int Ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-st + 1] - (int)p[-st - 1]) + ((int)p[st + 1] - (int)p[st - 1]);
        const int Iy = ((int)p[st] - (int)p[-st]) * 2 + ((int)p[st - 1] - (int)p[-st - 1]) + ((int)p[st + 1] - (int)p[-st + 1]);
        a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
    }
    const float scale = __fdiv_rn(1.f, __fmul_rn((float)(4 * 7), 255.f));
    const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float fa = (float)a, fb = (float)b, fc = (float)c, sum = __fadd_rn(fa, fb);
    return __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, sum), sum)), s4);
}
__device__ __forceinline__ uint32_t cv_ford(float f) { const uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }   // monotone

__global__ void __launch_bounds__(1024) k_cretain(CvLevels L, CvImgs I, const uint32_t *__restrict__ cand, const int32_t *__restrict__ ncand,
                                                  uint32_t *__restrict__ fxy, float *__restrict__ fresp, int32_t *__restrict__ fcount, int32_t *__restrict__ status)
{
    __shared__ unsigned long long s_key[CV_LIST_CAP];
    __shared__ int s_hist[256];
    __shared__ int s_cut, s_n2;
    ORBX_PDL_ENTRY();
    const int l = blockIdx.x, tid = threadIdx.x;
    const int n = min(ncand[l], L.cand_cap[l]);
    const uint32_t *cl = cand + L.cand_off[l];
    const int N = L.N[l];
    if (tid < 256) s_hist[tid] = 0;
    if (tid == 0) s_n2 = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) atomicAdd(&s_hist[orbx_ps(cl[i])], 1);
    __syncthreads();
    if (tid == 0) {                                   // retainBest(2N): the score of the 2N-th best; everything >= it is kept
        int cut = 0;
        if (n > 2 * N) { int acc = 0; for (cut = 255; cut > 0; cut--) { acc += s_hist[cut]; if (acc >= 2 * N) break; } if (2 * N == 0) cut = 256; }
        s_cut = cut;
    }
    __syncthreads();
    const int cut = s_cut;
    for (int i = tid; i < n; i += blockDim.x) {
        const uint32_t c = cl[i];
        if (orbx_ps(c) >= cut) {
            const int o = atomicAdd(&s_n2, 1);
            if (o < CV_LIST_CAP) {
                const int x = orbx_px(c), y = orbx_py(c);
                const float hr = cv_harris(I.img[l], I.step[l], x, y);
                s_key[o] = ((unsigned long long)(~cv_ford(hr)) << 32) | (unsigned)((y << 16) | x);      // ascending key = (Harris descending, y, x)
            }
        }
    }
    __syncthreads();
    int n2 = s_n2;
    if (n2 > CV_LIST_CAP) { if (tid == 0) atomicOr(status, ORBX_DS_NODE_OVERFLOW); n2 = CV_LIST_CAP; }
    int np2 = 1;
    while (np2 < n2) np2 <<= 1;
    for (int i = n2 + tid; i < np2; i += blockDim.x) s_key[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)                // bitonic sort, ascending
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long a = s_key[i], b = s_key[p];
                    if (((i & k) == 0) ? (a > b) : (a < b)) { s_key[i] = b; s_key[p] = a; }
                }
            }
            __syncthreads();
        }
    // retainBest(N) by Harris: the first N and every further point whose response equals the N-th
    int keep = n2;
    if (n2 > N) {
        if (N == 0) keep = 0;
        else {
            const unsigned thr = (unsigned)(s_key[N - 1] >> 32);
            __shared__ int s_keep;
            if (tid == 0) s_keep = N;
            __syncthreads();
            for (int i = N + tid; i < n2; i += blockDim.x) if ((unsigned)(s_key[i] >> 32) == thr) atomicMax(&s_keep, i + 1);
            __syncthreads();
            keep = s_keep;
        }
    }
    for (int i = tid; i < keep; i += blockDim.x) {
        const unsigned long long k = s_key[i];
        const uint32_t o = ~(unsigned)(k >> 32);
        fxy[(size_t)l * CV_LIST_CAP + i] = (unsigned)k;
        fresp[(size_t)l * CV_LIST_CAP + i] = __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
    }
    if (tid == 0) fcount[l] = keep;
}

// ---- GaussianBlur(7x7, sigma 2) through OpenCV's float path: 32 x 32 output tile, row pass into shared floats, column pass ----
__global__ void __launch_bounds__(256) k_cblur_f32(CvLevels L, CvImgs I, uint8_t *__restrict__ blur, const int *__restrict__ btile_first)
{
    __shared__ float s_row[38][33];
    ORBX_PDL_ENTRY();
    int l = 0;
    while (l + 1 < L.nlevels && (int)blockIdx.x >= btile_first[l + 1]) l++;
    const int w = L.w[l], h = L.h[l], st = I.step[l];
    const int tx = (w + 31) / 32, t = blockIdx.x - btile_first[l];
    const int x0 = (t % tx) * 32, y0 = (t / tx) * 32;
    const uint8_t *img = I.img[l];
    const float k0 = __uint_as_float(1032826801u), k1 = __uint_as_float(1040595070u), k2 = __uint_as_float(1044597305u), k3 = __uint_as_float(1046301408u);
    const float kk[7] = { k0, k1, k2, k3, k2, k1, k0 };
    for (int i = threadIdx.x; i < 38 * 32; i += 256) {
        const int r = i >> 5, c = i & 31;
        int y = y0 + r - 3, x = x0 + c;
        y = y < 0 ? -y : (y >= h ? 2 * h - 2 - y : y);
        if (h == 1) y = 0;
        float s = 0.f;
        if (x < w) {
            const uint8_t *row = img + (size_t)y * st;
            auto px = [&](int xx) { xx = xx < 0 ? -xx : (xx >= w ? 2 * w - 2 - xx : xx); return (float)row[w == 1 ? 0 : xx]; };
            s = __fmul_rn(kk[0], px(x - 3));
#pragma unroll
            for (int j = 1; j < 7; j++) s = __fmaf_rn(px(x + j - 3), kk[j], s);
        }
        s_row[r][c] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
        const int r = i >> 5, c = i & 31;
        const int y = y0 + r, x = x0 + c;
        if (x >= w || y >= h) continue;
        float s = __fmul_rn(kk[3], s_row[r + 3][c]);
#pragma unroll
        for (int j = 1; j <= 3; j++) s = __fmaf_rn(__fadd_rn(s_row[r + 3 + j][c], s_row[r + 3 - j][c]), kk[3 + j], s);
        const int v = __float2int_rn(s);
        blur[L.boff[l] + (size_t)y * L.pitch[l] + x] = (uint8_t)min(max(v, 0), 255);
    }
}

// k_describe.cu
void launch_describe_c(orbx_handle *h, int nlevels, const uint8_t *const *img, const int *step, const uint8_t *const *blur, const int *bstep,
                       const float *scale, const uint32_t *d_fxy, const float *d_fresp, const int32_t *d_fcount, int list_cap, int max_total,
                       orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_count);

static inline size_t al(size_t v, size_t a) { return (v + a - 1) / a * a; }

void orbx_cvorb_destroy(orbx_handle *h)
{
    orbx_cvorb *c = h->cv;
    if (!c) return;
    void *dev[] = { c->d_pyr, c->d_score, c->d_blur, c->d_xtab, c->d_ytab, c->d_cand, c->d_ncand, c->d_fxy, c->d_fresp, c->d_fcount };
    for (void *p : dev) if (p) cudaFree(p);
    delete c;
    h->cv = nullptr;
}

// level geometry of cv::ORB for one frame size; false: a level vanishes
static bool cv_geometry(const orbx_handle *h, int w, int hgt, CvLevels &L, std::vector<ResizeTab> *xt, std::vector<ResizeTab> *yt)
{
    memset(&L, 0, sizeof(L));
    const orbx_params &p = h->prm;
    L.nlevels = p.nlevels;
    const double sf = (double)p.scale_factor;
    // features per level: ORB_Impl::detectAndCompute -> computeKeyPoints (orb.cpp): float arithmetic, cvRound per level, remainder on the last
    const float factor = (float)(1.0 / sf);
    float nd = (float)p.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)p.nlevels));
    int sum = 0;
    size_t off = 0, boff = 0, coff = 0; int tiles = 0;
    for (int l = 0; l < p.nlevels; l++) {
        const float scale = (float)pow(sf, (double)l);               // getScale(level, firstLevel = 0, scaleFactor)
        L.scale[l] = scale;
        L.w[l] = (int)lrintf((float)w / scale); L.h[l] = (int)lrintf((float)hgt / scale);
        if (L.w[l] < 1 || L.h[l] < 1) return false;
        L.pitch[l] = (int)al((size_t)L.w[l], 128);
        if (l > 0) { L.off[l] = off; off += al((size_t)L.pitch[l] * L.h[l], 256); }
        L.boff[l] = boff; boff += al((size_t)L.pitch[l] * L.h[l], 256);
        if (l < p.nlevels - 1) { L.N[l] = (int)lrintf(nd); sum += L.N[l]; nd *= factor; }
        else L.N[l] = std::max(p.nfeatures - sum, 0);
        L.cand_cap[l] = (int)al((size_t)std::max(4096, L.w[l] * L.h[l] / 16), 64);
        L.cand_off[l] = coff; coff += (size_t)L.cand_cap[l];
        L.tile_first[l] = tiles;
        tiles += ((L.w[l] + CV_TILE_W - 1) / CV_TILE_W) * ((L.h[l] + CV_TILE_H - 1) / CV_TILE_H);
        if (l > 0 && xt && yt) {
            L.xtab[l] = (int)xt->size(); xt->resize(xt->size() + L.w[l]); exact_table(L.w[l - 1], L.w[l], xt->data() + L.xtab[l]);
            L.ytab[l] = (int)yt->size(); yt->resize(yt->size() + L.h[l]); exact_table(L.h[l - 1], L.h[l], yt->data() + L.ytab[l]);
        }
    }
    L.tile_first[p.nlevels] = tiles;
    return true;
}

orbx_status orbx_cvorb_create(orbx_handle *h)
{
    const orbx_params &p = h->prm;
    if (p.ini_th_fast < 1) { h->err = "profile C needs a FAST threshold >= 1"; return ORBX_E_INVALID; }
    orbx_cvorb *c = new orbx_cvorb();
    memset(c, 0, sizeof(*c));
    h->cv = c;
    CvLevels L;
    if (!cv_geometry(h, p.max_width, p.max_height, L, nullptr, nullptr)) { h->err = "unsupported max_width x max_height geometry (a pyramid level vanishes)"; return ORBX_E_UNSUPPORTED; }
    size_t pyr = 0, lvl = 0, cand = 0;
    for (int l = 0; l < p.nlevels; l++) { const size_t b = al((size_t)L.pitch[l] * L.h[l], 256) + 4096; if (l) pyr += b; lvl += b; cand += (size_t)L.cand_cap[l] + 4096; }
    c->pyr_cap = pyr + pyr / 8 + 4096; c->lvl_cap = lvl + lvl / 8 + 4096; c->cand_cap_total = cand + cand / 8;
    c->tab_cap = 2 * (p.max_width + p.max_height) * 6 + 1024;
    ORBX_CUDA(h, cudaMalloc(&c->d_pyr, c->pyr_cap));
    ORBX_CUDA(h, cudaMalloc(&c->d_score, c->lvl_cap));
    ORBX_CUDA(h, cudaMalloc(&c->d_blur, c->lvl_cap));
    ORBX_CUDA(h, cudaMalloc(&c->d_xtab, sizeof(ResizeTab) * c->tab_cap));
    ORBX_CUDA(h, cudaMalloc(&c->d_ytab, sizeof(ResizeTab) * c->tab_cap));
    ORBX_CUDA(h, cudaMalloc(&c->d_cand, c->cand_cap_total * sizeof(uint32_t)));
    ORBX_CUDA(h, cudaMalloc(&c->d_ncand, (ORBX_MAX_LEVELS * 2 + 8) * sizeof(int32_t)));
    ORBX_CUDA(h, cudaMalloc(&c->d_fxy, (size_t)ORBX_MAX_LEVELS * CV_LIST_CAP * sizeof(uint32_t)));
    ORBX_CUDA(h, cudaMalloc(&c->d_fresp, (size_t)ORBX_MAX_LEVELS * CV_LIST_CAP * sizeof(float)));
    ORBX_CUDA(h, cudaMalloc(&c->d_fcount, ORBX_MAX_LEVELS * sizeof(int32_t)));
    c->width = c->height = -1;
    return ORBX_OK;
}

static orbx_status cv_set_geometry(orbx_handle *h, int w, int hgt)
{
    orbx_cvorb *c = h->cv;
    if (c->width == w && c->height == hgt) return ORBX_OK;
    if (w > h->prm.max_width || hgt > h->prm.max_height) { h->err = "frame larger than max_width x max_height"; return ORBX_E_INVALID; }
    CvLevels L; std::vector<ResizeTab> xt, yt;
    if (!cv_geometry(h, w, hgt, L, &xt, &yt)) { h->err = "unsupported frame geometry (a pyramid level vanishes: cv::resize throws)"; return ORBX_E_UNSUPPORTED; }
    size_t pyr = 0, lvl = 0, cand = 0;
    for (int l = 0; l < L.nlevels; l++) { const size_t b = al((size_t)L.pitch[l] * L.h[l], 256); if (l) pyr += b; lvl += b; cand += (size_t)L.cand_cap[l]; }
    if (pyr > c->pyr_cap || lvl > c->lvl_cap || cand > c->cand_cap_total || (int)xt.size() > c->tab_cap || (int)yt.size() > c->tab_cap) { h->err = "frame geometry does not fit the arenas sized at create"; return ORBX_E_INVALID; }
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    if (!xt.empty()) ORBX_CUDA(h, cudaMemcpy(c->d_xtab, xt.data(), xt.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice));
    if (!yt.empty()) ORBX_CUDA(h, cudaMemcpy(c->d_ytab, yt.data(), yt.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice));
    c->L = L; c->width = w; c->height = hgt;
    return ORBX_OK;
}

// one frame: d_gray (rows of `step` bytes) -> keypoints / descriptors (level order; Harris descending inside a level) + count, all on the device
orbx_status orbx_cvorb_run(orbx_handle *h, const uint8_t *d_gray, int width, int height, size_t step,
                           orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_count)
{
    orbx_status st = cv_set_geometry(h, width, height);
    if (st != ORBX_OK) return st;
    orbx_cvorb *c = h->cv;
    const CvLevels &L = c->L;
    h->pdl_chain = true;
    CvImgs I, Bl;
    const uint8_t *imgs[ORBX_MAX_LEVELS], *blurs[ORBX_MAX_LEVELS]; int steps[ORBX_MAX_LEVELS], bsteps[ORBX_MAX_LEVELS];
    for (int l = 0; l < L.nlevels; l++) {
        I.img[l] = l ? c->d_pyr + L.off[l] : d_gray; I.step[l] = l ? L.pitch[l] : (int)step;
        Bl.img[l] = c->d_blur + L.boff[l]; Bl.step[l] = L.pitch[l];
        imgs[l] = I.img[l]; steps[l] = I.step[l]; blurs[l] = Bl.img[l]; bsteps[l] = Bl.step[l];
    }
    ORBX_CUDA(h, cudaMemsetAsync(c->d_ncand, 0, ORBX_MAX_LEVELS * sizeof(int32_t), h->stream));
    for (int l = 1; l < L.nlevels; l++) {
        ProfScope ps(h, ORBX_K_RESIZE);
        dim3 grid((L.w[l] + 63) / 64, (L.h[l] + 3) / 4);
        orbx_launch_pdl(h, k_resize_exact, grid, dim3(256), 0, h->stream, I.img[l - 1], I.step[l - 1], L.w[l - 1], L.h[l - 1],
                        (uint8_t *)(c->d_pyr + L.off[l]), L.pitch[l], L.w[l], L.h[l], (const ResizeTab *)(c->d_xtab + L.xtab[l]), (const ResizeTab *)(c->d_ytab + L.ytab[l]));
    }
    const int tiles = L.tile_first[L.nlevels];
    { ProfScope ps(h, ORBX_K_FAST); orbx_launch_pdl(h, k_cfast_score, dim3(tiles), dim3(CV_TILE_W * CV_TILE_H), 0, h->stream, L, I, c->d_score, h->prm.ini_th_fast); }
    { ProfScope ps(h, ORBX_K_FAST); orbx_launch_pdl(h, k_cfast_nms, dim3(tiles), dim3(CV_TILE_W * CV_TILE_H), 0, h->stream, L, (const uint8_t *)c->d_score, c->d_cand, c->d_ncand, h->d_status); }
    { ProfScope ps(h, ORBX_K_QUADTREE); orbx_launch_pdl(h, k_cretain, dim3(L.nlevels), dim3(1024), 0, h->stream, L, I, (const uint32_t *)c->d_cand, (const int32_t *)c->d_ncand,
                                                        c->d_fxy, c->d_fresp, c->d_fcount, h->d_status); }
    {
        // blur tiles of 32 x 32, flattened over the levels; the table rides in d_ncand's tail
        int bt[ORBX_MAX_LEVELS + 1], n = 0;
        for (int l = 0; l < L.nlevels; l++) { bt[l] = n; n += ((L.w[l] + 31) / 32) * ((L.h[l] + 31) / 32); }
        bt[L.nlevels] = n;
        int *d_bt = c->d_ncand + ORBX_MAX_LEVELS;
        ORBX_CUDA(h, cudaMemcpyAsync(d_bt, bt, (L.nlevels + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        ProfScope ps(h, ORBX_K_BLUR);
        k_cblur_f32<<<n, 256, 0, h->stream>>>(L, I, c->d_blur, d_bt);
    }
    int max_total = 0;
    for (int l = 0; l < L.nlevels; l++) max_total += CV_LIST_CAP;
    launch_describe_c(h, L.nlevels, imgs, steps, blurs, bsteps, L.scale, c->d_fxy, c->d_fresp, c->d_fcount, CV_LIST_CAP, std::min(max_total, cap + 1),
                      d_kps, d_desc, cap, d_count);
    ORBX_CUDA(h, cudaGetLastError());
    return ORBX_OK;
}

// stage access for the parity tests: pyramid level / blurred level of the last frame, tightly packed
orbx_status orbx_cvorb_get_level(orbx_handle *h, int level, int blurred, uint8_t *out, size_t out_step, const uint8_t *l0, size_t l0_step)
{
    orbx_cvorb *c = h->cv;
    if (!c || c->width < 0 || level < 0 || level >= c->L.nlevels) return ORBX_E_INVALID;
    ORBX_CUDA(h, cudaStreamSynchronize(h->stream));
    const uint8_t *src; size_t step;
    if (blurred) { src = c->d_blur + c->L.boff[level]; step = c->L.pitch[level]; }
    else if (level == 0) { src = l0; step = l0_step; }
    else { src = c->d_pyr + c->L.off[level]; step = c->L.pitch[level]; }
    ORBX_CUDA(h, cudaMemcpy2D(out, out_step, src, step, (size_t)c->L.w[level], (size_t)c->L.h[level], cudaMemcpyDeviceToHost));
    return ORBX_OK;
}
void orbx_cvorb_level_size(const orbx_handle *h, int w, int hgt, int level, int *lw, int *lh)
{
    const float scale = (float)pow((double)h->prm.scale_factor, (double)level);
    *lw = (int)lrintf((float)w / scale); *lh = (int)lrintf((float)hgt / scale);
}
