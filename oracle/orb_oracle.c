/* orb_oracle.c — CPU ORACLE (test infrastructure, NOT product code).  See orb_oracle.h.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/dynamic_visual_slam/).  OpenCV-internal arithmetic (resize, FAST, GaussianBlur,
 * fastAtan2, cvRound) is restated from SURVEY.md App. A and pinned against cv2 4.13.0 by the tests.
 * Build: gcc -O2 -ffp-contract=off -fopenmp  (NO -march=native, NO -ffast-math: fp32 steps must
 * round individually, as in the reference's x86-64 baseline build).
 */
#include "orb_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const int8_t k_pattern[1024] = {
#include "../include/orbx_pattern.inc"
};

enum { PATCH_SIZE = 31, HALF_PATCH_SIZE = 15, EDGE_THRESHOLD = 19 };   /* ORBextractor.cpp:71-73 */

/* cvRound: round-half-to-even (SSE cvtss2si under the default rounding mode) */
static inline int cv_round_f(float v)  { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }

float orc_cosf(float x) { return cosf(x); }
float orc_sinf(float x) { return sinf(x); }

/* The trig of computeOrbDescriptor (ORBextractor.cpp:112-113: `angle = kpt.angle * factorPI; a = cos(angle), b = sin(angle)` in float)
 * over a whole range of float bit patterns of kpt.angle: wrapping sums of the result bit patterns, with this machine's libm.  The CUDA
 * restatement of glibc's cosf / sinf is pinned against it over every fp32 angle in [0, 360] degrees (tests/test_gpu_parity.py). */
void orc_trig_checksum(uint32_t first, uint32_t last, uint64_t *sum_cos, uint64_t *sum_sin, int nthreads)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const long long total = (long long)last - (long long)first + 1;
    uint64_t sc = 0, ss = 0;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for reduction(+ : sc, ss) num_threads(nthreads) schedule(static)
    for (long long i = 0; i < total; i++) {
        union { uint32_t u; float f; } d, c, n;
        d.u = first + (uint32_t)i;
        const float rad = d.f * factorPI;
        c.f = cosf(rad); n.f = sinf(rad);
        sc += c.u; ss += n.u;
    }
    *sum_cos = sc; *sum_sin = ss;
}

/* ------------------------------------------------------------------------------------------ */
/* ORBextractor::ORBextractor — ORBextractor.cpp:409-469                                      */
int orc_extractor_init(orc_extractor *ex, int nfeatures, float scaleFactor, int nlevels,
                       int iniThFAST, int minThFAST)
{
    if (!ex || nlevels < 1 || nlevels > ORC_MAX_LEVELS || nfeatures < 1) return -1;
    memset(ex, 0, sizeof(*ex));
    ex->nfeatures = nfeatures; ex->nlevels = nlevels;
    ex->iniThFAST = iniThFAST; ex->minThFAST = minThFAST;
    ex->scaleFactor = (double)scaleFactor;
    ex->scale[0] = 1.0f; ex->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        ex->scale[i]  = (float)((double)ex->scale[i - 1] * ex->scaleFactor);   /* :418 */
        ex->sigma2[i] = ex->scale[i] * ex->scale[i];
    }
    for (int i = 0; i < nlevels; i++) {
        ex->inv_scale[i]  = 1.0f / ex->scale[i];
        ex->inv_sigma2[i] = 1.0f / ex->sigma2[i];
    }
    float factor = (float)(1.0f / ex->scaleFactor);                              /* :434 */
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int level = 0; level < nlevels - 1; level++) {
        ex->nfeat_level[level] = cv_round_f(nDesired);
        sum += ex->nfeat_level[level];
        nDesired *= factor;
    }
    ex->nfeat_level[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;

    /* umax — :451-468 */
    int v, v0;
    int vmax = (int)floorf(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceilf(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) ex->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (ex->umax[v0] == ex->umax[v0 + 1]) ++v0;
        ex->umax[v] = v0;
        ++v0;
    }
    return 0;
}

/* ComputePyramid level sizes — ORBextractor.cpp:1173-1174 */
void orc_level_size(const orc_extractor *ex, int w, int h, int level, int *lw, int *lh)
{
    float s = ex->inv_scale[level];
    *lw = cv_round_f((float)w * s);
    *lh = cv_round_f((float)h * s);
}

/* Where the reference itself is defined.  For small or very elongated images ORBextractor.cpp leaves defined behaviour:
 *   1  a level rounds to zero columns or rows: cv::resize throws on the empty dsize (:1174, :1182);
 *   2  a level has exactly 32 rows (maxBorderY - minBorderY == 0) or borders of opposite sign with |W/H| >= 0.5: the root count
 *      nIni (:559) is infinite or negative and vpIniNodes.resize(nIni) throws std::length_error (:566) — observed with the compiled
 *      reference (oracle/_ref; tests/test_oracle_vs_ref.py);
 *   3  a level holds FAST cells but W/H < 0.5 rounds to nIni = 0 roots: vpIniNodes[...] indexes an empty vector (:586).
 * 0 = supported: every division of the reference is finite where its result is used.  The CUDA library applies the same rule
 * (ORBX_E_UNSUPPORTED) so that the two fail on exactly the same inputs. */
int orc_geometry_status(const orc_extractor *ex, int w, int h)
{
    for (int l = 0; l < ex->nlevels; l++) {
        int lw, lh;
        orc_level_size(ex, w, h, l, &lw, &lh);
        if (lw <= 0 || lh <= 0) return 1;
        const int Wb = lw - 2 * (EDGE_THRESHOLD - 3), Hb = lh - 2 * (EDGE_THRESHOLD - 3);
        if (Hb == 0) return 2;
        const int nIni = (int)roundf((float)Wb / (float)Hb);
        if (nIni < 0) return 2;
        if (nIni == 0 && Wb >= 35 && Hb >= 35) return 3;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* cv::resize(..., INTER_LINEAR) on CV_8UC1 — called at ORBextractor.cpp:1182.  SURVEY App. A.1 */
void orc_resize_tables(int ssize, int dsize, int32_t *ofs, int16_t *coef, int horizontal)
{
    double inv_scale = (double)dsize / ssize;
    double scale = 1. / inv_scale;
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (horizontal) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        }
        ofs[d] = s;
        float c0 = 1.f - f, c1 = f;
        int a0 = cv_round_f(c0 * 2048.f), a1 = cv_round_f(c1 * 2048.f);
        coef[2 * d]     = (int16_t)(a0 > 32767 ? 32767 : a0 < -32768 ? -32768 : a0);
        coef[2 * d + 1] = (int16_t)(a1 > 32767 ? 32767 : a1 < -32768 ? -32768 : a1);
    }
}

static inline int clipi(int v, int lo, int hi) { return v < lo ? lo : (v >= hi ? hi - 1 : v); }

void orc_resize_linear(const uint8_t *src, int sw, int sh, size_t sstep,
                       uint8_t *dst, int dw, int dh, size_t dstep)
{
    int32_t *xofs = (int32_t *)malloc(sizeof(int32_t) * dw);
    int16_t *xa   = (int16_t *)malloc(sizeof(int16_t) * 2 * dw);
    int32_t *yofs = (int32_t *)malloc(sizeof(int32_t) * dh);
    int16_t *yb   = (int16_t *)malloc(sizeof(int16_t) * 2 * dh);
    int32_t *r0   = (int32_t *)malloc(sizeof(int32_t) * dw);
    int32_t *r1   = (int32_t *)malloc(sizeof(int32_t) * dw);
    orc_resize_tables(sw, dw, xofs, xa, 1);
    orc_resize_tables(sh, dh, yofs, yb, 0);
    int prev0 = -1000000, prev1 = -1000000;
    for (int y = 0; y < dh; y++) {
        int sy0 = clipi(yofs[y], 0, sh), sy1 = clipi(yofs[y] + 1, 0, sh);
        if (sy0 == prev1 && sy0 != prev0) {            /* reuse the lower row as the new upper row */
            int32_t *t = r0; r0 = r1; r1 = t; prev0 = sy0; prev1 = -1000000;
        }
        if (sy0 != prev0) {
            const uint8_t *S = src + (size_t)sy0 * sstep;
            for (int x = 0; x < dw; x++) {
                int sx = xofs[x], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
                r0[x] = S[sx] * xa[2 * x] + S[sx1] * xa[2 * x + 1];
            }
            prev0 = sy0;
        }
        if (sy1 != prev1) {
            const uint8_t *S = src + (size_t)sy1 * sstep;
            for (int x = 0; x < dw; x++) {
                int sx = xofs[x], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
                r1[x] = S[sx] * xa[2 * x] + S[sx1] * xa[2 * x + 1];
            }
            prev1 = sy1;
        }
        int b0 = yb[2 * y], b1 = yb[2 * y + 1];
        uint8_t *D = dst + (size_t)y * dstep;
        for (int x = 0; x < dw; x++) {
            int v = (((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2;
            D[x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(xofs); free(xa); free(yofs); free(yb); free(r0); free(r1);
}

/* ------------------------------------------------------------------------------------------ */
/* cv::FAST(roi, kps, threshold, nonmaxSuppression=true), TYPE_9_16 — called at
 * ORBextractor.cpp:826-827 and :845-846.  SURVEY App. A.2.  Output in raster order, coordinates
 * relative to the ROI, score = (max arc-min) - 1.                                              */
static const int k_ring_dx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
static const int k_ring_dy[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

/* S = max over the 16 arcs of 9 contiguous ring pixels of min(center - ring) resp. min(ring - center) */
static inline int fast_arc_score(const uint8_t *p, const int *off)
{
    int d[25];
    int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[off[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = -256;
    for (int k = 0; k < 16; k++) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; j++) { int t = d[k + j]; if (t < mn) mn = t; if (t > mx) mx = t; }
        if (mn > best) best = mn;            /* darker-ring arc:  center - ring all > t  */
        if (-mx > best) best = -mx;          /* brighter-ring arc: ring - center all > t */
    }
    return best;
}

int orc_fast_roi(const uint8_t *img, size_t step, int w, int h, int threshold,
                 orc_cand *out, int cap)
{
    if (w < 7 || h < 7) return 0;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = k_ring_dy[k] * (int)step + k_ring_dx[k];
    uint8_t *sc = (uint8_t *)calloc((size_t)w * h, 1);
    for (int y = 3; y < h - 3; y++) {
        const uint8_t *row = img + (size_t)y * step;
        for (int x = 3; x < w - 3; x++) {
            const uint8_t *p = row + x;
            int v = p[0], hi = v + threshold, lo = v - threshold;
            /* exact necessary condition: every opposite pair must hold one arc member */
            int a = p[off[0]], b = p[off[8]];
            int br = (a > hi) | (b > hi), dk = (a < lo) | (b < lo);
            if (!(br | dk)) continue;
            a = p[off[4]]; b = p[off[12]];
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            if (!(br | dk)) continue;
            a = p[off[2]]; b = p[off[10]];
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            if (!(br | dk)) continue;
            a = p[off[6]]; b = p[off[14]];
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            if (!(br | dk)) continue;
            int s = fast_arc_score(p, off);
            if (s > threshold) sc[(size_t)y * w + x] = (uint8_t)(s - 1);
        }
    }
    int n = 0;
    for (int y = 3; y < h - 3; y++) {
        const uint8_t *c = sc + (size_t)y * w, *u = c - w, *d = c + w;
        for (int x = 3; x < w - 3; x++) {
            int s = c[x];
            if (!s) continue;
            if (s > c[x - 1] && s > c[x + 1] && s > u[x - 1] && s > u[x] && s > u[x + 1] &&
                s > d[x - 1] && s > d[x] && s > d[x + 1]) {
                if (n < cap) { out[n].x = x; out[n].y = y; out[n].score = s; }
                n++;
            }
        }
    }
    free(sc);
    return n;
}

/* the per-cell FAST loop of ComputeKeyPointsOctTree — ORBextractor.cpp:785-872.
 * Output coordinates are relative to (minBorderX, minBorderY), in the reference's push order. */
int orc_fast_cells(const orc_extractor *ex, const uint8_t *img, size_t step, int w, int h,
                   orc_cand *out, int cap)
{
    const float W = 35;
    const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
    const int maxBorderX = w - EDGE_THRESHOLD + 3, maxBorderY = h - EDGE_THRESHOLD + 3;
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols < 1 || nRows < 1) return 0;
    const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
    int n = 0;
    int cellcap = (wCell + 6) * (hCell + 6);
    orc_cand *cell = (orc_cand *)malloc(sizeof(orc_cand) * cellcap);
    for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBorderY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBorderY - 3) continue;
        if (maxY > maxBorderY) maxY = (float)maxBorderY;
        for (int j = 0; j < nCols; j++) {
            const float iniX = (float)(minBorderX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBorderX - 6) continue;
            if (maxX > maxBorderX) maxX = (float)maxBorderX;
            int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
            const uint8_t *roi = img + (size_t)y0 * step + x0;
            int m = orc_fast_roi(roi, step, x1 - x0, y1 - y0, ex->iniThFAST, cell, cellcap);
            if (m == 0) m = orc_fast_roi(roi, step, x1 - x0, y1 - y0, ex->minThFAST, cell, cellcap);
            for (int k = 0; k < m; k++) {
                if (n < cap) {
                    out[n].x = cell[k].x + j * wCell;
                    out[n].y = cell[k].y + i * hCell;
                    out[n].score = cell[k].score;
                }
                n++;
            }
        }
    }
    free(cell);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* libstdc++ std::sort (introsort, _S_threshold = 16) restated for the one call at
 * ORBextractor.cpp:700 with compareNodes (:538-553): ascending (count, UL.x); ties stay in
 * whatever order this exact algorithm leaves them.                                             */
typedef struct { int32_t cnt, ulx, payload; } srt_t;
static inline int srt_less(const srt_t *a, const srt_t *b)
{
    if (a->cnt < b->cnt) return 1;
    if (a->cnt > b->cnt) return 0;
    return a->ulx < b->ulx;
}
static inline void srt_swap(srt_t *a, srt_t *b) { srt_t t = *a; *a = *b; *b = t; }

static void srt_adjust_heap(srt_t *first, long hole, long len, srt_t value)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (srt_less(first + child, first + (child - 1))) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    long parent = (hole - 1) / 2;
    while (hole > top && srt_less(first + parent, &value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void srt_heapsort(srt_t *first, srt_t *last)
{
    long len = last - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            srt_t v = first[parent];
            srt_adjust_heap(first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        srt_t v = *last;
        *last = *first;
        srt_adjust_heap(first, 0, last - first, v);
    }
}
static void srt_median_to_first(srt_t *result, srt_t *a, srt_t *b, srt_t *c)
{
    if (srt_less(a, b)) {
        if (srt_less(b, c)) srt_swap(result, b);
        else if (srt_less(a, c)) srt_swap(result, c);
        else srt_swap(result, a);
    } else if (srt_less(a, c)) srt_swap(result, a);
    else if (srt_less(b, c)) srt_swap(result, c);
    else srt_swap(result, b);
}
static srt_t *srt_partition(srt_t *first, srt_t *last, srt_t *pivot)
{
    for (;;) {
        while (srt_less(first, pivot)) ++first;
        --last;
        while (srt_less(pivot, last)) --last;
        if (!(first < last)) return first;
        srt_swap(first, last);
        ++first;
    }
}
static void srt_introsort_loop(srt_t *first, srt_t *last, long depth)
{
    while (last - first > 16) {
        if (depth == 0) { srt_heapsort(first, last); return; }
        --depth;
        srt_t *mid = first + (last - first) / 2;
        srt_median_to_first(first, first + 1, mid, last - 1);
        srt_t *cut = srt_partition(first + 1, last, first);
        srt_introsort_loop(cut, last, depth);
        last = cut;
    }
}
static void srt_unguarded_linear_insert(srt_t *last)
{
    srt_t val = *last;
    srt_t *next = last - 1;
    while (srt_less(&val, next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void srt_insertion_sort(srt_t *first, srt_t *last)
{
    if (first == last) return;
    for (srt_t *i = first + 1; i != last; ++i) {
        if (srt_less(i, first)) {
            srt_t val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(srt_t));
            *first = val;
        } else srt_unguarded_linear_insert(i);
    }
}
static void srt_sort(srt_t *first, srt_t *last)
{
    if (first == last) return;
    long n = last - first, lg = 0;
    while ((n >> (lg + 1)) > 0) lg++;
    srt_introsort_loop(first, last, lg * 2);
    if (last - first > 16) {
        srt_insertion_sort(first, first + 16);
        for (srt_t *i = first + 16; i != last; ++i) srt_unguarded_linear_insert(i);
    } else srt_insertion_sort(first, last);
}
void orc_introsort_pairs(int32_t *cnt, int32_t *ulx, int32_t *payload, int n)
{
    srt_t *a = (srt_t *)malloc(sizeof(srt_t) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) { a[i].cnt = cnt[i]; a[i].ulx = ulx[i]; a[i].payload = payload[i]; }
    srt_sort(a, a + n);
    for (int i = 0; i < n; i++) { cnt[i] = a[i].cnt; ulx[i] = a[i].ulx; payload[i] = a[i].payload; }
    free(a);
}

/* ------------------------------------------------------------------------------------------ */
/* Frontend feature culling for the backend — reference frontend.cpp:1168-1218.
 * out = the query index of every match, in match order (:1181-1190), then the unmatched keypoints (index order, :1193-1198) sorted with
 * std::sort and the comparator a.first > b.first on (response, index) pairs (:1201-1202) — equal responses stay where libstdc++'s
 * introsort leaves them — taken while fewer than max_new were added and response >= min_response (:1209-1210).                     */
static uint32_t cull_key(float response)
{
    union { float f; uint32_t u; } v;
    v.f = response + 0.0f;                                  /* -0 -> +0: the float comparator does not tell them apart */
    const uint32_t asc = v.u ^ ((v.u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
    return ~asc;                                            /* ascending key = descending response */
}
int orc_cull_keyframe(const float *response, int n, const int32_t *match_query, int n_matches, int max_new, float min_response, int32_t *out_index)
{
    unsigned char *matched = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    int m = 0;
    for (int i = 0; i < n_matches; i++) {
        const int q = match_query[i];
        if (q < 0 || q >= n) { free(matched); return -1; }
        matched[q] = 1;
        out_index[m++] = q;
    }
    srt_t *a = (srt_t *)malloc(sizeof(srt_t) * (size_t)(n > 0 ? n : 1));
    int nu = 0;
    for (int i = 0; i < n; i++)
        if (!matched[i]) { a[nu].cnt = (int32_t)(cull_key(response[i]) ^ 0x80000000u); a[nu].ulx = 0; a[nu].payload = i; nu++; }
    srt_sort(a, a + nu);
    int added = 0;
    for (int i = 0; i < nu; i++) {
        if (added >= max_new || response[a[i].payload] < min_response) break;
        out_index[m++] = a[i].payload;
        added++;
    }
    free(a); free(matched);
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* ExtractorNode / DivideNode / DistributeOctTree — ORBextractor.hpp:31-42, .cpp:480-536, 555-779.
 * Literal std::list emulation (doubly linked), push_front / erase exactly as the reference.    */
typedef struct qnode {
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    int *keys; int nkeys;
    int bNoMore;
    struct qnode *prev, *next;
} qnode;
typedef struct { qnode *head, *tail; int size; } qlist;

static qnode *qnode_new(int reserve)
{
    qnode *n = (qnode *)calloc(1, sizeof(qnode));
    n->keys = (int *)malloc(sizeof(int) * (reserve > 0 ? reserve : 1));
    return n;
}
static void qlist_push_front(qlist *l, qnode *n)
{
    n->prev = NULL; n->next = l->head;
    if (l->head) l->head->prev = n; else l->tail = n;
    l->head = n; l->size++;
}
static void qlist_push_back(qlist *l, qnode *n)
{
    n->next = NULL; n->prev = l->tail;
    if (l->tail) l->tail->next = n; else l->head = n;
    l->tail = n; l->size++;
}
static qnode *qlist_erase(qlist *l, qnode *n)   /* returns the following element */
{
    qnode *nx = n->next;
    if (n->prev) n->prev->next = n->next; else l->head = n->next;
    if (n->next) n->next->prev = n->prev; else l->tail = n->prev;
    l->size--;
    free(n->keys); free(n);
    return nx;
}

static void divide_node(const qnode *p, const orc_cand *c, qnode *n[4])
{
    const int halfX = (int)ceilf((float)(p->URx - p->ULx) / 2);
    const int halfY = (int)ceilf((float)(p->BRy - p->ULy) / 2);
    for (int k = 0; k < 4; k++) n[k] = qnode_new(p->nkeys);
    n[0]->ULx = p->ULx;          n[0]->ULy = p->ULy;
    n[0]->URx = p->ULx + halfX;  n[0]->URy = p->ULy;
    n[0]->BLx = p->ULx;          n[0]->BLy = p->ULy + halfY;
    n[0]->BRx = p->ULx + halfX;  n[0]->BRy = p->ULy + halfY;
    n[1]->ULx = n[0]->URx; n[1]->ULy = n[0]->URy;
    n[1]->URx = p->URx;    n[1]->URy = p->URy;
    n[1]->BLx = n[0]->BRx; n[1]->BLy = n[0]->BRy;
    n[1]->BRx = p->URx;    n[1]->BRy = p->ULy + halfY;
    n[2]->ULx = n[0]->BLx; n[2]->ULy = n[0]->BLy;
    n[2]->URx = n[0]->BRx; n[2]->URy = n[0]->BRy;
    n[2]->BLx = p->BLx;    n[2]->BLy = p->BLy;
    n[2]->BRx = n[0]->BRx; n[2]->BRy = p->BLy;
    n[3]->ULx = n[2]->URx; n[3]->ULy = n[2]->URy;
    n[3]->URx = n[1]->BRx; n[3]->URy = n[1]->BRy;
    n[3]->BLx = n[2]->BRx; n[3]->BLy = n[2]->BRy;
    n[3]->BRx = p->BRx;    n[3]->BRy = p->BRy;
    for (int i = 0; i < p->nkeys; i++) {
        int id = p->keys[i];
        float kx = (float)c[id].x, ky = (float)c[id].y;
        qnode *dst;
        if (kx < (float)n[0]->URx) dst = (ky < (float)n[0]->BRy) ? n[0] : n[2];
        else                        dst = (ky < (float)n[0]->BRy) ? n[1] : n[3];
        dst->keys[dst->nkeys++] = id;
    }
    for (int k = 0; k < 4; k++) if (n[k]->nkeys == 1) n[k]->bNoMore = 1;
}

int orc_distribute_octtree(const orc_cand *cands, int n, int minX, int maxX, int minY, int maxY,
                           int N, orc_cand *out, int cap)
{
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));       /* :559 */
    if (nIni < 1) return -3;                 /* the reference divides by zero here */
    const float hX = (float)(maxX - minX) / nIni;
    qlist L = { 0, 0, 0 };
    qnode **ini = (qnode **)malloc(sizeof(qnode *) * nIni);
    for (int i = 0; i < nIni; i++) {
        qnode *ni = qnode_new(n);
        ni->ULx = (int)(hX * (float)i);       ni->ULy = 0;
        ni->URx = (int)(hX * (float)(i + 1)); ni->URy = 0;
        ni->BLx = ni->ULx; ni->BLy = maxY - minY;
        ni->BRx = ni->URx; ni->BRy = maxY - minY;
        qlist_push_back(&L, ni);
        ini[i] = ni;
    }
    for (int i = 0; i < n; i++) {
        int r = (int)((float)cands[i].x / hX);                                /* :584 */
        if (r >= nIni) r = nIni - 1;         /* (out-of-bounds in the reference; unreachable) */
        ini[r]->keys[ini[r]->nkeys++] = i;
    }
    free(ini);
    for (qnode *lit = L.head; lit;) {                                         /* :589-600 */
        if (lit->nkeys == 1) { lit->bNoMore = 1; lit = lit->next; }
        else if (lit->nkeys == 0) lit = qlist_erase(&L, lit);
        else lit = lit->next;
    }
    int bFinish = 0;
    int vcap = 4 * (L.size > 0 ? L.size : 1) + 16, vn = 0;
    srt_t *vSize = (srt_t *)malloc(sizeof(srt_t) * vcap);
    qnode **vPtr = (qnode **)malloc(sizeof(qnode *) * vcap);
#define VPUSH(node) do { if (vn == vcap) { vcap *= 2; vSize = (srt_t *)realloc(vSize, sizeof(srt_t) * vcap); \
        vPtr = (qnode **)realloc(vPtr, sizeof(qnode *) * vcap); } \
        vSize[vn].cnt = (node)->nkeys; vSize[vn].ulx = (node)->ULx; vSize[vn].payload = vn; vPtr[vn] = (node); vn++; } while (0)
    while (!bFinish) {
        int prevSize = L.size;
        int nToExpand = 0;
        vn = 0;
        for (qnode *lit = L.head; lit;) {                                     /* :620-678 */
            if (lit->bNoMore) { lit = lit->next; continue; }
            qnode *ch[4];
            divide_node(lit, cands, ch);
            for (int k = 0; k < 4; k++) {
                if (ch[k]->nkeys > 0) {
                    qlist_push_front(&L, ch[k]);
                    if (ch[k]->nkeys > 1) { nToExpand++; VPUSH(ch[k]); }
                } else { free(ch[k]->keys); free(ch[k]); }
            }
            lit = qlist_erase(&L, lit);
        }
        if (L.size >= N || L.size == prevSize) bFinish = 1;                   /* :682-685 */
        else if (L.size + nToExpand * 3 > N) {                                /* :686 */
            while (!bFinish) {
                prevSize = L.size;
                int pn = vn;
                srt_t *pv = (srt_t *)malloc(sizeof(srt_t) * (pn > 0 ? pn : 1));
                qnode **pp = (qnode **)malloc(sizeof(qnode *) * (pn > 0 ? pn : 1));
                memcpy(pv, vSize, sizeof(srt_t) * pn);
                memcpy(pp, vPtr, sizeof(qnode *) * pn);
                vn = 0;
                srt_sort(pv, pv + pn);                                        /* :700 */
                for (int j = pn - 1; j >= 0; j--) {
                    qnode *node = pp[pv[j].payload];
                    qnode *ch[4];
                    divide_node(node, cands, ch);
                    for (int k = 0; k < 4; k++) {
                        if (ch[k]->nkeys > 0) {
                            qlist_push_front(&L, ch[k]);
                            if (ch[k]->nkeys > 1) VPUSH(ch[k]);
                        } else { free(ch[k]->keys); free(ch[k]); }
                    }
                    qlist_erase(&L, node);
                    if (L.size >= N) break;
                }
                free(pv); free(pp);
                if (L.size >= N || L.size == prevSize) bFinish = 1;
            }
        }
    }
#undef VPUSH
    free(vSize); free(vPtr);
    int m = 0;
    for (qnode *lit = L.head; lit;) {                                         /* :757-776 */
        int best = lit->keys[0];
        int maxResponse = cands[best].score;
        for (int k = 1; k < lit->nkeys; k++) {
            if (cands[lit->keys[k]].score > maxResponse) { best = lit->keys[k]; maxResponse = cands[best].score; }
        }
        if (m < cap) out[m] = cands[best];
        m++;
        qnode *nx = lit->next;
        free(lit->keys); free(lit);
        lit = nx;
    }
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* cv::fastAtan2 (scalar path) — called at ORBextractor.cpp:102.  SURVEY App. A.4              */
float orc_fast_atan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* IC_Angle — ORBextractor.cpp:76-103 */
float orc_ic_angle(const uint8_t *img, size_t step_, int cx, int cy, const int *umax)
{
    int m_01 = 0, m_10 = 0;
    const int step = (int)step_;
    const uint8_t *center = img + (size_t)cy * step_ + cx;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
}

/* GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on a continuous CV_8UC1 — ORBextractor.cpp:1132-1133.
 * 8.8 fixed-point taps, one rounding after the column pass.                    */
static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
    return p;
}
void orc_gaussian_blur7(const uint8_t *src, int w, int h, size_t sstep, uint8_t *dst, size_t dstep)
{
    static const int K[7] = { 18, 34, 48, 56, 48, 34, 18 };
    uint16_t *rows = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t *S = src + (size_t)y * sstep;
        uint16_t *R = rows + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            int acc = 0;
            if (x >= 3 && x < w - 3) for (int i = 0; i < 7; i++) acc += K[i] * S[x + i - 3];
            else for (int i = 0; i < 7; i++) acc += K[i] * S[reflect101(x + i - 3, w)];
            R[x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t *r[7];
        for (int j = 0; j < 7; j++) r[j] = rows + (size_t)reflect101(y + j - 3, h) * w;
        uint8_t *D = dst + (size_t)y * dstep;
        for (int x = 0; x < w; x++) {
            uint32_t c = 0;
            for (int j = 0; j < 7; j++) c += (uint32_t)K[j] * r[j][x];
            D[x] = (uint8_t)((c + 32768u) >> 16);
        }
    }
    free(rows);
}

/* cv::findFundamentalMat(..., FM_RANSAC, thresh, conf) scores a model with FMEstimatorCallback::computeError (calib3d/src/fundam.cpp):
 * symmetric epipolar distance in double, stored as float; inlier <=> err <= (float)(thresh*thresh) — reference call sites
 * frontend.cpp:1134-1154, :625-645.  F row-major 3x3, pts1 = the first point set (x2' F x1 = 0).  Returns the inlier count. */
int orc_fmat_inliers(const float *p1, const float *p2, int n, const double *F, double thresh, uint8_t *mask)
{
    const float t = (float)(thresh * thresh);
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        const double x1 = p1[2 * i], y1 = p1[2 * i + 1], x2 = p2[2 * i], y2 = p2[2 * i + 1];
        double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
        const double s2 = 1. / (a * a + b * b), d2 = x2 * a + y2 * b + c;
        a = F[0] * x2 + F[3] * y2 + F[6]; b = F[1] * x2 + F[4] * y2 + F[7]; c = F[2] * x2 + F[5] * y2 + F[8];
        const double s1 = 1. / (a * a + b * b), d1 = x1 * a + y1 * b + c;
        const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
        const float err = (float)(e1 > e2 ? e1 : e2);
        const int in = err <= t;
        if (mask) mask[i] = (uint8_t)in;
        cnt += in;
    }
    return cnt;
}

/* Frontend::estimateCameraPose, frontend.cpp:858-892: the 3D-2D correspondences handed to cv::solvePnPRansac.  For every match, in match order:
 * the previous frame's keypoint (trainIdx) is back-projected with the depth at its rounded pixel (std::round on floats: half away from zero),
 * kept iff the pixel lies inside the depth image and 0.3 < d <= 3.0 m; X = (u - cx) * d / fx in float (rgb_fx_ ... are float members, :278).
 * kps as 28-byte records (x, y first), matches as (queryIdx, trainIdx, imgIdx, distance).  Returns the number of correspondences. */
int orc_pnp_points(const orc_keypoint *prev_kps, const orc_keypoint *curr_kps, const orc_dmatch *m, int nm, const uint16_t *depth, int w, int h,
                   size_t dstep, float fx, float fy, float cx, float cy, float *p3, float *p2)
{
    int n = 0;
    for (int i = 0; i < nm; i++) {
        const float px = prev_kps[m[i].trainIdx].x, py = prev_kps[m[i].trainIdx].y;
        const int xp = (int)roundf(px), yp = (int)roundf(py);
        if (xp < 0 || yp < 0 || xp >= w || yp >= h) continue;
        const float d = (float)*(const uint16_t *)((const uint8_t *)depth + (size_t)yp * dstep + 2 * (size_t)xp) * 0.001f;
        if (d <= 0.3f || d > 3.0f) continue;
        p3[3 * n] = (px - cx) * d / fx; p3[3 * n + 1] = (py - cy) * d / fy; p3[3 * n + 2] = d;
        p2[2 * n] = curr_kps[m[i].queryIdx].x; p2[2 * n + 1] = curr_kps[m[i].queryIdx].y;
        n++;
    }
    return n;
}

/* cv::solvePnPRansac(points3d, points2d, K, dist, rvec, tvec, false, 100, 4.0, 0.99, inliers) (frontend.cpp:911-923) scores a pose with
 * PnPRansacCallback::computeError (calib3d/src/solvepnp.cpp): cv::projectPoints in double (x = R0 X + R1 Y + R2 Z + t0 ...; x /= z via z = 1/z;
 * u = x fx + cx — the distortion terms vanish for the zero coefficients of a rectified stream), projections stored as float, squared pixel
 * distance accumulated in float; inlier <=> err <= (float)(thresh * thresh).  R row-major 3x3 (= cv::Rodrigues(rvec)), t 3.  Pinned against
 * cv2.projectPoints in tests/test_pnp.py.  Returns the inlier count. */
int orc_pnp_inliers(const float *p3, const float *p2, int n, const double *R, const double *t, double fx, double fy, double cx, double cy,
                    double thresh, uint8_t *mask)
{
    const float t2 = (float)(thresh * thresh);
    int cnt = 0;
    for (int i = 0; i < n; i++) {
        const double X = p3[3 * i], Y = p3[3 * i + 1], Z = p3[3 * i + 2];
        double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
        double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
        double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
        z = z ? 1. / z : 1;
        x *= z; y *= z;
        const float u = (float)(x * fx + cx), v = (float)(y * fy + cy);
        const float dx = p2[2 * i] - u, dy = p2[2 * i + 1] - v;
        float e = dx * dx;
        e += dy * dy;
        const int in = e <= t2;
        if (mask) mask[i] = (uint8_t)in;
        cnt += in;
    }
    return cnt;
}

/* ---- profile C (cv::ORB, the reference's gtest test/test_dbow2_integration.cpp:19,38 and BASELINE configs[0]) primitives ---- */
/* cv::resize(..., INTER_LINEAR_EXACT) on CV_8UC1: ufixedpoint16 coefficients (8 fractional bits) from the double-precision source
 * coordinate, horizontal pass exact in 16 bits, vertical pass in 32 bits with ONE rounding (+32768 >> 16).  Pinned bit for bit against
 * cv2.resize(INTER_LINEAR_EXACT) in tests/test_cvorb_oracle.py. */
void orc_resize_exact_tables(int ssize, int dsize, int32_t *ofs, int16_t *coef /*2 per i, c0 + c1 = 256*/)
{
    const double scale = 1.0 / ((double)dsize / (double)ssize);
    for (int x = 0; x < dsize; x++) {
        const double fval = scale * ((double)x + 0.5) - 0.5;
        int ival = (int)floor(fval);
        int c0, c1;
        if (ival >= 0 && ssize > 1) {
            if (ival < ssize - 1) { c1 = (int)lrint((fval - (double)ival) * 256.0); c0 = 256 - c1; }
            else { ival = ssize - 2; c0 = 0; c1 = 256; }
        } else { ival = 0; c0 = 256; c1 = 0; }
        ofs[x] = ival; coef[2 * x] = (int16_t)c0; coef[2 * x + 1] = (int16_t)c1;
    }
}
void orc_resize_linear_exact(const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw, int dh, size_t dstep)
{
    int32_t *xofs = (int32_t *)malloc(sizeof(int32_t) * dw), *yofs = (int32_t *)malloc(sizeof(int32_t) * dh);
    int16_t *xa = (int16_t *)malloc(sizeof(int16_t) * 2 * dw), *ya = (int16_t *)malloc(sizeof(int16_t) * 2 * dh);
    orc_resize_exact_tables(sw, dw, xofs, xa);
    orc_resize_exact_tables(sh, dh, yofs, ya);
    for (int y = 0; y < dh; y++) {
        const uint8_t *S0 = src + (size_t)yofs[y] * sstep, *S1 = src + (size_t)(yofs[y] + 1 < sh ? yofs[y] + 1 : sh - 1) * sstep;
        uint8_t *D = dst + (size_t)y * dstep;
        for (int x = 0; x < dw; x++) {
            const int x0 = xofs[x], x1 = x0 + 1 < sw ? x0 + 1 : sw - 1;
            const uint32_t h0 = (uint32_t)S0[x0] * xa[2 * x] + (uint32_t)S0[x1] * xa[2 * x + 1];
            const uint32_t h1 = (uint32_t)S1[x0] * xa[2 * x] + (uint32_t)S1[x1] * xa[2 * x + 1];
            const uint32_t v = (h0 * (uint32_t)ya[2 * y] + h1 * (uint32_t)ya[2 * y + 1] + 32768u) >> 16;
            D[x] = (uint8_t)(v > 255u ? 255u : v);
        }
    }
    free(xofs); free(yofs); free(xa); free(ya);
}
/* cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on a SUB-MATRIX (cv::ORB blurs its pyramid layers in place inside one big buffer):
 * OpenCV then runs sepFilter2D with the CV_32F kernel getGaussianKernel(7, 2).  fp32 arithmetic of the AVX2/FMA build of cv2 4.13.0
 * (found by matching cv2.ORB descriptors, 5000 of 5000 rows): row pass = k0*p0, then fused multiply-adds left to right; column pass =
 * k3*r3, then fused multiply-adds of the symmetric pairs (r[3+i] + r[3-i]) * k[3+i]; one rounding to u8 (round half to even).
 * A scalar OpenCV build differs in 1-3 px per 120k, hence north_star's ">= 99.9 % of descriptor rows" bar for profile C. */
static const uint32_t k_gauss7_f32_bits[7] = { 1032826801u, 1040595070u, 1044597305u, 1046301408u, 1044597305u, 1040595070u, 1032826801u };
void orc_gaussian_blur7_f32(const uint8_t *src, int w, int h, size_t sstep, uint8_t *dst, size_t dstep)
{
    float k[7];
    memcpy(k, k_gauss7_f32_bits, sizeof(k));
    float *rows = (float *)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t *S = src + (size_t)y * sstep;
        for (int x = 0; x < w; x++) {
            float s = k[0] * (float)S[reflect101(x - 3, w)];
            for (int i = 1; i < 7; i++) s = fmaf((float)S[reflect101(x + i - 3, w)], k[i], s);
            rows[(size_t)y * w + x] = s;
        }
    }
    for (int y = 0; y < h; y++) {
        const float *r[7];
        for (int j = 0; j < 7; j++) r[j] = rows + (size_t)reflect101(y + j - 3, h) * w;
        uint8_t *D = dst + (size_t)y * dstep;
        for (int x = 0; x < w; x++) {
            float s = k[3] * r[3][x];
            for (int i = 1; i <= 3; i++) s = fmaf(r[3 + i][x] + r[3 - i][x], k[3 + i], s);
            long v = lrintf(s);
            D[x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(rows);
}

/* computeOrbDescriptor — ORBextractor.cpp:106-146 */
void orc_descriptor(const uint8_t *blur, size_t step_, int cx, int cy, float angle_deg, uint8_t *desc)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = angle_deg * factorPI;
    float a = cosf(angle), b = sinf(angle);
    const uint8_t *center = blur + (size_t)cy * step_ + cx;
    const int step = (int)step_;
    const int8_t *pat = k_pattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int j = 0; j < 8; j++) {
            float x0 = (float)pat[4 * j], y0 = (float)pat[4 * j + 1];
            float x1 = (float)pat[4 * j + 2], y1 = (float)pat[4 * j + 3];
            int t0 = center[cv_round_f(x0 * b + y0 * a) * step + cv_round_f(x0 * a - y0 * b)];
            int t1 = center[cv_round_f(x1 * b + y1 * a) * step + cv_round_f(x1 * a - y1 * b)];
            val |= (t0 < t1) << j;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* ORBextractor::operator() — ORBextractor.cpp:1086-1167 (ComputePyramid :1169-1194,
 * ComputeKeyPointsOctTree :781-896).  The 19-px REFLECT_101 border is never read on this path
 *  and is not materialised.                                                   */
int orc_extract(const orc_extractor *ex, const uint8_t *gray, int w, int h, size_t step,
                orc_keypoint *kps, uint8_t *desc, int cap, orc_trace *tr)
{
    if (!gray || w <= 0 || h <= 0) return -1;                                  /* :1090-1091 */
    if (orc_geometry_status(ex, w, h) != 0) return -3;                         /* the reference throws or faults here */
    const int nl = ex->nlevels;
    int lw[ORC_MAX_LEVELS] = {0}, lh[ORC_MAX_LEVELS] = {0};
    uint8_t *pyr[ORC_MAX_LEVELS];
    size_t total = 0;
    for (int l = 0; l < nl; l++) { orc_level_size(ex, w, h, l, &lw[l], &lh[l]); total += (size_t)lw[l] * lh[l]; }
    uint8_t *buf = (uint8_t *)malloc(total);
    uint8_t *blur = (uint8_t *)malloc((size_t)lw[0] * lh[0]);
    size_t off = 0;
    for (int l = 0; l < nl; l++) { pyr[l] = buf + off; off += (size_t)lw[l] * lh[l]; }
    for (int y = 0; y < h; y++) memcpy(pyr[0] + (size_t)y * lw[0], gray + (size_t)y * step, (size_t)w);
    for (int l = 1; l < nl; l++)                                               /* :1182 */
        orc_resize_linear(pyr[l - 1], lw[l - 1], lh[l - 1], (size_t)lw[l - 1], pyr[l], lw[l], lh[l], (size_t)lw[l]);
    if (tr) for (int l = 0; l < nl; l++) { tr->lw[l] = lw[l]; tr->lh[l] = lh[l]; }
    if (tr && tr->pyramid) memcpy(tr->pyramid, buf, total);

    int ccap = ex->nfeatures * 10 > 65536 ? ex->nfeatures * 10 : 65536;
    if (ccap < w * h / 4 + 1024) ccap = w * h / 4 + 1024;                     /* noise frames: ~10 % of the pixels are FAST keypoints */
    orc_cand *cands = (orc_cand *)malloc(sizeof(orc_cand) * ccap);
    orc_cand *sel = (orc_cand *)malloc(sizeof(orc_cand) * ccap);
    int n_out = 0, overflow = 0;
    size_t boff = 0;
    /* the reference runs selection+orientation for all levels first, then blur+descriptors per
     * level; the stages are independent across levels, so one loop gives identical results. */
    for (int l = 0; l < nl; l++) {
        const int minBX = EDGE_THRESHOLD - 3, minBY = minBX;
        const int maxBX = lw[l] - EDGE_THRESHOLD + 3, maxBY = lh[l] - EDGE_THRESHOLD + 3;
        int nc = 0, ns = 0;
        if (maxBX - minBX >= 35 && maxBY - minBY >= 35) {
            nc = orc_fast_cells(ex, pyr[l], (size_t)lw[l], lw[l], lh[l], cands, ccap);
            if (nc > ccap) { overflow = 1; nc = ccap; }
            ns = orc_distribute_octtree(cands, nc, minBX, maxBX, minBY, maxBY, ex->nfeat_level[l], sel, ccap);
            if (ns < 0) ns = 0;
        }
        if (tr) {
            tr->ncands[l] = nc; tr->nkeys[l] = ns;
            if (tr->cands) memcpy(tr->cands + (size_t)l * tr->cand_cap, cands,
                                  sizeof(orc_cand) * (nc < tr->cand_cap ? nc : tr->cand_cap));
        }
        int need_blur = ns > 0 || (tr && tr->blurred);
        if (need_blur) orc_gaussian_blur7(pyr[l], lw[l], lh[l], (size_t)lw[l], blur, (size_t)lw[l]);   /* :1132-1133 */
        if (tr && tr->blurred) memcpy(tr->blurred + boff, blur, (size_t)lw[l] * lh[l]);
        boff += (size_t)lw[l] * lh[l];
        const int scaledPatchSize = (int)(PATCH_SIZE * ex->scale[l]);          /* :880 */
        const float scale = ex->scale[l];
        for (int k = 0; k < ns; k++) {
            if (n_out >= cap) { overflow = 1; break; }
            orc_keypoint kp;
            kp.x = (float)sel[k].x + (float)minBX;                             /* :886-887 */
            kp.y = (float)sel[k].y + (float)minBY;
            kp.size = (float)scaledPatchSize;
            kp.response = (float)sel[k].score;
            kp.octave = l; kp.class_id = -1;
            int cx = cv_round_f(kp.x), cy = cv_round_f(kp.y);
            kp.angle = orc_ic_angle(pyr[l], (size_t)lw[l], cx, cy, ex->umax);   /* :893-895 */
            orc_descriptor(blur, (size_t)lw[l], cx, cy, kp.angle, desc + (size_t)n_out * 32);   /* :1138 */
            if (l != 0) { kp.x *= scale; kp.y *= scale; }                      /* :1147-1149 */
            kps[n_out++] = kp;
        }
    }
    free(cands); free(sel); free(buf); free(blur);
    return overflow ? -2 : n_out;
}

int orc_extract_batch(const orc_extractor *ex, const uint8_t *gray, int nframes, int w, int h,
                      orc_keypoint *kps, uint8_t *desc, int cap, int32_t *counts, int nthreads)
{
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    #pragma omp parallel for schedule(dynamic, 1) reduction(|:bad)
    for (int f = 0; f < nframes; f++) {
        int n = orc_extract(ex, gray + (size_t)f * w * h, w, h, (size_t)w,
                            kps + (size_t)f * cap, desc + (size_t)f * cap * 32, cap, NULL);
        counts[f] = n;
        if (n < 0) bad = 1;
    }
    return bad ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Frontend::isValidDepth / filterDepth — frontend.cpp:457-473, 503-527 */
int orc_filter_depth(const orc_keypoint *kps, const uint8_t *desc, int n,
                     const uint16_t *depth, int dw, int dh, size_t dstep,
                     float min_depth, float max_depth,
                     orc_keypoint *okps, uint8_t *odesc, int32_t *orig_idx)
{
    int m = 0;
    for (int i = 0; i < n; i++) {
        int x = (int)roundf(kps[i].x), y = (int)roundf(kps[i].y);
        if (x < 0 || y < 0 || x >= dw || y >= dh) continue;
        float d = depth[(size_t)y * dstep + x] * 0.001f;
        if (d < min_depth || d > max_depth || isnan(d) || isinf(d) || d < 0.0f) continue;
        okps[m] = kps[i];
        if (desc && odesc) memcpy(odesc + (size_t)m * 32, desc + (size_t)i * 32, 32);
        if (orig_idx) orig_idx[m] = i;
        m++;
    }
    return m;
}

/* Backend::categorizeObservation — backend.cpp:1011-1029 (float pixel vs double box, inclusive) */
int orc_categorize(float px, float py, const orc_box *boxes, int nboxes)
{
    for (int i = 0; i < nboxes; i++) {
        const orc_box *b = &boxes[i];
        if ((double)px >= b->cx - b->w / 2 && (double)px <= b->cx + b->w / 2 &&
            (double)py >= b->cy - b->h / 2 && (double)py <= b->cy + b->h / 2)
            return b->class_id;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* cv::BFMatcher(NORM_HAMMING) — frontend.cpp:220,1123,614 ; backend.cpp:222,1072.  App. A.8   */
int orc_hamming(const uint8_t *a, const uint8_t *b)
{
    uint64_t x[4], y[4];
    memcpy(x, a, 32); memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
           __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

int orc_knn2(const uint8_t *q, int nq, const uint8_t *t, int nt, orc_dmatch *out, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; i++) {
        int d0 = 1 << 30, d1 = 1 << 30, i0 = -1, i1 = -1;
        const uint8_t *qi = q + (size_t)i * 32;
        for (int j = 0; j < nt; j++) {
            int d = orc_hamming(qi, t + (size_t)j * 32);
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
            else if (d < d1) { d1 = d; i1 = j; }
        }
        out[2 * i].queryIdx = i; out[2 * i].trainIdx = i0; out[2 * i].imgIdx = 0;
        out[2 * i].distance = i0 >= 0 ? (float)d0 : 0.f;
        out[2 * i + 1].queryIdx = i; out[2 * i + 1].trainIdx = i1; out[2 * i + 1].imgIdx = 0;
        out[2 * i + 1].distance = i1 >= 0 ? (float)d1 : 0.f;
    }
    return nq;
}

int orc_match(const uint8_t *q, int nq, const uint8_t *t, int nt, orc_dmatch *out, int nthreads)
{
    if (nt <= 0) return 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; i++) {
        int d0 = 1 << 30, i0 = -1;
        const uint8_t *qi = q + (size_t)i * 32;
        for (int j = 0; j < nt; j++) {
            int d = orc_hamming(qi, t + (size_t)j * 32);
            if (d < d0) { d0 = d; i0 = j; }
        }
        out[i].queryIdx = i; out[i].trainIdx = i0; out[i].imgIdx = 0; out[i].distance = (float)d0;
    }
    return nq;
}

/* ------------------------------------------------------------------------------------------ */
/* Seeded synthetic inputs: integer-only, world-anchored so frame f is the world
 * shifted by (4f, 2f) px — consecutive frames overlap and true matches exist.                  */
static inline uint32_t syn_hash(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t h = a * 0x9E3779B1u ^ b * 0x85EBCA77u ^ c * 0xC2B2AE3Du;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}
static inline uint32_t syn_vnoise(uint32_t X, uint32_t Y, int shift, uint32_t seed)
{
    uint32_t P = 1u << shift, gx = X >> shift, gy = Y >> shift, fx = X & (P - 1), fy = Y & (P - 1);
    uint32_t v00 = syn_hash(gx, gy, seed) & 255, v10 = syn_hash(gx + 1, gy, seed) & 255;
    uint32_t v01 = syn_hash(gx, gy + 1, seed) & 255, v11 = syn_hash(gx + 1, gy + 1, seed) & 255;
    uint32_t top = v00 * (P - fx) + v10 * fx, bot = v01 * (P - fx) + v11 * fx;
    return (top * (P - fy) + bot * fy) >> (2 * shift);
}
static inline uint8_t syn_gray_px(uint32_t seed, uint32_t X, uint32_t Y)
{
    uint32_t acc = 8 * syn_vnoise(X, Y, 6, seed + 1) + 4 * syn_vnoise(X, Y, 5, seed + 2) +
                   2 * syn_vnoise(X, Y, 4, seed + 3) + 2 * syn_vnoise(X, Y, 3, seed + 4);
    int base = 40 + (int)((acc >> 4) * 5 >> 3);
    int val = base;
    uint32_t cx0 = (X >> 5) - 1, cy0 = (Y >> 5) - 1;
    for (uint32_t dy = 0; dy < 2; dy++)
        for (uint32_t dx = 0; dx < 2; dx++) {
            uint32_t cx = cx0 + dx, cy = cy0 + dy;
            uint32_t hsh = syn_hash(cx, cy, seed ^ 0xABCD1234u);
            if ((hsh & 3u) == 0) continue;
            uint32_t x0 = (cx << 5) + ((hsh >> 2) & 31), y0 = (cy << 5) + ((hsh >> 7) & 31);
            uint32_t rw = 5 + ((hsh >> 12) & 31) % 27, rh = 5 + ((hsh >> 17) & 31) % 27;
            if (X - x0 < rw && Y - y0 < rh) {
                val = (int)(hsh >> 24) + ((base - 128) >> 2);
                val = val < 0 ? 0 : val > 255 ? 255 : val;
            }
        }
    return (uint8_t)val;
}
void orc_synth_gray(uint32_t seed, int frame, int w, int h, uint8_t *out, size_t step)
{
    const uint32_t OX = (1u << 20) + 4u * (uint32_t)frame, OY = (1u << 20) + 2u * (uint32_t)frame;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            out[(size_t)y * step + x] = syn_gray_px(seed, OX + (uint32_t)x, OY + (uint32_t)y);
}
void orc_synth_depth(uint32_t seed, int frame, int w, int h, uint16_t *out, size_t step)
{
    const uint32_t OX = (1u << 20) + 4u * (uint32_t)frame, OY = (1u << 20) + 2u * (uint32_t)frame;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint32_t X = OX + (uint32_t)x, Y = OY + (uint32_t)y;
            uint32_t d = 600 + 12 * syn_vnoise(X, Y, 6, seed ^ 0x0D0D0D0Du);
            if (syn_hash(X >> 4, Y >> 4, seed ^ 0xDDDD0001u) % 100u < 15u) d = 0;
            out[(size_t)y * step + x] = (uint16_t)d;
        }
}
/* YOLO-style dynamic-object boxes for BASELINE configs[4] (four per frame, ~15 % of the area, class "person"):
 * centre on a quarter-pixel grid, size 15-25 % of the frame per axis; box 2 is of another class (1) so that the first-containing-
 * box rule (backend.cpp:1015-1026) matters where it overlaps a "person" box (class 0).  Returns the number written (4). */
int orc_synth_boxes(uint32_t seed, int frame, int w, int h, orc_box *out)
{
    for (int b = 0; b < 4; b++) {
        uint32_t h0 = syn_hash((uint32_t)frame, (uint32_t)b, seed ^ 0xB0B0B0B0u), h1 = syn_hash((uint32_t)frame, (uint32_t)b, seed ^ 0x0B0B0B0Bu);
        out[b].cx = (double)(h0 % (uint32_t)(4 * w)) / 4.0;
        out[b].cy = (double)((h0 >> 13) % (uint32_t)(4 * h)) / 4.0;
        out[b].w = (double)w * (double)(150 + h1 % 101) / 1000.0;
        out[b].h = (double)h * (double)(150 + (h1 >> 11) % 101) / 1000.0;
        out[b].class_id = b == 2 ? 1 : 0;
        out[b].pad = 0;
    }
    return 4;
}
/* Backend::categorizeObservation + filtered_objects_ over a keypoint list (backend.cpp:1011-1029, 746-751): stable compaction of the
 * keypoints whose first containing box is not of a class in drop_mask */
int orc_filter_boxes(const orc_keypoint *kps, const uint8_t *desc, int n, const orc_box *boxes, int nboxes, uint64_t drop_mask,
                     orc_keypoint *okps, uint8_t *odesc)
{
    int m = 0;
    for (int i = 0; i < n; i++) {
        const int c = orc_categorize(kps[i].x, kps[i].y, boxes, nboxes);
        if (c >= 0 && c < 64 && ((drop_mask >> c) & 1ull)) continue;
        okps[m] = kps[i];
        memcpy(odesc + (size_t)m * 32, desc + (size_t)i * 32, 32);
        m++;
    }
    return m;
}
/* landmark-database rows: counter-based, 32 B per row, uniform bytes */
void orc_synth_descriptors(uint32_t seed, uint64_t first_row, int nrows, uint8_t *out)
{
    for (int r = 0; r < nrows; r++) {
        uint64_t row = first_row + (uint64_t)r;
        for (int k = 0; k < 8; k++) {
            uint32_t v = syn_hash((uint32_t)row, (uint32_t)(row >> 32) * 8u + (uint32_t)k, seed);
            memcpy(out + (size_t)r * 32 + 4 * k, &v, 4);
        }
    }
}

/* ---- Backend::reprojectPoint (reference backend.cpp:1153-1173) ----
 * point_camera = R.t() * (point_world - t) in double; cv::Mat's 3x3 * 3x1 product accumulates (a0*b0 + a1*b1) + a2*b2 without FMA
 * (pinned against cv2.gemm in tests/test_oracle_vs_cv2.py); z <= 0 gives (-1, -1); u = (float)(fx*x/z + cx).  R row-major. */
void orc_reproject(const float *p, const double *R, const double *t, double fx, double fy, double cx, double cy, float *uv)
{
    const double d0 = (double)p[0] - t[0], d1 = (double)p[1] - t[1], d2 = (double)p[2] - t[2];
    double pc[3];
    for (int i = 0; i < 3; i++) pc[i] = (R[0 * 3 + i] * d0 + R[1 * 3 + i] * d1) + R[2 * 3 + i] * d2;
    if (pc[2] <= 0) { uv[0] = -1.f; uv[1] = -1.f; return; }
    uv[0] = (float)(fx * pc[0] / pc[2] + cx);
    uv[1] = (float)(fy * pc[1] / pc[2] + cy);
}

/* ---- Backend::associateObservation (reference backend.cpp:1064-1120) for a batch of observations of one category ----
 * candidates: every landmark with Hamming distance < max_desc (:1068-1077); among them the one with the smallest reprojection
 * error cv::norm(obs.pixel - reprojection) (double), if that error is < max_reproj (:1091-1111).  The reference walks an
 * unordered_map, so the winner among EXACTLY equal errors is unspecified there; here the lowest landmark row wins.
 * out_idx[i] = landmark row or -1, out_err[i] = its error (DBL_MAX if none), out_dist[i] = its Hamming distance. */
void orc_associate(const uint8_t *q, const float *qpx, int nq, const uint8_t *rows, const float *pos, int nrows,
                   const double *R, const double *t, double fx, double fy, double cx, double cy,
                   double max_desc, double max_reproj, int32_t *out_idx, double *out_err, float *out_dist, int nthreads)
{
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
    for (int i = 0; i < nq; i++) {
        int best = -1; double best_err = DBL_MAX; float best_d = 0.f;
        for (int j = 0; j < nrows; j++) {
            const float d = (float)orc_hamming(q + (size_t)i * 32, rows + (size_t)j * 32);
            if (!((double)d < max_desc)) continue;
            float uv[2];
            orc_reproject(pos + (size_t)j * 3, R, t, fx, fy, cx, cy, uv);
            const float ex = qpx[2 * i] - uv[0], ey = qpx[2 * i + 1] - uv[1];
            const double err = sqrt((double)ex * ex + (double)ey * ey);
            if (err < max_reproj && err < best_err) { best = j; best_err = err; best_d = d; }
        }
        out_idx[i] = best; out_err[i] = best_err; out_dist[i] = best_d;
    }
}

/* cv::cvtColor(BGR2GRAY) on CV_8UC3 (reference frontend.cpp:1084): (3735*B + 19235*G + 9798*R + 16384) >> 15 */
void orc_bgr2gray(const uint8_t *bgr, int w, int h, size_t sstep, uint8_t *gray, size_t dstep)
{
    for (int y = 0; y < h; y++) {
        const uint8_t *s = bgr + (size_t)y * sstep; uint8_t *d = gray + (size_t)y * dstep;
        for (int x = 0; x < w; x++) d[x] = (uint8_t)((3735 * s[3 * x] + 19235 * s[3 * x + 1] + 9798 * s[3 * x + 2] + 16384) >> 15);
    }
}

/* ---- Frontend::publishKeyframe, landmark / observation packing (reference frontend.cpp:731-776) ----
 * per keypoint i: x = round(pt.x), y = round(pt.y) (half away from zero); d = depth_u16(y, x) * 0.001f; camera point
 * ((pt.x - cx) * d / fx, (pt.y - cy) * d / fy, d) in float (rgb_fx_ .. are float members, :278); kept iff z > 0.3 && z < 3.0
 * (float z against double literals); world = R * p + t in double (cv::Mat: ((r0*p0 + r1*p1) + r2*p2) + t, no FMA).
 * Record = Landmark {id = i, position} + Observation {id = i, pixel_x, pixel_y (double from float), descriptor[32]}.
 * The reference indexes the depth image unchecked (the keypoints it passes were depth-filtered); pixels outside the image are skipped here. */
int orc_pack_keyframe(const orc_keypoint *kps, const uint8_t *desc, int n, const uint16_t *depth, int dw, int dh, size_t dstep_elems,
                      float fx, float fy, float cx, float cy, const double *R, const double *t, orc_kfrecord *out)
{
    int m = 0;
    for (int i = 0; i < n; i++) {
        const float px = kps[i].x, py = kps[i].y;
        const int x = (int)roundf(px), y = (int)roundf(py);
        if (x < 0 || y < 0 || x >= dw || y >= dh) continue;
        const float d = (float)depth[(size_t)y * dstep_elems + x] * 0.001f;
        const float X = (px - cx) * d / fx, Y = (py - cy) * d / fy, Z = d;
        if (!((double)Z > 0.3 && (double)Z < 3.0)) continue;
        orc_kfrecord *r = &out[m++];
        r->landmark_id = (uint64_t)i;
        for (int k = 0; k < 3; k++)
            r->position[k] = ((R[3 * k] * (double)X + R[3 * k + 1] * (double)Y) + R[3 * k + 2] * (double)Z) + t[k];
        r->pixel_x = (double)px; r->pixel_y = (double)py;
        memcpy(r->descriptor, desc + (size_t)i * 32, 32);
    }
    return m;
}

/* Harris corner response as cv::ORB computes it (OpenCV features2d orb.cpp HarrisResponses; the reference's ORBextractor.hpp:48 names
 * HARRIS_SCORE but never computes it): blockSize x blockSize sums of Ix^2, Iy^2, IxIy with
 * Ix = 2(I[x+1]-I[x-1]) + (I[y-1][x+1]-I[y-1][x-1]) + (I[y+1][x+1]-I[y+1][x-1]) (Iy alike), integer; response in fp32 =
 * (a*b - c*c - k*(a+b)^2) * (1/(4*blockSize*255))^4.  Pinned against cv2.ORB_create(nlevels=1) responses (exact). */
float orc_harris_response(const uint8_t *img, size_t step, int x0, int y0, int blockSize, float k)
{
    const int r = blockSize / 2;
    int a = 0, b = 0, c = 0;
    for (int i = 0; i < blockSize; i++) for (int j = 0; j < blockSize; j++) {
        const uint8_t *p = img + (size_t)(y0 - r + i) * step + (x0 - r + j);
        const int st = (int)step;
        const int Ix = (p[1] - p[-1]) * 2 + (p[-st + 1] - p[-st - 1]) + (p[st + 1] - p[st - 1]);
        const int Iy = (p[st] - p[-st]) * 2 + (p[st - 1] - p[-st - 1]) + (p[st + 1] - p[-st + 1]);
        a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
    }
    const float scale = 1.f / ((1 << 2) * blockSize * 255.f);
    const float s4 = scale * scale * scale * scale;
    return ((float)a * b - (float)c * c - k * ((float)a + b) * ((float)a + b)) * s4;
}
