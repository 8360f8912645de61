// k_describe.cu — intensity-centroid orientation + rBRIEF-256 + keypoint assembly, one warp per keypoint.
// Replaces IC_Angle / computeOrientation (reference ORBextractor.cpp:76-103, 471-478), computeOrbDescriptor /
// computeDescriptors (:106-146, 1077-1084), the tail of ComputeKeyPointsOctTree (:874-890) and the
// per-level assembly loop of operator() (:1123-1164).  Two variants: k_describe_fused (default) also evaluates the 7x7 Gaussian of
// :1132-1133 itself, only at the pixels the descriptor reads, from a staged window of the un-blurred level; k_describe reads a
// blurred pyramid written by k_blur7 (ORBX_OPT_FUSED_BLUR = 0).  Same bits either way.
//
// Floating point on this stage must round exactly like the reference's x86-64 build:
//   * cv::fastAtan2: fp32 polynomial, every operation rounded on its own (no FMA)   — SURVEY App. A.4
//   * cosf/sinf: glibc's algorithm (double-precision range reduction + polynomial, rounded once to
//     fp32) restated below; pinned against glibc over every float32 angle in [0,360] degrees
//   * x*b + y*a: fp32 multiply, fp32 add, cvRound = round-half-even                  — SURVEY App. A.5
#include "orbx_internal.h"
#include <cstring>

// read with 128-bit __ldg per lane (lane i owns pairs 8i..8i+7): a __constant__ table would serialise the
// 32 distinct addresses of a warp
__device__ __align__(16) const signed char c_pattern[1024] = {
#include "../../include/orbx_pattern.inc"
};
__device__ __constant__ int c_umax[16];

// per patch row v = lane - 15 of IC_Angle's disc: byte masks of the 31 columns u = -15 .. 15 (byte 4 i + j of the row = column 4 i + j - 15),
// 0xFF where |u| <= umax[|v|] (ORBextractor.cpp:76-103); read with two 128-bit __ldg per lane by the fused kernel
__device__ __align__(16) uint32_t g_ic_mask[32][8];

void upload_umax(const int *umax)
{
    cudaMemcpyToSymbol(c_umax, umax, sizeof(int) * 16);
    uint32_t m[32][8];
    memset(m, 0, sizeof(m));
    for (int v = -ORBX_HALF_PATCH; v <= ORBX_HALF_PATCH; v++)
        for (int u = -ORBX_HALF_PATCH; u <= ORBX_HALF_PATCH; u++)
            if ((u < 0 ? -u : u) <= umax[v < 0 ? -v : v]) m[v + ORBX_HALF_PATCH][(u + ORBX_HALF_PATCH) >> 2] |= 0xFFu << (8 * ((u + ORBX_HALF_PATCH) & 3));
    cudaMemcpyToSymbol(g_ic_mask, m, sizeof(m));
}

// ---- glibc sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h algorithm, |x| < 120) ----
__device__ __forceinline__ float glibc_sincos_poly(double x, double x2, int neg, int n)
{
    // neg selects the table entry that computes -cos for free
    if ((n & 1) == 0) {
        const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
        const double x3 = __dmul_rn(x, x2);
        const double s1 = __dadd_rn(S2, __dmul_rn(x2, S3));
        const double x7 = __dmul_rn(x3, x2);
        const double s = __dadd_rn(x, __dmul_rn(x3, S1));
        return __double2float_rn(__dadd_rn(s, __dmul_rn(x7, s1)));
    } else {
        const double sg = neg ? -1.0 : 1.0;
        const double C0 = sg * 0x1p0, C1 = sg * -0x1.ffffffd0c621cp-2, C2 = sg * 0x1.55553e1068f19p-5;
        const double C3 = sg * -0x1.6c087e89a359dp-10, C4 = sg * 0x1.99343027bf8c3p-16;
        const double x4 = __dmul_rn(x2, x2);
        const double c2 = __dadd_rn(C3, __dmul_rn(x2, C4));
        const double c1 = __dadd_rn(C0, __dmul_rn(x2, C1));
        const double x6 = __dmul_rn(x4, x2);
        const double c = __dadd_rn(c1, __dmul_rn(x4, C2));
        return __double2float_rn(__dadd_rn(c, __dmul_rn(x6, c2)));
    }
}
__device__ __forceinline__ unsigned abstop12(float f) { return (__float_as_uint(f) >> 20) & 0x7ff; }

__device__ float glibc_cosf(float y)
{
    double x = (double)y;
    if (abstop12(y) < 0x3f4u) {                       // |y| < ~pi/4 (top-12-bit compare, as glibc)
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return glibc_sincos_poly(x, __dmul_rn(x, x), 0, 1);
    }
    const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __dadd_rn(x, -__dmul_rn((double)n, 0x1.921FB54442D18p0));
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return glibc_sincos_poly(__dmul_rn(x, s), __dmul_rn(x, x), (n & 2) ? 1 : 0, n ^ 1);
}
__device__ float glibc_sinf(float y)
{
    double x = (double)y;
    if (abstop12(y) < 0x3f4u) {
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return glibc_sincos_poly(x, __dmul_rn(x, x), 0, 0);
    }
    const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __dadd_rn(x, -__dmul_rn((double)n, 0x1.921FB54442D18p0));
    const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return glibc_sincos_poly(__dmul_rn(x, s), __dmul_rn(x, x), (n & 2) ? 1 : 0, n);
}

// ---- cv::fastAtan2, scalar path ----
__device__ float cv_fast_atan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale);
    const float p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

struct DescParams {
    const uint8_t *l0; size_t l0_step, l0_fstride;
    const uint8_t *pyr; size_t pyr_slab;
    const uint8_t *blur; size_t blur_slab;
    const uint32_t *sel; int sel_slab;
    const int32_t *nsel;
    orbx_keypoint *kps; uint8_t *desc; int cap;     // per-frame capacity
    int32_t *counts;
    int32_t *status;
    const int32_t *map; int map_slab;               // fused kernel, filter-first order (k_keep_list): output slot -> index in the selected list; counts[] is then an input
};

#define DESC_WARPS 4

__global__ void __launch_bounds__(DESC_WARPS * 32) k_describe(DescParams P, const FrameGeom *__restrict__ G)
{
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int nl = G->nlevels;
    int gidx = blockIdx.x * DESC_WARPS + (threadIdx.x >> 5);
    // locate (level, index in level): levels are concatenated in order (ORBextractor.cpp:1123).  Lane l holds level l's count; an
    // inclusive warp scan gives the level boundaries, a ballot the level that contains keypoint gidx.
    int level, k, total;
    {
        const int c = lane < nl ? P.nsel[f * nl + lane] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < ORBX_MAX_LEVELS; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        total = __shfl_sync(0xffffffffu, incl, nl - 1);
        level = __popc(__ballot_sync(0xffffffffu, lane < nl && incl <= gidx));
        const int before = __shfl_sync(0xffffffffu, incl, max(level - 1, 0));
        k = gidx - (level > 0 ? before : 0);
        if (level >= nl) level = -1;
    }
    if (gidx == 0 && lane == 0) {
        P.counts[f] = total <= P.cap ? total : 0;
        if (total > P.cap) atomicOr(P.status, ORBX_DS_KP_OVERFLOW);
    }
    if (level < 0 || total > P.cap) return;
    const LevelGeom &g = G->lv[level];
    const uint32_t c = P.sel[(size_t)f * P.sel_slab + g.sel_off + k];
    // pt += (minBorderX, minBorderY) — ORBextractor.cpp:886-887; integer-valued, cvRound is the identity
    const int cx = orbx_px(c) + ORBX_BORDER, cy = orbx_py(c) + ORBX_BORDER;
    const uint8_t *img; size_t step;
    if (level == 0) { img = P.l0 + (size_t)f * P.l0_fstride; step = P.l0_step; }
    else { img = P.pyr + (size_t)f * P.pyr_slab + g.off; step = (size_t)g.pitch; }

    const uint8_t *bctr = P.blur + (size_t)f * P.blur_slab + g.boff + (size_t)cy * g.bpitch + cx;
    const int bstep = g.bpitch;

    // ---- IC_Angle: lane = column u+15, loop rows v ----
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - ORBX_HALF_PATCH;
        const int au = u < 0 ? -u : u;
        const uint8_t *ctr = img + (size_t)cy * step + cx + u;
        // all 31 row loads are issued before the first use (the stage is load-latency-bound: 4 in flight measured 81 % long-scoreboard stalls)
        int px[2 * ORBX_HALF_PATCH + 1];
#pragma unroll
        for (int v = -ORBX_HALF_PATCH; v <= ORBX_HALF_PATCH; v++) {
            const int av = v < 0 ? -v : v;
            px[v + ORBX_HALF_PATCH] = au <= c_umax[av] ? (int)__ldg(ctr + (ptrdiff_t)v * (ptrdiff_t)step) : 0;
        }
        int colsum = 0;
#pragma unroll
        for (int v = -ORBX_HALF_PATCH; v <= ORBX_HALF_PATCH; v++) { colsum += px[v + ORBX_HALF_PATCH]; m01 += v * px[v + ORBX_HALF_PATCH]; }
        m10 = u * colsum;
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = cv_fast_atan2((float)m01, (float)m10);

    // ---- rBRIEF: lane i computes descriptor byte i ----
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float ang = __fmul_rn(angle, factorPI);
    const float a = glibc_cosf(ang), b = glibc_sinf(ang);
    signed char pat[32];
    {
        const int4 *pp = reinterpret_cast<const int4 *>(c_pattern + lane * 32);
        *reinterpret_cast<int4 *>(pat) = __ldg(pp);
        *reinterpret_cast<int4 *>(pat + 16) = __ldg(pp + 1);
    }
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float x0 = (float)pat[4 * j], y0 = (float)pat[4 * j + 1], x1 = (float)pat[4 * j + 2], y1 = (float)pat[4 * j + 3];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = __ldg(bctr + r0 * bstep + c0), t1 = __ldg(bctr + r1 * bstep + c1);
        val |= (t0 < t1) << j;
    }
    uint8_t *drow = P.desc + ((size_t)f * P.cap + gidx) * ORBX_DESC_BYTES;
    drow[lane] = (uint8_t)val;
    if (lane == 0) {
        orbx_keypoint kp;
        kp.x = (float)cx; kp.y = (float)cy;
        if (level != 0) { kp.x = __fmul_rn(kp.x, g.scale); kp.y = __fmul_rn(kp.y, g.scale); }   // :1147-1149
        kp.size = g.size; kp.angle = angle; kp.response = (float)orbx_ps(c);
        kp.octave = level; kp.class_id = -1;
        P.kps[(size_t)f * P.cap + gidx] = kp;
    }
}

// ---- fused variant: the 7x7 Gaussian of ORBextractor.cpp:1132-1133 evaluated ONLY where a descriptor reads it ----
// The reference blurs every level (2.85 Mpx per 1280x720 frame) and then reads 512 blurred pixels around each of ~1000 keypoints.
// The rotated pattern stays within 18 px of the keypoint (max |cvRound(rotated coordinate)| over all angles), so the 43 x 43 window
// of the UN-blurred level around it holds every input of those 512 values; the same window holds the radius-15 disc of IC_Angle.
// Per keypoint (one warp): the window is staged in shared memory with 16-byte loads (rows outside the level are fetched from their
// BORDER_REFLECT_101 mirror, the at most two outside columns per side are mirrored in place — keypoints lie in [19, w-19), App. A.6);
// row pass of the whole window (2 x IDP.4A per pixel, exact u16 row sums) into a second buffer; then each of the 512 samples is
// one 7-tap column pass on that buffer: (sum + 32768) >> 16 — bit for bit the value cv::GaussianBlur leaves at that pixel.
// Against blur + describe as two kernels this is ~3x fewer instructions and no blurred pyramid in HBM (no 2 x 2.85 MB per frame
// written and re-read, no scattered DRAM gathers).  k_blur7 stays: orbx_get_blurred_level materialises levels with it on demand.
#ifndef DESC_STRIDE
#define DESC_STRIDE 3                // batches: rows per warp of the fused kernel (grid-stride walk)
#endif
#define DF_R 21
#define DF_ROWS (2 * DF_R + 1)       // 43
#define DF_PITCH 76                  // window tile pitch: 15 alignment bytes + 43 columns fit 64; 19 words: lanes on consecutive rows (IC_Angle) hit 32
                                     // different banks, and so do lanes on consecutive ROW PAIRS (row pass: 38 words = 6 banks apart, times 16 pairs =
                                     // the 16 even banks, the neighbour group on the odd ones; six pairs x five groups for the last rows)
#define DF_RCOLS 40                  // row-pass columns: ten 4-pixel groups starting at the word that holds column x-18
#define DF_CP 46                     // row sums are kept COLUMN-major (u16 [column][row], 46 entries per column: 43 rows, row 43 of the last row pair,
                                     // and an even count so that a row pair is one aligned 32-bit store): the seven taps of a sample's column pass
                                     // are then four consecutive 32-bit words = 4 LDS + 4 IDP.2A per sample
#define DF_WARP_BYTES (16 + DF_ROWS * DF_PITCH + 16 + ((DF_RCOLS * DF_CP * 2 + 16 + 15) & ~15))

// row sums of the four pixels of word C (L, R = the words on either side): taps 18,34,48,56,48,34,18 on bytes x-3 .. x+3.  The taps are laid out per
// output byte against the three words as they are (ten IDP.4A, no funnel shifts): byte 0 needs L and C, bytes 1 and 2 all three, byte 3 C and R.
__device__ __forceinline__ void df_hpass4(uint32_t L, uint32_t C, uint32_t R, uint32_t &h0, uint32_t &h1, uint32_t &h2, uint32_t &h3)
{
    h0 = __dp4a(L, 0x30221200u, __dp4a(C, 0x12223038u, 0u));
    h1 = __dp4a(L, 0x22120000u, __dp4a(C, 0x22303830u, __dp4a(R, 0x00000012u, 0u)));
    h2 = __dp4a(L, 0x12000000u, __dp4a(C, 0x30383022u, __dp4a(R, 0x00001222u, 0u)));
    h3 = __dp4a(C, 0x38302212u, __dp4a(R, 0x00122230u, 0u));                    // row sums <= 255 * 256 fit 16 bits
}

// blurred value from the column-major row sums: taps 18,34,48,56,48,34,18 on entries e .. e+6; the four words that hold them start at
// entry e & ~1, so the tap weights sit one half-word later when e is odd
__device__ __forceinline__ uint32_t df_sample(const uint32_t *rw, int e)
{
    const uint32_t *q = rw + (e >> 1);
    const bool odd = e & 1;
    uint32_t acc = 32768u;
    acc = __dp2a_lo(q[0], odd ? 0x1200u : 0x2212u, acc);
    acc = __dp2a_lo(q[1], odd ? 0x3022u : 0x3830u, acc);
    acc = __dp2a_lo(q[2], odd ? 0x3038u : 0x2230u, acc);
    acc = __dp2a_lo(q[3], odd ? 0x1222u : 0x0012u, acc);
    return acc >> 16;
}

__global__ void __launch_bounds__(DESC_WARPS * 32) k_describe_fused(DescParams P, const FrameGeom *__restrict__ G)
{
    __shared__ __align__(16) uint8_t s_all[DESC_WARPS * DF_WARP_BYTES];
    ORBX_PDL_ENTRY();
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nl = G->nlevels;
    // level boundaries of the frame's selected list: levels are concatenated in order (ORBextractor.cpp:1123).  Lane l holds level l's count; an
    // inclusive warp scan gives the boundaries, a ballot (per keypoint, below) the level that contains list index gidx.
    int incl = lane < nl ? P.nsel[f * nl + lane] : 0;
#pragma unroll
    for (int o = 1; o < ORBX_MAX_LEVELS; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    const int total = __shfl_sync(0xffffffffu, incl, nl - 1);
    int nrows = total;                                                 // output rows of this frame
    if (P.map) nrows = P.counts[f];                                    // filter-first order: k_keep_list has applied the depth / box filter to the selected
                                                                       // POSITIONS and left the survivors' list indices, in order, with their count — only
                                                                       // those get an angle and a descriptor, straight into their final rows
    else {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            P.counts[f] = total <= P.cap ? total : 0;
            if (total > P.cap) atomicOr(P.status, ORBX_DS_KP_OVERFLOW);
        }
        if (total > P.cap) return;
    }
    uint8_t *tile = s_all + wid * DF_WARP_BYTES + 16;                              // 16 bytes of slack on either side of the tile
    uint16_t *rows = reinterpret_cast<uint16_t *>(tile + DF_ROWS * DF_PITCH + 16);
    // a warp walks the frame's rows with the grid's stride (the launch holds about a third of the rows' worth of warps per frame: fewer, longer
    // CTAs, and none that finds nothing to do)
    for (int slot = blockIdx.x * DESC_WARPS + wid; slot < nrows; slot += gridDim.x * DESC_WARPS) {
    const int gidx = P.map ? P.map[(size_t)f * P.map_slab + slot] : slot;        // index in the frame's selected list
    const int level = __popc(__ballot_sync(0xffffffffu, lane < nl && incl <= gidx));
    if (level >= nl) break;                                                        // (a list index past the list: cannot happen)
    const int before = __shfl_sync(0xffffffffu, incl, max(level - 1, 0));
    const int k = gidx - (level > 0 ? before : 0);
    const LevelGeom &g = G->lv[level];
    const uint32_t c = P.sel[(size_t)f * P.sel_slab + g.sel_off + k];
    // pt += (minBorderX, minBorderY) — ORBextractor.cpp:886-887; integer-valued, cvRound is the identity
    const int cx = orbx_px(c) + ORBX_BORDER, cy = orbx_py(c) + ORBX_BORDER;
    const uint8_t *img; int step;
    if (level == 0) { img = P.l0 + (size_t)f * P.l0_fstride; step = (int)P.l0_step; }
    else { img = P.pyr + (size_t)f * P.pyr_slab + g.off; step = g.pitch; }
    const int w = g.w, hgt = g.h;

    // ---- stage the window: image columns from a 16-byte boundary, rows cy-21 .. cy+21 (mirrored outside the level) ----
    const int xw = cx - DF_R;                                                      // window column 0
    const int ax = xw & 15, xal = xw - ax;                                         // its byte in the tile; image column of tile byte 0 (may be -16)
    {
        // 43 rows x four 16-byte chunks.  A lane = (chunk, row of a quad, which quad): the two quads of an iteration lie 16 rows apart (19 x 16 words =
        // 16 banks: no conflict between their 4-word chunks; rows 4 apart would collide three chunks over); all six loads of a lane in flight together
        uint4 v[6];
        const int ch = lane & 3, sub = (lane >> 2) & 3, half = lane >> 4;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int r = 4 * (k < 4 ? k + 4 * half : (k == 4 ? 8 + half : 10 + half)) + sub;
            int y = cy - DF_R + r;
            y = y < 0 ? -y : (y >= hgt ? 2 * hgt - 2 - y : y);
            const int xs = xal + 16 * ch;
            v[k] = make_uint4(0u, 0u, 0u, 0u);
            if (r < DF_ROWS && xs >= 0 && xs + 16 <= step) v[k] = __ldg(reinterpret_cast<const uint4 *>(img + (size_t)y * step + xs));
        }
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int r = 4 * (k < 4 ? k + 4 * half : (k == 4 ? 8 + half : 10 + half)) + sub;
            if (r < DF_ROWS) {
                uint32_t *d = reinterpret_cast<uint32_t *>(tile + r * DF_PITCH + 16 * ch);
                d[0] = v[k].x; d[1] = v[k].y; d[2] = v[k].z; d[3] = v[k].w;
            }
        }
    }
    __syncwarp();
    if (xw < 0 || xw + 2 * DF_R >= w) {                                            // columns outside the level: x -> -x resp. 2(w-1) - x
        for (int r = lane; r < DF_ROWS; r += 32) {
            uint8_t *row = tile + r * DF_PITCH - xal;                              // row[x] = pixel x of that (mirrored) image row
            for (int x = xw; x < 0; x++) row[x] = row[-x];
            for (int x = w; x <= xw + 2 * DF_R; x++) row[x] = row[2 * w - 2 - x];
        }
        __syncwarp();
    }

    // ---- IC_Angle on the staged window: lane = patch row v + 15; the row's 31 columns as eight words brought to the column grid by one
    // warp-uniform funnel shift, masked to the disc, then two IDP.4A per word: S = sum of I, T = sum of (u + 15) I; m10 = sum over rows of T - 15 S ----
    int m10 = 0, m01 = 0;
    if (lane < 2 * ORBX_HALF_PATCH + 1) {
        const int p0 = ax + DF_R - ORBX_HALF_PATCH;                                // tile byte of column u = -15
        const uint32_t *rw32 = reinterpret_cast<const uint32_t *>(tile + (DF_R - ORBX_HALF_PATCH + lane) * DF_PITCH) + (p0 >> 2);
        const int sh = 8 * (p0 & 3);
        const uint4 ma = __ldg(reinterpret_cast<const uint4 *>(g_ic_mask[lane])), mb = __ldg(reinterpret_cast<const uint4 *>(g_ic_mask[lane]) + 1);
        const uint32_t mk[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
        uint32_t wv[9];
#pragma unroll
        for (int i = 0; i < 9; i++) wv[i] = rw32[i];
        uint32_t S = 0, T = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t a = __funnelshift_r(wv[i], wv[i + 1], sh) & mk[i];
            S = __dp4a(a, 0x01010101u, S);
            T = __dp4a(a, 0x03020100u + 0x04040404u * (uint32_t)i, T);
        }
        m10 = (int)T - ORBX_HALF_PATCH * (int)S;
        m01 = (lane - ORBX_HALF_PATCH) * (int)S;
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = cv_fast_atan2((float)m01, (float)m10);

    // ---- row pass of the window: ten 4-pixel groups x 22 row pairs (rows 2 p, 2 p + 1; row 43 is padding).  A lane's unit is one group x one row
    // pair: six words in, ten IDP.4A per row, and the two rows of a column leave as ONE 32-bit store into the column-major buffer.  Five iterations
    // take row pairs 0..15 of two adjacent groups (lanes 0-15 / 16-31: rows two apart sit two banks apart at the 17-word pitch, the neighbour group
    // fills the odd banks), two more take row pairs 16..21 of all ten groups ----
    const int b0 = (ax + 3) & ~3;                                                  // tile byte of row-pass column 0 (<= the byte of column x-18)
    {
        const uint32_t *tw = reinterpret_cast<const uint32_t *>(tile) + (b0 >> 2);
        uint32_t *rw32 = reinterpret_cast<uint32_t *>(rows);
#pragma unroll
        for (int it = 0; it < 7; it++) {
            int gq, rp;
            bool act = true;
            if (it < 5) { gq = 2 * it + (lane >> 4); rp = lane & 15; }
            else { const int g5 = (lane * 43) >> 8; gq = 5 * (it - 5) + g5; rp = 16 + lane - 6 * g5; act = lane < 30; }   // lane / 6: five groups x six row pairs
            if (act) {
                const uint32_t *q = tw + (2 * rp) * (DF_PITCH / 4) + gq;
                uint32_t a0, a1, a2, a3, c0, c1, c2, c3;
                df_hpass4(q[-1], q[0], q[1], a0, a1, a2, a3);
                df_hpass4(q[DF_PITCH / 4 - 1], q[DF_PITCH / 4], q[DF_PITCH / 4 + 1], c0, c1, c2, c3);
                uint32_t *o = rw32 + (4 * gq) * (DF_CP / 2) + rp;
                o[0] = __byte_perm(a0, c0, 0x5410); o[DF_CP / 2] = __byte_perm(a1, c1, 0x5410);                 // (row 2 p, row 2 p + 1) of one column
                o[2 * (DF_CP / 2)] = __byte_perm(a2, c2, 0x5410); o[3 * (DF_CP / 2)] = __byte_perm(a3, c3, 0x5410);
            }
        }
    }
    __syncwarp();

    // ---- rBRIEF on blurred values computed at the sample points: lane i computes descriptor byte i ----
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float ang = __fmul_rn(angle, factorPI);
    const float a = glibc_cosf(ang), b = glibc_sinf(ang);
    signed char pat[32];
    {
        const int4 *pp = reinterpret_cast<const int4 *>(c_pattern + lane * 32);
        *reinterpret_cast<int4 *>(pat) = __ldg(pp);
        *reinterpret_cast<int4 *>(pat + 16) = __ldg(pp + 1);
    }
    const int e0 = (ax + DF_R - b0) * DF_CP + DF_R - 3;                            // entry of the first tap of the keypoint's own pixel
    const uint32_t *rw = reinterpret_cast<const uint32_t *>(rows);
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float x0 = (float)pat[4 * j], y0 = (float)pat[4 * j + 1], x1 = (float)pat[4 * j + 2], y1 = (float)pat[4 * j + 3];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        val |= (df_sample(rw, e0 + c0 * DF_CP + r0) < df_sample(rw, e0 + c1 * DF_CP + r1)) << j;
    }
    uint8_t *drow = P.desc + ((size_t)f * P.cap + slot) * ORBX_DESC_BYTES;
    drow[lane] = (uint8_t)val;
    if (lane == 0) {
        orbx_keypoint kp;
        kp.x = (float)cx; kp.y = (float)cy;
        if (level != 0) { kp.x = __fmul_rn(kp.x, g.scale); kp.y = __fmul_rn(kp.y, g.scale); }   // :1147-1149
        kp.size = g.size; kp.angle = angle; kp.response = (float)orbx_ps(c);
        kp.octave = level; kp.class_id = -1;
        P.kps[(size_t)f * P.cap + slot] = kp;
    }
    __syncwarp();                                                                  // the tile and the row sums are rewritten by the next keypoint
    }
}

void launch_describe_to(orbx_handle *h, int nframes, const uint8_t *l0, size_t l0_step, size_t l0_fstride,
                        orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_counts, const int32_t *d_map, int map_slab)
{
    DescParams P;
    P.map = h->opt_fused_blur ? d_map : nullptr; P.map_slab = map_slab;
    P.l0 = l0; P.l0_step = l0_step; P.l0_fstride = l0_fstride;
    P.pyr = h->d_pyr; P.pyr_slab = h->pyr_slab;
    P.blur = h->d_blur; P.blur_slab = h->blur_slab;
    P.sel = h->d_sel; P.sel_slab = h->geo.sel_entries; P.nsel = h->d_nsel;
    P.kps = d_kps; P.desc = d_desc; P.cap = cap; P.counts = d_counts; P.status = h->d_status;
    const int maxk = h->geo.sel_entries < cap ? h->geo.sel_entries : cap;
    dim3 grid((maxk + DESC_WARPS - 1) / DESC_WARPS, nframes);
    // fused kernel: warps stride over the frame's rows; a third of the CTAs for batches, one warp per row when few frames must fill the machine
    dim3 gridf(nframes >= 16 ? (grid.x + DESC_STRIDE - 1) / DESC_STRIDE : grid.x, nframes);
    ProfScope ps(h, ORBX_K_DESCRIBE);
    if (h->opt_fused_blur) orbx_launch_pdl(h, k_describe_fused, gridf, dim3(DESC_WARPS * 32), 0, h->stream, P, (const FrameGeom *)h->d_geo);
    else k_describe<<<grid, DESC_WARPS * 32, 0, h->stream>>>(P, h->d_geo);
}

// ---- profile C (cv::ORB): IC_Angle on the level, rBRIEF on the float-blurred level, keypoint assembly (orbx_cvorb.cu) ----
// per-level lists [level][list_cap] of (x | y << 16) + Harris response, already in output order; one warp per keypoint
struct DescCParams {
    const uint8_t *img[ORBX_MAX_LEVELS]; int step[ORBX_MAX_LEVELS];
    const uint8_t *blur[ORBX_MAX_LEVELS]; int bstep[ORBX_MAX_LEVELS];
    float scale[ORBX_MAX_LEVELS];
    const uint32_t *fxy; const float *fresp; const int32_t *fcount; int list_cap, nlevels;
    orbx_keypoint *kps; uint8_t *desc; int cap; int32_t *count; int32_t *status;
};
__global__ void __launch_bounds__(DESC_WARPS * 32) k_describe_c(DescCParams P)
{
    ORBX_PDL_ENTRY();
    const int lane = threadIdx.x & 31;
    const int nl = P.nlevels;
    const int gidx = blockIdx.x * DESC_WARPS + (threadIdx.x >> 5);
    int level, k, total;
    {
        const int c = lane < nl ? P.fcount[lane] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < ORBX_MAX_LEVELS; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        total = __shfl_sync(0xffffffffu, incl, nl - 1);
        level = __popc(__ballot_sync(0xffffffffu, lane < nl && incl <= gidx));
        const int before = __shfl_sync(0xffffffffu, incl, max(level - 1, 0));
        k = gidx - (level > 0 ? before : 0);
        if (level >= nl) level = -1;
    }
    if (gidx == 0 && lane == 0) {
        P.count[0] = total <= P.cap ? total : 0;
        if (total > P.cap) atomicOr(P.status, ORBX_DS_KP_OVERFLOW);
    }
    if (level < 0 || total > P.cap) return;
    const uint32_t c = P.fxy[(size_t)level * P.list_cap + k];
    const int cx = (int)(c & 0xFFFFu), cy = (int)(c >> 16);
    const uint8_t *img = P.img[level]; const size_t step = (size_t)P.step[level];
    const uint8_t *bctr = P.blur[level] + (size_t)cy * P.bstep[level] + cx;
    const int bstep = P.bstep[level];
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - ORBX_HALF_PATCH;
        const int au = u < 0 ? -u : u;
        const uint8_t *ctr = img + (size_t)cy * step + cx + u;
        int px[2 * ORBX_HALF_PATCH + 1];
#pragma unroll
        for (int v = -ORBX_HALF_PATCH; v <= ORBX_HALF_PATCH; v++) {
            const int av = v < 0 ? -v : v;
            px[v + ORBX_HALF_PATCH] = au <= c_umax[av] ? (int)__ldg(ctr + (ptrdiff_t)v * (ptrdiff_t)step) : 0;
        }
        int colsum = 0;
#pragma unroll
        for (int v = -ORBX_HALF_PATCH; v <= ORBX_HALF_PATCH; v++) { colsum += px[v + ORBX_HALF_PATCH]; m01 += v * px[v + ORBX_HALF_PATCH]; }
        m10 = u * colsum;
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = cv_fast_atan2((float)m01, (float)m10);
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float ang = __fmul_rn(angle, factorPI);
    const float a = glibc_cosf(ang), b = glibc_sinf(ang);
    signed char pat[32];
    {
        const int4 *pp = reinterpret_cast<const int4 *>(c_pattern + lane * 32);
        *reinterpret_cast<int4 *>(pat) = __ldg(pp);
        *reinterpret_cast<int4 *>(pat + 16) = __ldg(pp + 1);
    }
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float x0 = (float)pat[4 * j], y0 = (float)pat[4 * j + 1], x1 = (float)pat[4 * j + 2], y1 = (float)pat[4 * j + 3];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = __ldg(bctr + r0 * bstep + c0), t1 = __ldg(bctr + r1 * bstep + c1);
        val |= (t0 < t1) << j;
    }
    P.desc[(size_t)gidx * ORBX_DESC_BYTES + lane] = (uint8_t)val;
    if (lane == 0) {
        orbx_keypoint kp;
        const float s = P.scale[level];
        kp.x = __fmul_rn((float)cx, s); kp.y = __fmul_rn((float)cy, s);            // keypoints[i].pt *= scale (orb.cpp)
        kp.size = __fmul_rn((float)ORBX_PATCH, s); kp.angle = angle; kp.response = P.fresp[(size_t)level * P.list_cap + k];
        kp.octave = level; kp.class_id = -1;
        P.kps[gidx] = kp;
    }
}
void launch_describe_c(orbx_handle *h, int nlevels, const uint8_t *const *img, const int *step, const uint8_t *const *blur, const int *bstep,
                       const float *scale, const uint32_t *d_fxy, const float *d_fresp, const int32_t *d_fcount, int list_cap, int max_total,
                       orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_count)
{
    DescCParams P;
    memset(&P, 0, sizeof(P));
    for (int l = 0; l < nlevels; l++) { P.img[l] = img[l]; P.step[l] = step[l]; P.blur[l] = blur[l]; P.bstep[l] = bstep[l]; P.scale[l] = scale[l]; }
    P.fxy = d_fxy; P.fresp = d_fresp; P.fcount = d_fcount; P.list_cap = list_cap; P.nlevels = nlevels;
    P.kps = d_kps; P.desc = d_desc; P.cap = cap; P.count = d_count; P.status = h->d_status;
    ProfScope ps(h, ORBX_K_DESCRIBE);
    k_describe_c<<<(max_total + DESC_WARPS - 1) / DESC_WARPS, DESC_WARPS * 32, 0, h->stream>>>(P);
}

// ---- Harris response of given points of one pyramid level (cv::ORB's HarrisResponses; HARRIS_SCORE of ORBextractor.hpp:48) ----
// integer 7x7 sums of the Sobel-like derivatives, then the fp32 expression in OpenCV's operation order (no FMA)
__global__ void k_harris(const uint8_t *img, size_t step, int w, int h, const int32_t *xy, int n, int bs, float k, float *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x0 = xy[2 * i], y0 = xy[2 * i + 1], r = bs / 2;
    if (x0 - r - 1 < 0 || y0 - r - 1 < 0 || x0 + r + 1 >= w || y0 + r + 1 >= h) { out[i] = 0.f; return; }
    int a = 0, b = 0, c = 0;
    const int st = (int)step;
    for (int u = 0; u < bs; u++) for (int v = 0; v < bs; v++) {
        const uint8_t *p = img + (size_t)(y0 - r + u) * step + (x0 - r + v);
        const int Ix = ((int)__ldg(p + 1) - (int)__ldg(p - 1)) * 2 + ((int)__ldg(p - st + 1) - (int)__ldg(p - st - 1)) + ((int)__ldg(p + st + 1) - (int)__ldg(p + st - 1));
        const int Iy = ((int)__ldg(p + st) - (int)__ldg(p - st)) * 2 + ((int)__ldg(p + st - 1) - (int)__ldg(p - st - 1)) + ((int)__ldg(p + st + 1) - (int)__ldg(p - st + 1));
        a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
    }
    const float scale = __fdiv_rn(1.f, __fmul_rn((float)(4 * bs), 255.f));
    const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float fa = (float)a, fb = (float)b, fc = (float)c, sum = __fadd_rn(fa, fb);
    const float v = __fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(k, sum), sum));
    out[i] = __fmul_rn(v, s4);
}
void launch_harris(orbx_handle *h, const uint8_t *img, size_t step, int w, int hgt, const int32_t *d_xy, int n, int bs, float k, float *d_out)
{
    if (n <= 0) return;
    ProfScope ps(h, ORBX_K_OTHER);
    k_harris<<<(n + 127) / 128, 128, 0, h->stream>>>(img, step, w, hgt, d_xy, n, bs, k, d_out);
}

// ---- device self-tests of the floating-point restatements ----
__global__ void k_test_trig(const float *in, int n, float *oc, float *os)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { oc[i] = glibc_cosf(in[i]); os[i] = glibc_sinf(in[i]); }
}
__global__ void k_test_atan2(const float *y, const float *x, int n, float *o)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = cv_fast_atan2(y[i], x[i]);
}
// sum of the result bit patterns over angle_deg bit patterns [first, last]
__global__ void k_trig_checksum(uint32_t first, uint32_t last, unsigned long long *sums)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    unsigned long long sc = 0, ss = 0;
    const unsigned long long total = (unsigned long long)last - first + 1ull;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float deg = __uint_as_float(first + (uint32_t)i);
        const float rad = __fmul_rn(deg, factorPI);
        sc += __float_as_uint(glibc_cosf(rad));
        ss += __float_as_uint(glibc_sinf(rad));
    }
    for (int o = 16; o > 0; o >>= 1) { sc += __shfl_xor_sync(0xffffffffu, sc, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sums[0], sc); atomicAdd(&sums[1], ss); }
}
void launch_test_trig(orbx_handle *h, const float *d_in, int n, float *d_c, float *d_s)
{
    k_test_trig<<<(n + 255) / 256, 256, 0, h->stream>>>(d_in, n, d_c, d_s); h->launches++;
}
void launch_test_atan2(orbx_handle *h, const float *d_y, const float *d_x, int n, float *d_o)
{
    k_test_atan2<<<(n + 255) / 256, 256, 0, h->stream>>>(d_y, d_x, n, d_o); h->launches++;
}
void launch_trig_checksum(orbx_handle *h, uint32_t first, uint32_t last, unsigned long long *d_sums)
{
    k_trig_checksum<<<h->sm_count * 8, 256, 0, h->stream>>>(first, last, d_sums); h->launches++;
}
