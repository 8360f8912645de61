"""cvorb_oracle.py — ORACLE (test infrastructure): stage-wise restatement of cv::ORB ("profile C"), the extractor of the reference's
own gtest (reference dynamic_visual_slam/test/test_dbow2_integration.cpp:19,38: cv::ORB::create(...)->detectAndCompute) and of
BASELINE.json configs[0] / north_star stage 3 (Harris scoring + top-N retention).

The arithmetic lives in OpenCV (features2d/src/orb.cpp, imgproc resize/smooth), not under /root/reference; it is restated here from its
published algorithm and PINNED against python cv2 4.13.0's cv2.ORB_create(...).detectAndCompute in tests/test_cvorb_oracle.py:
  * pyramid: level sizes cvRound(cols / scale_l), chained cv::resize(INTER_LINEAR_EXACT) — 8.8 fixed-point coefficients, one rounding;
  * per level: whole-image FAST-9/16 (threshold 20, 3x3 NMS), points closer than 31 px to the border dropped,
    KeyPointsFilter::retainBest(2 N_l) by FAST score (every point tied with the cut-off value is kept), HarrisResponses (7x7 block,
    k = 0.04, fp32 expression in OpenCV's operation order), retainBest(N_l) by Harris;
  * IC_Angle (cv::fastAtan2), pt *= scale_l, size = 31 * scale_l;
  * cv::GaussianBlur(7x7, sigma 2) on the level AS A SUB-MATRIX, which takes OpenCV's float path (sepFilter2D with the CV_32F Gaussian
    kernel: fp32 fused multiply-adds left to right, then the symmetric pairs top/bottom, one rounding to u8 — the arithmetic of cv2
    4.13.0's AVX2/FMA build, found by matching its descriptors), unlike the frontend's extractor whose .clone() takes the fixed-point path;
  * rBRIEF-256 with the same fp32 rotation arithmetic as the frontend's extractor.
Output order inside a level is std::nth_element's in OpenCV; this restatement returns each level sorted by (response descending, y, x)
and comparisons are on SETS, as north_star asks ("identical except for ties at the retention cutoff").
Only tests/, __graft_entry__ and bench.py's CPU legs import this module.
"""
import numpy as np

import c_oracle as co

GAUSS7_F32 = np.array([1032826801, 1040595070, 1044597305, 1046301408, 1044597305, 1040595070, 1032826801], np.uint32).view(np.float32)
"""cv::getGaussianKernel(7, 2, CV_32F) as float32 bit patterns (cv2 4.13.0); orb_oracle.c holds the same constants"""


def resize_linear_exact(src, dw, dh):
    return co.resize_linear_exact(src, dw, dh)


def blur7_float(img):
    """cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) through sepFilter2D's float path (orc_gaussian_blur7_f32)"""
    return co.gaussian_blur7_f32(img)


def retain_best(resp, n):
    """indices kept by KeyPointsFilter::retainBest(n): the n largest responses and everything tied with the n-th"""
    if n < 0 or len(resp) <= n:
        return np.arange(len(resp))
    if n == 0:
        return np.zeros(0, np.int64)
    cut = np.sort(resp)[::-1][n - 1]
    return np.nonzero(resp >= cut)[0]


class CvOrb:
    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, edgeThreshold=31, patchSize=31, fastThreshold=20):
        self.nfeatures, self.nlevels, self.edge, self.patch, self.fast_th = nfeatures, nlevels, edgeThreshold, patchSize, fastThreshold
        self.scaleFactor = np.float32(scaleFactor)
        sf = float(self.scaleFactor)                                  # the double member is initialised from a float argument
        self.scale = [np.float32(pow(sf, l)) for l in range(nlevels)]
        factor = np.float32(1.0 / sf)
        nd = np.float32(nfeatures) * (np.float32(1) - factor) / (np.float32(1) - np.float32(pow(float(factor), float(nlevels))))
        self.nfeat, s = [], 0
        for l in range(nlevels - 1):
            self.nfeat.append(int(np.rint(nd)))
            s += self.nfeat[-1]
            nd = np.float32(nd * factor)
        self.nfeat.append(max(nfeatures - s, 0))
        self.umax = co.COracle().umax

    def level_size(self, w, h, l):
        return int(np.rint(np.float32(w) / self.scale[l])), int(np.rint(np.float32(h) / self.scale[l]))

    def extract(self, img, trace=None):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        pyr = [img]
        for l in range(1, self.nlevels):
            lw, lh = self.level_size(w, h, l)
            pyr.append(resize_linear_exact(pyr[-1], lw, lh))
        kps, descs, per_level = [], [], []
        for l, lv in enumerate(pyr):
            lh, lw = lv.shape
            c = co.fast_roi(lv, self.fast_th, cap=max(lw * lh // 4, 4096))
            keep = (c["x"] >= self.edge) & (c["x"] < lw - self.edge) & (c["y"] >= self.edge) & (c["y"] < lh - self.edge)
            c = c[keep]
            c = c[retain_best(c["score"].astype(np.float32), 2 * self.nfeat[l])]
            harris = np.array([co.harris_response(lv, int(x), int(y), 7, 0.04) for x, y in zip(c["x"], c["y"])], np.float32)
            sel = retain_best(harris, self.nfeat[l])
            c, harris = c[sel], harris[sel]
            order = np.lexsort((c["x"], c["y"], -harris.astype(np.float64)))
            c, harris = c[order], harris[order]
            blur = blur7_float(lv)
            k = np.zeros(len(c), co.KP_DTYPE)
            d = np.zeros((len(c), 32), np.uint8)
            for i in range(len(c)):
                cx, cy = int(c["x"][i]), int(c["y"][i])
                ang = np.float32(co.ic_angle(lv, cx, cy, self.umax))
                k[i] = (np.float32(cx) * self.scale[l], np.float32(cy) * self.scale[l], np.float32(self.patch) * self.scale[l], ang, harris[i], l, -1)
                d[i] = co.descriptor(blur, cx, cy, ang)
            kps.append(k); descs.append(d); per_level.append(len(c))
        if trace is not None:
            trace.update(pyramid=pyr, per_level=per_level)
        return np.concatenate(kps), np.concatenate(descs)
