"""cv_oracle.py — ORACLE layer L-A (test infrastructure, NOT product code).

A literal Python restatement of the reference's ORB_SLAM3::ORBextractor
(dynamic_visual_slam/src/ORBextractor.cpp) that calls the *same OpenCV primitives the reference
calls* through python cv2 (4.13.0 in the build container): cv2.resize(INTER_LINEAR),
cv2.FastFeatureDetector (TYPE_9_16, nms), cv2.GaussianBlur(7x7, sigma 2), cv2.fastAtan2,
cv2.BFMatcher(NORM_HAMMING).  The std::list / std::sort logic of DistributeOctTree is restated
with Python lists and the REAL libstdc++ std::sort (oracle/stdsort_shim.cpp).

This layer needs cv2 and is used (a) to pin the dependency-free C oracle (orb_oracle.c) bit-for-bit
and (b) to generate the golden vectors under tests/golden/ (oracle/gen_golden.py).
Only tests/, gen_golden.py and bench.py's reference arm import it.
"""
import ctypes
import math
import os

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32 = np.float32

PATCH_SIZE, HALF_PATCH_SIZE, EDGE_THRESHOLD = 31, 15, 19        # ORBextractor.cpp:71-73


def _pattern():
    txt = open(os.path.join(_HERE, "..", "include", "orbx_pattern.inc")).read()
    txt = txt[txt.index("*/") + 2:]
    vals = [int(v) for v in txt.replace("\n", "").split(",") if v.strip()]
    assert len(vals) == 1024
    return np.array(vals, dtype=np.int32).reshape(512, 2)


def cv_round(v):
    """cvRound: round half to even (numpy rint does the same)."""
    return int(np.rint(v))


_shim = None


def real_std_sort(cnt, ulx):
    """Permutation produced by the real libstdc++ std::sort + compareNodes (ORBextractor.cpp:700)."""
    global _shim
    if _shim is None:
        _shim = ctypes.CDLL(os.path.join(_HERE, "libstdsort_shim.so"))
    n = len(cnt)
    c = (ctypes.c_int * n)(*cnt)
    u = (ctypes.c_int * n)(*ulx)
    p = (ctypes.c_int * n)(*range(n))
    _shim.real_std_sort(c, u, p, n)
    return list(p)


class _Node:
    __slots__ = ("UL", "UR", "BL", "BR", "keys", "noMore")

    def __init__(self):
        self.UL = self.UR = self.BL = self.BR = (0, 0)
        self.keys = []
        self.noMore = False

    def divide(self):                                               # ORBextractor.cpp:480-536
        halfX = int(math.ceil(_f32(self.UR[0] - self.UL[0]) / _f32(2)))
        halfY = int(math.ceil(_f32(self.BR[1] - self.UL[1]) / _f32(2)))
        n1, n2, n3, n4 = _Node(), _Node(), _Node(), _Node()
        n1.UL = self.UL
        n1.UR = (self.UL[0] + halfX, self.UL[1])
        n1.BL = (self.UL[0], self.UL[1] + halfY)
        n1.BR = (self.UL[0] + halfX, self.UL[1] + halfY)
        n2.UL, n2.UR, n2.BL, n2.BR = n1.UR, self.UR, n1.BR, (self.UR[0], self.UL[1] + halfY)
        n3.UL, n3.UR, n3.BL, n3.BR = n1.BL, n1.BR, self.BL, (n1.BR[0], self.BL[1])
        n4.UL, n4.UR, n4.BL, n4.BR = n3.UR, n2.BR, n3.BR, self.BR
        for kp in self.keys:
            if kp[0] < n1.UR[0]:
                (n1 if kp[1] < n1.BR[1] else n3).keys.append(kp)
            elif kp[1] < n1.BR[1]:
                n2.keys.append(kp)
            else:
                n4.keys.append(kp)
        for n in (n1, n2, n3, n4):
            if len(n.keys) == 1:
                n.noMore = True
        return n1, n2, n3, n4


def distribute_octtree(keys, minX, maxX, minY, maxY, N):
    """DistributeOctTree — ORBextractor.cpp:555-779.  keys: list of (x, y, response)."""
    r = _f32(maxX - minX) / _f32(maxY - minY)                        # :559, C round(): half away from zero
    nIni = int(math.floor(float(r) + 0.5))
    hX = _f32(maxX - minX) / _f32(nIni)
    lNodes = []
    ini = []
    for i in range(nIni):
        ni = _Node()
        ni.UL = (int(hX * _f32(i)), 0)
        ni.UR = (int(hX * _f32(i + 1)), 0)
        ni.BL = (ni.UL[0], maxY - minY)
        ni.BR = (ni.UR[0], maxY - minY)
        lNodes.append(ni)
        ini.append(ni)
    for kp in keys:
        ini[int(_f32(kp[0]) / hX)].keys.append(kp)
    kept = []
    for n in lNodes:
        if len(n.keys) == 1:
            n.noMore = True
            kept.append(n)
        elif len(n.keys) > 0:
            kept.append(n)
    lNodes = kept
    finish = False
    while not finish:
        prevSize = len(lNodes)
        nToExpand = 0
        vSize = []                                   # (count, node)
        front = []                                   # nodes pushed to the front, latest first
        rest = []
        for n in lNodes:
            if n.noMore:
                rest.append(n)
                continue
            for c in n.divide():
                if len(c.keys) > 0:
                    front.insert(0, c)
                    if len(c.keys) > 1:
                        nToExpand += 1
                        vSize.append((len(c.keys), c))
        lNodes = front + rest
        if len(lNodes) >= N or len(lNodes) == prevSize:
            finish = True
        elif len(lNodes) + nToExpand * 3 > N:
            while not finish:
                prevSize = len(lNodes)
                prev = vSize
                vSize = []
                perm = real_std_sort([p[0] for p in prev], [p[1].UL[0] for p in prev])
                prev = [prev[i] for i in perm]
                for j in range(len(prev) - 1, -1, -1):
                    node = prev[j][1]
                    for c in node.divide():
                        if len(c.keys) > 0:
                            lNodes.insert(0, c)
                            if len(c.keys) > 1:
                                vSize.append((len(c.keys), c))
                    for idx, n in enumerate(lNodes):
                        if n is node:
                            del lNodes[idx]
                            break
                    if len(lNodes) >= N:
                        break
                if len(lNodes) >= N or len(lNodes) == prevSize:
                    finish = True
    out = []
    for n in lNodes:
        best = n.keys[0]
        for kp in n.keys[1:]:
            if kp[2] > best[2]:
                best = kp
        out.append(best)
    return out


class ORBextractorCV:
    """ORB_SLAM3::ORBextractor over cv2 primitives (reference ORBextractor.cpp:409-469 ctor)."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7):
        assert cv2 is not None, "cv2 is required for oracle layer L-A"
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.iniThFAST, self.minThFAST = iniThFAST, minThFAST
        sf = float(_f32(scaleFactor))                     # double member initialised from float
        self.scale = [_f32(1.0)]
        for i in range(1, nlevels):
            self.scale.append(_f32(float(self.scale[-1]) * sf))
        self.inv_scale = [_f32(1.0) / s for s in self.scale]
        self.sigma2 = [s * s for s in self.scale]
        factor = _f32(1.0 / sf)
        nd = _f32(nfeatures) * (_f32(1) - factor) / (_f32(1) - _f32(math.pow(float(factor), float(nlevels))))
        self.nfeat = []
        tot = 0
        for _ in range(nlevels - 1):
            self.nfeat.append(cv_round(nd))
            tot += self.nfeat[-1]
            nd = _f32(nd * factor)
        self.nfeat.append(max(nfeatures - tot, 0))
        self.pattern = _pattern()
        umax = [0] * (HALF_PATCH_SIZE + 1)
        vmax = int(math.floor(_f32(HALF_PATCH_SIZE) * np.sqrt(_f32(2)) / _f32(2) + _f32(1)))
        vmin = int(math.ceil(_f32(HALF_PATCH_SIZE) * np.sqrt(_f32(2)) / _f32(2)))
        for v in range(vmax + 1):
            umax[v] = cv_round(math.sqrt(HALF_PATCH_SIZE * HALF_PATCH_SIZE - v * v))
        v0 = 0
        for v in range(HALF_PATCH_SIZE, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax
        self.fast_ini = cv2.FastFeatureDetector_create(iniThFAST, True)
        self.fast_min = cv2.FastFeatureDetector_create(minThFAST, True)
        self.pyramid = []

    def level_size(self, w, h, level):
        s = self.inv_scale[level]
        return cv_round(_f32(w) * s), cv_round(_f32(h) * s)

    def compute_pyramid(self, image):                               # ORBextractor.cpp:1169-1194
        h, w = image.shape
        self.pyramid = [np.ascontiguousarray(image)]
        for level in range(1, self.nlevels):
            lw, lh = self.level_size(w, h, level)
            self.pyramid.append(cv2.resize(self.pyramid[level - 1], (lw, lh), interpolation=cv2.INTER_LINEAR))
        return self.pyramid

    def fast_cells(self, img):
        """The cell loop of ComputeKeyPointsOctTree — ORBextractor.cpp:785-872."""
        h, w = img.shape
        W = _f32(35)
        minBX = minBY = EDGE_THRESHOLD - 3
        maxBX, maxBY = w - EDGE_THRESHOLD + 3, h - EDGE_THRESHOLD + 3
        width, height = _f32(maxBX - minBX), _f32(maxBY - minBY)
        nCols, nRows = int(width / W), int(height / W)
        if nCols < 1 or nRows < 1:
            return []
        wCell = int(math.ceil(width / _f32(nCols)))
        hCell = int(math.ceil(height / _f32(nRows)))
        out = []
        for i in range(nRows):
            iniY = minBY + i * hCell
            maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minBX + j * wCell
                maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                roi = img[iniY:maxY, iniX:maxX]
                kps = self.fast_ini.detect(roi)
                if len(kps) == 0:
                    kps = self.fast_min.detect(roi)
                for kp in kps:
                    out.append((int(kp.pt[0]) + j * wCell, int(kp.pt[1]) + i * hCell, int(kp.response)))
        return out

    def ic_angle(self, img, cx, cy):                                # ORBextractor.cpp:76-103
        m01 = m10 = 0
        patch = img[cy - HALF_PATCH_SIZE:cy + HALF_PATCH_SIZE + 1, cx - HALF_PATCH_SIZE:cx + HALF_PATCH_SIZE + 1].astype(np.int64)
        for v in range(-HALF_PATCH_SIZE, HALF_PATCH_SIZE + 1):
            d = self.umax[abs(v)]
            row = patch[v + HALF_PATCH_SIZE, HALF_PATCH_SIZE - d:HALF_PATCH_SIZE + d + 1]
            us = np.arange(-d, d + 1)
            m10 += int((us * row).sum())
            m01 += v * int(row.sum())
        return float(cv2.fastAtan2(float(m01), float(m10)))

    def descriptor(self, blur, cx, cy, angle_deg):                  # ORBextractor.cpp:106-146
        import ctypes as ct
        libm = _libm()
        factorPI = _f32(np.float64(np.pi) / np.float64(_f32(180.0)))
        ang = _f32(_f32(angle_deg) * factorPI)
        a = _f32(libm.cosf(ct.c_float(float(ang))))
        b = _f32(libm.sinf(ct.c_float(float(ang))))
        px = self.pattern[:, 0].astype(np.float32)
        py = self.pattern[:, 1].astype(np.float32)
        dy = np.rint((px * b).astype(np.float32) + (py * a).astype(np.float32)).astype(np.int64)
        dx = np.rint((px * a).astype(np.float32) - (py * b).astype(np.float32)).astype(np.int64)
        vals = blur[cy + dy, cx + dx].astype(np.int32)
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
        return np.packbits(bits.reshape(32, 8)[:, ::-1], axis=1).reshape(32)

    def __call__(self, image, trace=None):
        """ORBextractor::operator() — ORBextractor.cpp:1086-1167.  Returns (keypoints Nx7 float64
        table [x, y, size, angle, response, octave, class_id], descriptors Nx32 uint8)."""
        if image is None or image.size == 0:
            return -1
        assert image.dtype == np.uint8 and image.ndim == 2
        pyr = self.compute_pyramid(image)
        kps_all, desc_all = [], []
        if trace is not None:
            trace["pyramid"] = [p.copy() for p in pyr]
            trace["cands"], trace["selected"], trace["blurred"] = [], [], []
        for level, img in enumerate(pyr):
            h, w = img.shape
            minBX = minBY = EDGE_THRESHOLD - 3
            maxBX, maxBY = w - EDGE_THRESHOLD + 3, h - EDGE_THRESHOLD + 3
            cands = self.fast_cells(img) if (maxBX - minBX >= 35 and maxBY - minBY >= 35) else []
            sel = distribute_octtree(cands, minBX, maxBX, minBY, maxBY, self.nfeat[level]) if cands else []
            blur = cv2.GaussianBlur(img.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            if trace is not None:
                trace["cands"].append(cands)
                trace["selected"].append(sel)
                trace["blurred"].append(blur)
            size = float(int(_f32(PATCH_SIZE) * self.scale[level]))
            for (x, y, resp) in sel:
                fx, fy = _f32(x + minBX), _f32(y + minBY)
                cx, cy = cv_round(fx), cv_round(fy)
                ang = self.ic_angle(img, cx, cy)
                desc_all.append(self.descriptor(blur, cx, cy, ang))
                if level != 0:
                    fx, fy = _f32(fx * self.scale[level]), _f32(fy * self.scale[level])
                kps_all.append((float(fx), float(fy), size, ang, float(resp), level, -1))
        kps = np.array(kps_all, dtype=np.float64).reshape(-1, 7)
        desc = np.array(desc_all, dtype=np.uint8).reshape(-1, 32)
        return kps, desc


_libm_h = None


def _libm():
    global _libm_h
    if _libm_h is None:
        import ctypes as ct
        _libm_h = ct.CDLL("libm.so.6")
        _libm_h.cosf.restype = ct.c_float
        _libm_h.sinf.restype = ct.c_float
        _libm_h.cosf.argtypes = [ct.c_float]
        _libm_h.sinf.argtypes = [ct.c_float]
    return _libm_h


def bf_match(q, t):
    """cv::BFMatcher(NORM_HAMMING).match — frontend.cpp:1123. Returns list of (queryIdx, trainIdx, distance)."""
    m = cv2.BFMatcher(cv2.NORM_HAMMING).match(np.ascontiguousarray(q), np.ascontiguousarray(t))
    return [(d.queryIdx, d.trainIdx, d.distance) for d in m]


def bf_knn2(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(np.ascontiguousarray(q), np.ascontiguousarray(t), k=2)
    return [[(d.queryIdx, d.trainIdx, d.distance) for d in row] for row in m]
