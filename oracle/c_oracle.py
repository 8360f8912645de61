"""c_oracle.py — ctypes binding of the dependency-free C oracle (oracle/liborb_oracle.so).

ORACLE = test infrastructure.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module; the product never does.
"""
import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_LEVELS = 16


class Keypoint(ct.Structure):
    _fields_ = [("x", ct.c_float), ("y", ct.c_float), ("size", ct.c_float), ("angle", ct.c_float),
                ("response", ct.c_float), ("octave", ct.c_int32), ("class_id", ct.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
DM_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
CAND_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("score", "<i4")])
BOX_DTYPE = np.dtype([("cx", "<f8"), ("cy", "<f8"), ("w", "<f8"), ("h", "<f8"), ("class_id", "<i4"), ("pad", "<i4")])
assert KP_DTYPE.itemsize == 28 and DM_DTYPE.itemsize == 16 and BOX_DTYPE.itemsize == 40


class Extractor(ct.Structure):
    _fields_ = [("nfeatures", ct.c_int), ("nlevels", ct.c_int), ("iniThFAST", ct.c_int), ("minThFAST", ct.c_int),
                ("scaleFactor", ct.c_double),
                ("scale", ct.c_float * MAX_LEVELS), ("inv_scale", ct.c_float * MAX_LEVELS),
                ("sigma2", ct.c_float * MAX_LEVELS), ("inv_sigma2", ct.c_float * MAX_LEVELS),
                ("nfeat_level", ct.c_int * MAX_LEVELS), ("umax", ct.c_int * 16)]


class Trace(ct.Structure):
    _fields_ = [("pyramid", ct.c_void_p), ("blurred", ct.c_void_p), ("cands", ct.c_void_p),
                ("cand_cap", ct.c_int32), ("ncands", ct.c_int32 * MAX_LEVELS), ("nkeys", ct.c_int32 * MAX_LEVELS),
                ("lw", ct.c_int32 * MAX_LEVELS), ("lh", ct.c_int32 * MAX_LEVELS)]


def build(force=False):
    so = os.path.join(_HERE, "liborb_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("orb_oracle.c", "orb_oracle.h", "stdsort_shim.cpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ct.CDLL(build())
        _lib.orc_fast_atan2.restype = ct.c_float
        _lib.orc_fast_atan2.argtypes = [ct.c_float, ct.c_float]
        _lib.orc_ic_angle.restype = ct.c_float
        _lib.orc_cosf.restype = ct.c_float
        _lib.orc_sinf.restype = ct.c_float
        _lib.orc_cosf.argtypes = [ct.c_float]
        _lib.orc_sinf.argtypes = [ct.c_float]
    return _lib


def _p(a):
    return a.ctypes.data_as(ct.c_void_p)


class GeometryError(ValueError):
    pass


class COracle:
    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7):
        self.ex = Extractor()
        rc = lib().orc_extractor_init(ct.byref(self.ex), nfeatures, ct.c_float(scaleFactor), nlevels, iniThFAST, minThFAST)
        if rc != 0:
            raise ValueError("bad extractor parameters")
        self.nlevels = nlevels

    @property
    def nfeat(self):
        return list(self.ex.nfeat_level[:self.nlevels])

    @property
    def scale(self):
        return np.array(self.ex.scale[:self.nlevels], dtype=np.float32)

    @property
    def umax(self):
        return list(self.ex.umax)

    def level_size(self, w, h, level):
        lw, lh = ct.c_int(), ct.c_int()
        lib().orc_level_size(ct.byref(self.ex), w, h, level, ct.byref(lw), ct.byref(lh))
        return lw.value, lh.value

    def geometry_status(self, w, h):
        """0 = supported; 1/2/3 = the reference throws or faults on this frame size (orc_geometry_status)"""
        return int(lib().orc_geometry_status(ct.byref(self.ex), int(w), int(h)))

    def extract(self, gray, cap=20000, trace=False):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        h, w = gray.shape
        kps = np.zeros(cap, dtype=KP_DTYPE)
        desc = np.zeros((cap, 32), dtype=np.uint8)
        tr = None
        bufs = None
        if trace:
            sizes = [self.level_size(w, h, l) for l in range(self.nlevels)]
            total = sum(a * b for a, b in sizes)
            cand_cap = max(1 << 17, w * h // 4 + 1024)
            bufs = dict(pyr=np.zeros(total, np.uint8), blur=np.zeros(total, np.uint8),
                        cands=np.zeros((self.nlevels, cand_cap), CAND_DTYPE))
            tr = Trace()
            tr.pyramid, tr.blurred, tr.cands, tr.cand_cap = _p(bufs["pyr"]), _p(bufs["blur"]), _p(bufs["cands"]), cand_cap
        n = lib().orc_extract(ct.byref(self.ex), _p(gray), w, h, ct.c_size_t(gray.strides[0]), _p(kps), _p(desc), cap,
                              ct.byref(tr) if tr is not None else None)
        if n == -3:
            raise GeometryError("frame size %dx%d is outside the reference's defined domain (status %d)" % (w, h, self.geometry_status(w, h)))
        if n < 0:
            raise RuntimeError("orc_extract failed: %d" % n)
        out = dict(kps=kps[:n].copy(), desc=desc[:n].copy())
        if trace:
            off = 0
            pyr, blur, cands = [], [], []
            for l, (lw, lh) in enumerate(sizes):
                pyr.append(bufs["pyr"][off:off + lw * lh].reshape(lh, lw).copy())
                blur.append(bufs["blur"][off:off + lw * lh].reshape(lh, lw).copy())
                cands.append(bufs["cands"][l, :tr.ncands[l]].copy())
                off += lw * lh
            out.update(pyramid=pyr, blurred=blur, cands=cands, nkeys=list(tr.nkeys[:self.nlevels]))
        return out

    def extract_batch(self, frames, cap=4096, nthreads=0):
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        nf, h, w = frames.shape
        kps = np.zeros((nf, cap), dtype=KP_DTYPE)
        desc = np.zeros((nf, cap, 32), dtype=np.uint8)
        counts = np.zeros(nf, dtype=np.int32)
        rc = lib().orc_extract_batch(ct.byref(self.ex), _p(frames), nf, w, h, _p(kps), _p(desc), cap, _p(counts), nthreads)
        if rc != 0:
            raise RuntimeError("orc_extract_batch failed")
        return kps, desc, counts


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear(_p(src), src.shape[1], src.shape[0], ct.c_size_t(src.strides[0]), _p(dst), dw, dh, ct.c_size_t(dw))
    return dst


def fast_roi(img, threshold, cap=1 << 16):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.zeros(cap, CAND_DTYPE)
    n = lib().orc_fast_roi(_p(img), ct.c_size_t(img.strides[0]), img.shape[1], img.shape[0], threshold, _p(out), cap)
    return out[:n].copy()


def fast_cells(orc, img, cap=1 << 17):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.zeros(cap, CAND_DTYPE)
    n = lib().orc_fast_cells(ct.byref(orc.ex), _p(img), ct.c_size_t(img.strides[0]), img.shape[1], img.shape[0], _p(out), cap)
    return out[:n].copy()


def distribute_octtree(cands, minX, maxX, minY, maxY, N):
    cands = np.ascontiguousarray(cands, dtype=CAND_DTYPE)
    out = np.zeros(max(len(cands), 1) + 8, CAND_DTYPE)
    n = lib().orc_distribute_octtree(_p(cands), len(cands), minX, maxX, minY, maxY, N, _p(out), len(out))
    if n < 0:
        raise RuntimeError("orc_distribute_octtree failed: %d" % n)
    return out[:n].copy()


def gaussian_blur7(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.zeros_like(img)
    lib().orc_gaussian_blur7(_p(img), img.shape[1], img.shape[0], ct.c_size_t(img.strides[0]), _p(out), ct.c_size_t(out.strides[0]))
    return out


def resize_linear_exact(src, dw, dh):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_exact(_p(src), src.shape[1], src.shape[0], ct.c_size_t(src.strides[0]), _p(dst), dw, dh, ct.c_size_t(dw))
    return dst


def gaussian_blur7_f32(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.zeros_like(img)
    lib().orc_gaussian_blur7_f32(_p(img), img.shape[1], img.shape[0], ct.c_size_t(img.strides[0]), _p(out), ct.c_size_t(out.strides[0]))
    return out


def ic_angle(img, cx, cy, umax):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    f = lib().orc_ic_angle
    f.restype = ct.c_float
    return float(f(_p(img), ct.c_size_t(img.strides[0]), int(cx), int(cy), (ct.c_int * 16)(*umax)))


def descriptor(blur, cx, cy, angle_deg):
    blur = np.ascontiguousarray(blur, dtype=np.uint8)
    d = np.zeros(32, np.uint8)
    lib().orc_descriptor(_p(blur), ct.c_size_t(blur.strides[0]), int(cx), int(cy), ct.c_float(angle_deg), _p(d))
    return d


def fmat_inliers(p1, p2, F, thresh=2.0):
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    F = np.ascontiguousarray(F, np.float64).reshape(9)
    mask = np.zeros(max(len(p1), 1), np.uint8)
    n = lib().orc_fmat_inliers(_p(p1), _p(p2), len(p1), _p(F), ct.c_double(thresh), _p(mask))
    return n, mask[:len(p1)].copy()


def pnp_points(prev_kps, curr_kps, matches, depth, fx, fy, cx, cy):
    """Frontend::estimateCameraPose's correspondence loop (frontend.cpp:858-892): (points3d [n,3] f32, points2d [n,2] f32) in match order"""
    prev_kps = np.ascontiguousarray(prev_kps, KP_DTYPE); curr_kps = np.ascontiguousarray(curr_kps, KP_DTYPE)
    matches = np.ascontiguousarray(matches, DM_DTYPE); depth = np.ascontiguousarray(depth, np.uint16)
    p3 = np.zeros((max(len(matches), 1), 3), np.float32); p2 = np.zeros((max(len(matches), 1), 2), np.float32)
    n = lib().orc_pnp_points(_p(prev_kps), _p(curr_kps), _p(matches), len(matches), _p(depth), depth.shape[1], depth.shape[0], ct.c_size_t(depth.strides[0]),
                             ct.c_float(fx), ct.c_float(fy), ct.c_float(cx), ct.c_float(cy), _p(p3), _p(p2))
    return p3[:n].copy(), p2[:n].copy()


def pnp_inliers(p3, p2, R, t, fx, fy, cx, cy, thresh=4.0):
    """inlier count and mask of a pose under cv::solvePnPRansac's error (projectPoints without distortion, squared pixel error <= thresh^2)"""
    p3 = np.ascontiguousarray(p3, np.float32).reshape(-1, 3); p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    R = np.ascontiguousarray(R, np.float64).reshape(9); t = np.ascontiguousarray(t, np.float64).reshape(3)
    mask = np.zeros(len(p3), np.uint8)
    n = lib().orc_pnp_inliers(_p(p3), _p(p2), len(p3), _p(R), _p(t), ct.c_double(fx), ct.c_double(fy), ct.c_double(cx), ct.c_double(cy), ct.c_double(thresh), _p(mask))
    return int(n), mask


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(ct.c_float(y), ct.c_float(x)))


def filter_depth(kps, desc, depth, min_depth=0.3, max_depth=3.0):
    kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
    desc = np.ascontiguousarray(desc, dtype=np.uint8)
    depth = np.ascontiguousarray(depth, dtype=np.uint16)
    n = len(kps)
    okps = np.zeros(max(n, 1), KP_DTYPE)
    odesc = np.zeros((max(n, 1), 32), np.uint8)
    oidx = np.zeros(max(n, 1), np.int32)
    m = lib().orc_filter_depth(_p(kps), _p(desc), n, _p(depth), depth.shape[1], depth.shape[0],
                               ct.c_size_t(depth.strides[0] // 2), ct.c_float(min_depth), ct.c_float(max_depth),
                               _p(okps), _p(odesc), _p(oidx))
    return okps[:m].copy(), odesc[:m].copy(), oidx[:m].copy()


def categorize(px, py, boxes):
    boxes = np.ascontiguousarray(boxes, dtype=BOX_DTYPE)
    return lib().orc_categorize(ct.c_float(px), ct.c_float(py), _p(boxes), len(boxes))


def match(q, t, nthreads=0):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    out = np.zeros(len(q), DM_DTYPE)
    n = lib().orc_match(_p(q), len(q), _p(t), len(t), _p(out), nthreads)
    return out[:n].copy()


def knn2(q, t, nthreads=0):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    out = np.zeros((len(q), 2), DM_DTYPE)
    lib().orc_knn2(_p(q), len(q), _p(t), len(t), _p(out), nthreads)
    return out


def reproject(p3, R, t, fx, fy, cx, cy):
    p3 = np.ascontiguousarray(p3, np.float32)
    R = np.ascontiguousarray(R, np.float64)
    t = np.ascontiguousarray(t, np.float64).reshape(3)
    uv = np.zeros(2, np.float32)
    lib().orc_reproject(_p(p3), _p(R), _p(t), ct.c_double(fx), ct.c_double(fy), ct.c_double(cx), ct.c_double(cy), _p(uv))
    return uv


def associate(q, qpx, rows, pos, R, t, fx, fy, cx, cy, max_desc=50.0, max_reproj=5.0, nthreads=0):
    """Backend::associateObservation for a batch: returns (landmark row or -1, reprojection error, Hamming distance)."""
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    qpx = np.ascontiguousarray(qpx, np.float32).reshape(-1, 2)
    rows = np.ascontiguousarray(rows, np.uint8).reshape(-1, 32)
    pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
    R = np.ascontiguousarray(R, np.float64)
    t = np.ascontiguousarray(t, np.float64).reshape(3)
    idx = np.zeros(len(q), np.int32)
    err = np.zeros(len(q), np.float64)
    dist = np.zeros(len(q), np.float32)
    lib().orc_associate(_p(q), _p(qpx), len(q), _p(rows), _p(pos), len(rows), _p(R), _p(t), ct.c_double(fx), ct.c_double(fy),
                        ct.c_double(cx), ct.c_double(cy), ct.c_double(max_desc), ct.c_double(max_reproj), _p(idx), _p(err), _p(dist), nthreads)
    return idx, err, dist


KF_DTYPE = np.dtype([("landmark_id", "<u8"), ("position", "<f8", (3,)), ("pixel_x", "<f8"), ("pixel_y", "<f8"), ("descriptor", "u1", (32,))])
assert KF_DTYPE.itemsize == 80


def pack_keyframe(kps, desc, depth, fx, fy, cx, cy, R, t):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    depth = np.ascontiguousarray(depth, np.uint16)
    R = np.ascontiguousarray(R, np.float64)
    t = np.ascontiguousarray(t, np.float64).reshape(3)
    out = np.zeros(max(len(kps), 1), KF_DTYPE)
    lib().orc_pack_keyframe.restype = ct.c_int
    m = lib().orc_pack_keyframe(_p(kps), _p(desc), len(kps), _p(depth), depth.shape[1], depth.shape[0], ct.c_size_t(depth.strides[0] // 2),
                                ct.c_float(fx), ct.c_float(fy), ct.c_float(cx), ct.c_float(cy), _p(R), _p(t), _p(out))
    return out[:m].copy()


def harris_response(img, x, y, block=7, k=0.04):
    img = np.ascontiguousarray(img, np.uint8)
    f = lib().orc_harris_response
    f.restype = ct.c_float
    return float(f(_p(img), ct.c_size_t(img.strides[0]), int(x), int(y), int(block), ct.c_float(k)))


def bgr2gray(bgr):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    out = np.zeros((h, w), np.uint8)
    lib().orc_bgr2gray(_p(bgr), w, h, ct.c_size_t(bgr.strides[0]), _p(out), ct.c_size_t(w))
    return out


def synth_gray(seed, frame, w, h):
    out = np.zeros((h, w), np.uint8)
    lib().orc_synth_gray(ct.c_uint32(seed), frame, w, h, _p(out), ct.c_size_t(w))
    return out


def synth_depth(seed, frame, w, h):
    out = np.zeros((h, w), np.uint16)
    lib().orc_synth_depth(ct.c_uint32(seed), frame, w, h, _p(out), ct.c_size_t(w))
    return out


def synth_boxes(seed, frame, w, h):
    out = np.zeros(4, BOX_DTYPE)
    n = lib().orc_synth_boxes(ct.c_uint32(seed), int(frame), int(w), int(h), _p(out))
    return out[:n].copy()


def filter_boxes(kps, desc, boxes, drop_mask=1):
    kps = np.ascontiguousarray(kps, dtype=KP_DTYPE)
    desc = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
    boxes = np.ascontiguousarray(boxes, dtype=BOX_DTYPE)
    okps = np.zeros(max(len(kps), 1), KP_DTYPE)
    odesc = np.zeros((max(len(kps), 1), 32), np.uint8)
    m = lib().orc_filter_boxes(_p(kps), _p(desc), len(kps), _p(boxes), len(boxes), ct.c_uint64(drop_mask), _p(okps), _p(odesc))
    return okps[:m].copy(), odesc[:m].copy()


def synth_descriptors(seed, first_row, nrows):
    out = np.zeros((nrows, 32), np.uint8)
    lib().orc_synth_descriptors(ct.c_uint32(seed), ct.c_uint64(first_row), nrows, _p(out))
    return out


def cull_keyframe(response, match_query, max_new=200, min_response=50.0):
    """Frontend feature culling (reference frontend.cpp:1168-1218): indices into the filtered keypoints, matched first."""
    r = np.ascontiguousarray(response, np.float32)
    q = np.ascontiguousarray(match_query, np.int32)
    out = np.zeros(len(q) + max(0, max_new) + 1, np.int32)
    f = lib().orc_cull_keyframe
    f.restype = ct.c_int
    m = f(_p(r), len(r), _p(q), len(q), int(max_new), ct.c_float(min_response), _p(out))
    if m < 0:
        raise ValueError("match query index out of range")
    return out[:m].copy()


def trig_checksum(first_bits, last_bits, nthreads=0):
    """Wrapping sums of the bit patterns of cosf / sinf(angle * factorPI) over the float bit patterns [first, last] of the angle (libm)."""
    a, b = ct.c_uint64(), ct.c_uint64()
    lib().orc_trig_checksum(ct.c_uint32(first_bits), ct.c_uint32(last_bits), ct.byref(a), ct.byref(b), int(nthreads))
    return a.value, b.value


def cosf(x):
    return float(lib().orc_cosf(ct.c_float(x)))


def sinf(x):
    lib().orc_sinf.restype = ct.c_float
    lib().orc_sinf.argtypes = [ct.c_float]
    return float(lib().orc_sinf(ct.c_float(x)))


def introsort_pairs(cnt, ulx):
    n = len(cnt)
    c = np.array(cnt, np.int32)
    u = np.array(ulx, np.int32)
    p = np.arange(n, dtype=np.int32)
    lib().orc_introsort_pairs(_p(c), _p(u), _p(p), n)
    return list(p)
