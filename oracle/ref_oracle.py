"""ref_oracle.py — ctypes binding of oracle/_ref/libref_orbextractor.so: the REFERENCE's own ORBextractor.cpp, compiled
unmodified against the header shim (oracle/ref_shim, recipe: `make -C oracle _ref`).

ORACLE = test infrastructure.  Only tests/, __graft_entry__ and bench.py's CPU legs import this module; the product never does.
`available()` is False where neither the prebuilt library nor /root/reference exists.
"""
import ctypes as ct
import os
import subprocess

import numpy as np

from c_oracle import CAND_DTYPE, KP_DTYPE, _p

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_orbextractor.so")
REF_SRC = "/root/reference/dynamic_visual_slam/src/ORBextractor.cpp"
_lib = None


def build():
    """(Re)build when the reference sources are present (this container); on the GPU box the prebuilt file is used."""
    if os.path.exists(REF_SRC):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_ref"])
    return SO if os.path.exists(SO) else None


def available():
    return os.path.exists(SO) or os.path.exists(REF_SRC)


def lib():
    global _lib
    if _lib is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _lib = ct.CDLL(so)
        _lib.ref_create.restype = ct.c_void_p
        _lib.ref_create.argtypes = [ct.c_int, ct.c_float, ct.c_int, ct.c_int, ct.c_int]
        _lib.ref_destroy.argtypes = [ct.c_void_p]
        _lib.ref_source_path.restype = ct.c_char_p
    return _lib


class RefExtractor:
    """ORB_SLAM3::ORBextractor of the reference (ORBextractor.hpp:44-111), same constructor arguments."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7):
        self.args = (nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        self.nlevels = nlevels
        self.h = ct.c_void_p(lib().ref_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        f = [np.zeros(n, np.float32) for _ in range(4)]
        nf = np.zeros(n, np.int32)
        um = np.zeros(16, np.int32)
        lib().ref_tables(self.h, _p(f[0]), _p(f[1]), _p(f[2]), _p(f[3]), _p(nf), _p(um))
        return dict(scale=f[0], inv_scale=f[1], sigma2=f[2], inv_sigma2=f[3], nfeat=nf, umax=um)

    def extract(self, gray, lapping=(0, 0), cap=20000):
        """operator(): returns dict(kps, desc, ret) — `ret` is the reference's return value (monoIndex)."""
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        h, w = gray.shape if gray.ndim == 2 else (0, 0)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = ct.c_int(0)
        r = lib().ref_extract(self.h, _p(gray) if gray.size else None, w, h, ct.c_size_t(gray.strides[0] if gray.size else 0),
                              int(lapping[0]), int(lapping[1]), _p(kps), _p(desc), cap, ct.byref(n))
        if r == -2:
            raise RuntimeError("capacity")
        return dict(kps=kps[:n.value].copy(), desc=desc[:n.value].copy(), ret=r)

    def level(self, l, padded=False):
        w, h = ct.c_int(), ct.c_int()
        f = lib().ref_get_level_padded if padded else lib().ref_get_level
        if f(self.h, l, None, ct.byref(w), ct.byref(h)) != 0:
            raise IndexError(l)
        out = np.zeros((h.value, w.value), np.uint8)
        f(self.h, l, _p(out), ct.byref(w), ct.byref(h))
        return out

    def stage_trace(self, gray, cap=1 << 17):
        """ComputePyramid + ComputeKeyPointsOctTree with each cv::FAST call observed."""
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        h, w = gray.shape
        n = self.nlevels
        cands = np.zeros((n, cap), CAND_DTYPE)
        keys = np.zeros((n, cap), KP_DTYPE)
        nc, nk = np.zeros(n, np.int32), np.zeros(n, np.int32)
        ci, cm = np.zeros(n, np.int32), np.zeros(n, np.int32)
        rc = lib().ref_stage_trace(self.h, _p(gray), w, h, ct.c_size_t(gray.strides[0]), _p(cands), _p(nc), _p(keys), _p(nk), cap, _p(ci), _p(cm))
        if rc != 0:
            raise RuntimeError("ref_stage_trace: %d" % rc)
        return dict(cands=[cands[l, :nc[l]].copy() for l in range(n)], keys=[keys[l, :nk[l]].copy() for l in range(n)],
                    pyramid=[self.level(l) for l in range(n)], calls_ini=ci, calls_min=cm)

    def distribute_octtree(self, cands, minX, maxX, minY, maxY, N):
        cands = np.ascontiguousarray(cands, dtype=CAND_DTYPE)
        out = np.zeros(max(len(cands), 1) + 8, CAND_DTYPE)
        n = lib().ref_distribute_octtree(self.h, _p(cands), len(cands), minX, maxX, minY, maxY, N, _p(out), len(out))
        if n < 0:
            raise RuntimeError("ref_distribute_octtree: %d" % n)
        return out[:n].copy()

    def keypoints_old(self, gray, cap=1 << 15):
        gray = np.ascontiguousarray(gray, dtype=np.uint8)
        h, w = gray.shape
        keys = np.zeros((self.nlevels, cap), KP_DTYPE)
        nk = np.zeros(self.nlevels, np.int32)
        rc = lib().ref_keypoints_old(self.h, _p(gray), w, h, ct.c_size_t(gray.strides[0]), _p(keys), _p(nk), cap)
        if rc != 0:
            raise RuntimeError("ref_keypoints_old: %d" % rc)
        return [keys[l, :nk[l]].copy() for l in range(self.nlevels)]


def extract_batch(frames, args=(1000, 1.2, 8, 20, 7), cap=4096, nthreads=0):
    """Frame-parallel CPU baseline: one reference extractor per OpenMP thread."""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    nf, h, w = frames.shape
    kps = np.zeros((nf, cap), KP_DTYPE)
    desc = np.zeros((nf, cap, 32), np.uint8)
    counts = np.zeros(nf, np.int32)
    rc = lib().ref_extract_batch(int(args[0]), ct.c_float(args[1]), int(args[2]), int(args[3]), int(args[4]),
                                 _p(frames), nf, w, h, _p(kps), _p(desc), cap, _p(counts), int(nthreads))
    if rc != 0:
        raise RuntimeError("ref_extract_batch failed")
    return kps, desc, counts
