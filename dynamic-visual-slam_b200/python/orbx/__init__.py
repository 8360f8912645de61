"""orbx — thin ctypes binding of the C ABI in include/orbx.h (tests and bench harness only).

The product is liborbx.so (hand-written sm_100a kernels behind a C ABI); this module only marshals
numpy arrays / raw device pointers into it.  There is NO CPU fallback: if the shared library is
missing or no CUDA device is present, construction fails loudly.

Mirrors of the reference call shapes (paths relative to the reference's dynamic_visual_slam/):
  ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)   include/.../ORBextractor.hpp:50-51
  ORBextractor.__call__(image)  -> (keypoints, descriptors)             ORBextractor.hpp:58-60
  BFMatcher.match(query, train) -> DMatch[]                             frontend.cpp:1123
"""
import ctypes as ct
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ORBX_LIB selects another build of the same library (kernel experiments: csrc/Makefile variant builds); never a fallback
LIB_PATH = os.environ.get("ORBX_LIB") or os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "liborbx.so"))

OK, E_INVALID, E_CUDA, E_CAPACITY, E_EMPTY, E_NOMEM, E_UNSUPPORTED = range(7)
MAX_LEVELS = 16

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
DM_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
BOX_DTYPE = np.dtype([("cx", "<f8"), ("cy", "<f8"), ("w", "<f8"), ("h", "<f8"), ("class_id", "<i4"), ("pad", "<i4")])
TOP2_DTYPE = np.dtype([("dist0", "<u4"), ("idx0", "<u4"), ("dist1", "<u4"), ("idx1", "<u4")])
ASSOC_DTYPE = np.dtype([("reproj_error", "<f8"), ("landmark", "<i4"), ("distance", "<f4")])
POSE_DTYPE = np.dtype([("R", "<f8", (9,)), ("t", "<f8", (3,)), ("fx", "<f8"), ("fy", "<f8"), ("cx", "<f8"), ("cy", "<f8")])
assert KP_DTYPE.itemsize == 28 and DM_DTYPE.itemsize == 16 and BOX_DTYPE.itemsize == 40 and TOP2_DTYPE.itemsize == 16
KFPARAMS_DTYPE = np.dtype([("R", "<f8", (9,)), ("t", "<f8", (3,)), ("fx", "<f4"), ("fy", "<f4"), ("cx", "<f4"), ("cy", "<f4")])
KF_DTYPE = np.dtype([("landmark_id", "<u8"), ("position", "<f8", (3,)), ("pixel_x", "<f8"), ("pixel_y", "<f8"), ("descriptor", "u1", (32,))])
assert ASSOC_DTYPE.itemsize == 16 and POSE_DTYPE.itemsize == 128 and KFPARAMS_DTYPE.itemsize == 112 and KF_DTYPE.itemsize == 80


class Params(ct.Structure):
    _fields_ = [("nfeatures", ct.c_int32), ("scale_factor", ct.c_float), ("nlevels", ct.c_int32),
                ("ini_th_fast", ct.c_int32), ("min_th_fast", ct.c_int32),
                ("depth_min", ct.c_float), ("depth_max", ct.c_float),
                ("max_width", ct.c_int32), ("max_height", ct.c_int32), ("max_batch", ct.c_int32),
                ("max_keypoints", ct.c_int32), ("cand_divisor", ct.c_int32), ("device", ct.c_int32),
                ("host_chunk", ct.c_int32), ("profile", ct.c_int32), ("reserved_", ct.c_int32)]


class OrbxError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("orbx status %d: %s" % (status, msg))
        self.status = status


_lib = None


def load():
    """Load liborbx.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("liborbx.so not built at %s — run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    L = ct.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, f32, sz = ct.c_void_p, ct.c_int32, ct.c_int64, ct.c_uint32, ct.c_uint64, ct.c_float, ct.c_size_t
    sig = {
        "orbx_default_params": (None, [vp]),
        "orbx_create": (i32, [vp, vp]),
        "orbx_destroy": (None, [vp]),
        "orbx_last_error": (ct.c_char_p, [vp]),
        "orbx_version": (ct.c_char_p, []),
        "orbx_sync": (i32, [vp]),
        "orbx_set_option": (i32, [vp, i32, i32]),
        "orbx_cull_keyframe": (i32, [vp, vp, vp, i32, vp, i32, i32, f32, vp, vp, vp, i32, vp]),
        "orbx_cull_keyframe_device": (i32, [vp, vp, vp, i32, vp, i32, i32, f32, vp, vp, vp, i32, vp]),
        "orbx_stream": (vp, [vp]),
        "orbx_get_levels": (i32, [vp]),
        "orbx_get_scale_factor": (f32, [vp]),
        "orbx_get_scale_factors": (None, [vp, vp]),
        "orbx_get_inverse_scale_factors": (None, [vp, vp]),
        "orbx_get_scale_sigma_squares": (None, [vp, vp]),
        "orbx_get_inverse_scale_sigma_squares": (None, [vp, vp]),
        "orbx_get_features_per_level": (None, [vp, vp]),
        "orbx_level_size": (i32, [vp, i32, i32, i32, vp, vp]),
        "orbx_extract": (i32, [vp, vp, i32, i32, sz, vp, vp, i32, vp]),
        "orbx_extract_filtered": (i32, [vp, vp, i32, i32, sz, vp, sz, vp, i32, u64, vp, vp, i32, vp]),
        "orbx_extract_batch": (i32, [vp, vp, i32, i32, i32, sz, vp, sz, vp, vp, i32, vp]),
        "orbx_extract_bgr": (i32, [vp, vp, i32, i32, sz, vp, sz, vp, i32, u64, vp, vp, i32, vp]),
        "orbx_bgr2gray_device": (i32, [vp, vp, i32, i32, i32, sz, sz, vp, sz, sz]),
        "orbx_extract_batch_device": (i32, [vp, vp, i32, i32, i32, sz, sz, vp, sz, sz, vp, vp, i32, vp]),
        "orbx_match": (i32, [vp, vp, i32, vp, i32, i32, f32, f32, vp, vp]),
        "orbx_match_device": (i32, [vp, vp, i32, vp, i32, i32, f32, f32, vp, vp]),
        "orbx_match_pairs_device": (i32, [vp, vp, vp, i32, vp, vp, i32, i32, f32, f32, vp, vp]),
        "orbx_db_create": (i32, [vp, i64, u32, vp]),
        "orbx_db_destroy": (None, [vp]),
        "orbx_db_append": (i32, [vp, vp, i64]),
        "orbx_db_append_device": (i32, [vp, vp, i64]),
        "orbx_db_rows": (i64, [vp]),
        "orbx_db_query_top2_device": (i32, [vp, vp, i32, vp]),
        "orbx_db_query_top2": (i32, [vp, vp, i32, vp]),
        "orbx_merge_top2_device": (i32, [vp, vp, i32, i32, vp]),
        "orbx_db_query_radius": (i32, [vp, vp, i32, f32, vp, i32, vp]),
        "orbx_pack_keyframe": (i32, [vp, vp, vp, i32, vp, i32, i32, sz, vp, vp, i32, vp]),
        "orbx_pack_keyframe_device": (i32, [vp, i32, vp, vp, vp, i32, vp, i32, i32, sz, sz, vp, vp, vp, i32]),
        "orbx_db_set_positions": (i32, [vp, i64, i64, vp]),
        "orbx_db_set_positions_device": (i32, [vp, i64, i64, vp]),
        "orbx_db_associate": (i32, [vp, vp, vp, i32, vp, f32, ct.c_double, vp]),
        "orbx_db_associate_device": (i32, [vp, vp, vp, i32, vp, f32, ct.c_double, vp]),
        "orbx_merge_assoc_device": (i32, [vp, vp, i32, i32, vp]),
        "orbx_get_pyramid_level": (i32, [vp, i32, i32, vp, sz]),
        "orbx_get_blurred_level": (i32, [vp, i32, i32, vp, sz]),
        "orbx_get_candidates": (i32, [vp, i32, i32, vp, i32, vp]),
        "orbx_get_fast_scores": (i32, [vp, i32, i32, vp, ct.c_size_t]),
        "orbx_get_fast_edge_corners": (i32, [vp, i32, i32, vp, i32, vp]),
        "orbx_get_level_counts": (i32, [vp, i32, vp]),
        "orbx_harris_responses": (i32, [vp, i32, i32, vp, i32, i32, f32, vp]),
        "orbx_synth_gray_device": (i32, [vp, u32, i32, i32, i32, i32, vp, sz, sz]),
        "orbx_synth_depth_device": (i32, [vp, u32, i32, i32, i32, i32, vp, sz, sz]),
        "orbx_synth_descriptors_device": (i32, [vp, u32, u64, i64, vp]),
        "orbx_alloc_pinned": (vp, [sz]),
        "orbx_free_pinned": (None, [vp]),
        "orbx_alloc_device": (vp, [vp, sz]),
        "orbx_free_device": (None, [vp, vp]),
        "orbx_copy_to_device": (i32, [vp, vp, vp, sz]),
        "orbx_copy_to_host": (i32, [vp, vp, vp, sz]),
        "orbx_test_trig": (i32, [vp, vp, i32, vp, vp]),
        "orbx_test_atan2": (i32, [vp, vp, vp, i32, vp]),
        "orbx_test_trig_checksum": (i32, [vp, u32, u32, vp, vp]),
        "orbx_test_quadtree": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
        "orbx_bench_popc": (i32, [vp, vp]),
        "orbx_launch_count": (i64, [vp]),
        "orbx_track_batch_device": (i32, [vp, vp, i32, i32, i32, sz, sz, vp, sz, sz, vp, vp, i32, vp, vp, vp, f32]),
        "orbx_track_batch": (i32, [vp, vp, i32, i32, i32, sz, vp, sz, vp, vp, i32, vp, vp, vp, f32]),
        "orbx_track_reset": (None, [vp]),
        "orbx_extract_batch_submit": (i32, [vp, vp, i32, i32, i32, sz, vp, sz, vp, vp, i32, vp, vp]),
        "orbx_track_batch_submit": (i32, [vp, vp, i32, i32, i32, sz, vp, sz, vp, vp, i32, vp, vp, vp, f32, vp]),
        "orbx_batch_wait": (i32, [vp, i32]),
        "orbx_fmat_score": (i32, [vp, vp, vp, i32, vp, i32, ct.c_double, vp, vp, vp]),
        "orbx_pnp_points": (i32, [vp, vp, i32, vp, i32, vp, i32, vp, i32, i32, ct.c_size_t, ct.c_float, ct.c_float, ct.c_float, ct.c_float, vp, vp, vp]),
        "orbx_pnp_score": (i32, [vp, vp, vp, i32, vp, i32, ct.c_double, ct.c_double, ct.c_double, ct.c_double, ct.c_double, vp, vp, vp]),
        "orbx_fmat_ransac": (i32, [vp, vp, vp, i32, i32, ct.c_double, u32, vp, vp, vp]),
        "orbx_comm_get_unique_id": (i32, [vp]),
        "orbx_comm_last_error": (ct.c_char_p, []),
        "orbx_comm_create": (i32, [vp, i32, i32, vp, vp]),
        "orbx_comm_destroy": (None, [vp]),
        "orbx_comm_ranks": (i32, [vp]),
        "orbx_comm_set_transport": (i32, [vp, i32]),
        "orbx_comm_peer_memory": (i32, [vp]),
        "orbx_comm_rank": (i32, [vp]),
        "orbx_db_query_top2_sharded_device": (i32, [vp, vp, vp, i32, vp]),
        "orbx_db_associate_sharded_device": (i32, [vp, vp, vp, vp, i32, vp, f32, ct.c_double, vp]),
        "orbx_extract_batch_boxes_device": (i32, [vp, vp, i32, i32, i32, sz, sz, vp, sz, sz, vp, vp, i32, u64, vp, vp, i32, vp]),
        "orbx_extract_batch_boxes": (i32, [vp, vp, i32, i32, i32, sz, vp, sz, vp, vp, u64, vp, vp, i32, vp]),
        "orbx_track_batch_boxes_device": (i32, [vp, vp, i32, i32, i32, sz, sz, vp, sz, sz, vp, vp, i32, u64, vp, vp, i32, vp, vp, vp, f32]),
        "orbx_track_batch_boxes": (i32, [vp, vp, i32, i32, i32, sz, vp, sz, vp, vp, u64, vp, vp, i32, vp, vp, vp, f32]),
        "orbx_track_batch_boxes_submit": (i32, [vp, vp, i32, i32, i32, sz, vp, sz, vp, vp, u64, vp, vp, i32, vp, vp, vp, f32, vp]),
        "orbx_profile_enable": (None, [vp, i32]),
        "orbx_profile_kernels": (i32, []),
        "orbx_profile_name": (ct.c_char_p, [i32]),
        "orbx_profile_read": (i32, [vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


ABI_SYMBOLS = None  # filled by tests from include/orbx.h


def _p(a):
    return a.ctypes.data_as(ct.c_void_p) if a is not None else None


class PinnedArray:
    """numpy view of page-locked host memory from orbx_alloc_pinned (freed on close / garbage collection).
    Pinned inputs are DMA'd without a staging copy, and pinned depth maps are read in place by the GPU."""

    def __init__(self, shape, dtype):
        L = load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = L.orbx_alloc_pinned(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError("orbx_alloc_pinned failed")
        buf = (ct.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self):
        if self.ptr:
            self.array = None
            load().orbx_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ORBextractor:
    """ORB_SLAM3::ORBextractor over the C ABI (reference ORBextractor.hpp:44-111)."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7,
                 max_width=1280, max_height=720, max_batch=1, device=0, depth_min=0.3, depth_max=3.0,
                 max_keypoints=0, cand_divisor=0, host_chunk=0, profile=0):
        L = load()
        p = Params()
        L.orbx_default_params(ct.byref(p))
        p.nfeatures, p.scale_factor, p.nlevels = nfeatures, scaleFactor, nlevels
        p.ini_th_fast, p.min_th_fast = iniThFAST, minThFAST
        p.max_width, p.max_height, p.max_batch, p.device = max_width, max_height, max_batch, device
        p.depth_min, p.depth_max, p.max_keypoints, p.cand_divisor = depth_min, depth_max, max_keypoints, cand_divisor
        p.host_chunk = host_chunk
        p.profile = {"slam": 0, "cvorb": 1}.get(profile, profile)
        self._h = ct.c_void_p()
        st = L.orbx_create(ct.byref(p), ct.byref(self._h))
        if st != OK:
            raise OrbxError(st, (L.orbx_last_error(None) or b"").decode())
        self.L = L
        self.params = p
        self.nlevels = nlevels
        self.max_batch = max_batch

    # -- lifetime --
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.L.orbx_destroy(self._h)
            self._h = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != OK:
            raise OrbxError(st, (self.L.orbx_last_error(self._h) or b"").decode())

    @property
    def handle(self):
        return self._h

    def sync(self):
        self._check(self.L.orbx_sync(self._h))

    def set_serial(self, on=True):
        """ORBX_OPT_SERIAL: one stream, kernels in order (isolated per-kernel timings) instead of the overlapped schedule."""
        self._check(self.L.orbx_set_option(self._h, 1, 1 if on else 0))

    def set_fused_blur(self, on=True):
        """ORBX_OPT_FUSED_BLUR: evaluate the Gaussian inside the descriptor kernel (default) instead of blurring every level first."""
        self._check(self.L.orbx_set_option(self._h, 3, 1 if on else 0))

    def set_pdl(self, on=True):
        """ORBX_OPT_PDL: programmatic dependent launch of the step's kernel chain (default on)."""
        self._check(self.L.orbx_set_option(self._h, 4, 1 if on else 0))

    def set_overlap(self, on=True):
        """ORBX_OPT_OVERLAP: batches of >= 32 frames as two staggered half-batches on two streams (default off: measured slower)."""
        self._check(self.L.orbx_set_option(self._h, 5, 1 if on else 0))

    def set_match_mma(self, on=True):
        """ORBX_OPT_MATCH_MMA: 1 = tensor-memory (tcgen05) matcher for calls of >= 8 M pairs (default), 3 = for every call, 2 = the mma.sync int8 matcher,
        0 / False = always the POPC kernel."""
        self._check(self.L.orbx_set_option(self._h, 6, int(on)))

    def set_fast_dense(self, mode=1):
        """ORBX_OPT_FAST_DENSE: 0 = warp-per-cell FAST kernel (default), 1 = dense formulation for batches of >= 8 frames, 2 = for every call,
        3 = for every call with the NMS inside the tile kernel."""
        self._check(self.L.orbx_set_option(self._h, 7, int(mode)))

    def set_filter_first(self, on=True):
        """ORBX_OPT_FILTER_FIRST: 1 (default) = the depth / box filter runs on the selected positions before the descriptor kernel, which then
        describes the survivors only; 0 = the reference's order (describe everything, then drop rows).  Same output."""
        self._check(self.L.orbx_set_option(self._h, 8, 1 if on else 0))

    def set_fast_ctas(self, n):
        """ORBX_OPT_FAST_CTAS: resident FAST warps per SM in the overlapped schedule (0 = as many as fit)."""
        self._check(self.L.orbx_set_option(self._h, 2, int(n)))

    @property
    def stream(self):
        return self.L.orbx_stream(self._h)

    @property
    def launch_count(self):
        return int(self.L.orbx_launch_count(self._h))

    # -- getters (ORBextractor.hpp:62-82) --
    def GetLevels(self):
        return int(self.L.orbx_get_levels(self._h))

    def GetScaleFactor(self):
        return float(self.L.orbx_get_scale_factor(self._h))

    def _vec(self, fn, dtype=np.float32):
        out = np.zeros(self.nlevels, dtype)
        fn(self._h, _p(out))
        return out

    def GetScaleFactors(self):
        return self._vec(self.L.orbx_get_scale_factors)

    def GetInverseScaleFactors(self):
        return self._vec(self.L.orbx_get_inverse_scale_factors)

    def GetScaleSigmaSquares(self):
        return self._vec(self.L.orbx_get_scale_sigma_squares)

    def GetInverseScaleSigmaSquares(self):
        return self._vec(self.L.orbx_get_inverse_scale_sigma_squares)

    def features_per_level(self):
        return self._vec(self.L.orbx_get_features_per_level, np.int32)

    def level_size(self, w, h, level):
        lw, lh = ct.c_int32(), ct.c_int32()
        self._check(self.L.orbx_level_size(self._h, w, h, level, ct.byref(lw), ct.byref(lh)))
        return lw.value, lh.value

    # -- operator() --
    def __call__(self, image, depth=None, boxes=None, drop_class_mask=0, cap=4096):
        """Returns (keypoints[KP_DTYPE], descriptors[N,32] uint8).  Returns -1 for an empty image,
        as the reference's operator() does (ORBextractor.cpp:1090-1091)."""
        if image is None or image.size == 0:
            return -1
        if image.dtype != np.uint8 or image.ndim != 2:
            raise TypeError("image must be CV_8UC1")          # reference: assert(image.type() == CV_8UC1)
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        self._last_w, self._last_h = w, h
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = ct.c_int32()
        if depth is None and boxes is None:
            st = self.L.orbx_extract(self._h, _p(image), w, h, image.strides[0], _p(kps), _p(desc), cap, ct.byref(n))
        else:
            dptr, dstep = None, 0
            if depth is not None:
                if depth.dtype != np.uint16 or depth.shape != image.shape:
                    raise TypeError("depth must be CV_16UC1 of the image size")
                if depth.strides[1] != 2:
                    depth = np.ascontiguousarray(depth)
                dptr, dstep = _p(depth), depth.strides[0]
            bptr, nb = None, 0
            if boxes is not None and len(boxes):
                boxes = np.ascontiguousarray(boxes, dtype=BOX_DTYPE)
                bptr, nb = _p(boxes), len(boxes)
            st = self.L.orbx_extract_filtered(self._h, _p(image), w, h, image.strides[0], dptr, dstep, bptr, nb,
                                              ct.c_uint64(drop_class_mask), _p(kps), _p(desc), cap, ct.byref(n))
        self._check(st)
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_bgr(self, bgr, depth=None, cap=4096):
        """cvtColor(BGR2GRAY) on the device + operator() (+ filterDepth): the frontend's ingest, reference frontend.cpp:1084-1100."""
        bgr = np.ascontiguousarray(bgr, np.uint8)
        h, w, _ = bgr.shape
        self._last_w, self._last_h = w, h
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = ct.c_int32()
        dptr, dstep = None, 0
        if depth is not None:
            depth = np.ascontiguousarray(depth, np.uint16)
            dptr, dstep = _p(depth), depth.strides[0]
        self._check(self.L.orbx_extract_bgr(self._h, _p(bgr), w, h, bgr.strides[0], dptr, dstep, None, 0, ct.c_uint64(0),
                                            _p(kps), _p(desc), cap, ct.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def pack_keyframe(self, kps, desc, depth, fx, fy, cx, cy, R, t):
        """Landmark / observation records of Frontend::publishKeyframe (reference frontend.cpp:731-776)."""
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        depth = np.ascontiguousarray(depth, np.uint16)
        K = np.zeros(1, KFPARAMS_DTYPE)
        K["R"][0] = np.asarray(R, np.float64).reshape(9)
        K["t"][0] = np.asarray(t, np.float64).reshape(3)
        K["fx"], K["fy"], K["cx"], K["cy"] = fx, fy, cx, cy
        out = np.zeros(max(len(kps), 1), KF_DTYPE)
        n = ct.c_int32()
        self._check(self.L.orbx_pack_keyframe(self._h, _p(kps), _p(desc), len(kps), _p(depth), depth.shape[1], depth.shape[0], depth.strides[0],
                                              _p(K), _p(out), len(out), ct.byref(n)))
        return out[:n.value].copy()

    def cull_keyframe(self, kps, desc, match_query, max_new=200, min_response=50.0, cap=None):
        """Feature culling for the backend (reference frontend.cpp:1168-1218): (kps, desc, index) — matched keypoints in match order, then
        the best unmatched ones by response in the reference's std::sort order."""
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        q = np.ascontiguousarray(match_query, np.int32)
        cap = len(q) + min(int(max_new), len(kps)) if cap is None else int(cap)
        ok, od, oi = np.zeros(max(cap, 1), KP_DTYPE), np.zeros((max(cap, 1), 32), np.uint8), np.zeros(max(cap, 1), np.int32)
        n = ct.c_int32()
        self._check(self.L.orbx_cull_keyframe(self._h, _p(kps), _p(desc), len(kps), _p(q), len(q), int(max_new), ct.c_float(min_response),
                                              _p(ok), _p(od), _p(oi), cap, ct.byref(n)))
        return ok[:n.value].copy(), od[:n.value].copy(), oi[:n.value].copy()

    @staticmethod
    def pack_frame_boxes(frame_boxes):
        """list (one entry per frame) of BOX_DTYPE arrays -> (boxes, offsets[nframes + 1]) for the *_boxes batch calls"""
        off = np.zeros(len(frame_boxes) + 1, np.int32)
        off[1:] = np.cumsum([len(b) for b in frame_boxes])
        parts = [np.ascontiguousarray(b, BOX_DTYPE) for b in frame_boxes if len(b)]
        boxes = np.concatenate(parts) if parts else np.zeros(0, BOX_DTYPE)
        return np.ascontiguousarray(boxes), off

    def extract_batch(self, frames, depth=None, cap=2048, frame_boxes=None, drop_class_mask=0):
        """frame_boxes: per-frame YOLO box lists (BASELINE configs[4]); keypoints inside a box of a masked class are dropped after selection"""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        nf, h, w = frames.shape
        self._last_w, self._last_h = w, h
        kps = np.zeros((nf, cap), KP_DTYPE)
        desc = np.zeros((nf, cap, 32), np.uint8)
        counts = np.zeros(nf, np.int32)
        dptr, dstep = None, 0
        if depth is not None:
            depth = np.ascontiguousarray(depth, dtype=np.uint16)
            dptr, dstep = _p(depth), depth.strides[1]
        if frame_boxes is None:
            self._check(self.L.orbx_extract_batch(self._h, _p(frames), nf, w, h, frames.strides[1], dptr, dstep,
                                                  _p(kps), _p(desc), cap, _p(counts)))
        else:
            boxes, off = self.pack_frame_boxes(frame_boxes)
            self._check(self.L.orbx_extract_batch_boxes(self._h, _p(frames), nf, w, h, frames.strides[1], dptr, dstep,
                                                        _p(boxes), _p(off), ct.c_uint64(drop_class_mask), _p(kps), _p(desc), cap, _p(counts)))
        return kps, desc, counts

    def pnp_points(self, prev_kps, curr_kps, matches, prev_depth, fx, fy, cx, cy):
        """The 3D-2D correspondences of Frontend::estimateCameraPose (frontend.cpp:858-892): (points3d [n,3], points2d [n,2]) in match order."""
        prev_kps = np.ascontiguousarray(prev_kps, KP_DTYPE); curr_kps = np.ascontiguousarray(curr_kps, KP_DTYPE)
        matches = np.ascontiguousarray(matches, DM_DTYPE); depth = np.ascontiguousarray(prev_depth, np.uint16)
        nm = len(matches)
        p3 = np.zeros((max(nm, 1), 3), np.float32); p2 = np.zeros((max(nm, 1), 2), np.float32)
        n = ct.c_int32()
        self._check(self.L.orbx_pnp_points(self._h, _p(prev_kps), len(prev_kps), _p(curr_kps), len(curr_kps), _p(matches), nm, _p(depth), depth.shape[1], depth.shape[0],
                                           depth.strides[0], ct.c_float(fx), ct.c_float(fy), ct.c_float(cx), ct.c_float(cy), _p(p3), _p(p2), ct.byref(n)))
        return p3[:n.value].copy(), p2[:n.value].copy()

    def pnp_score(self, pts3d, pts2d, R, t, fx, fy, cx, cy, threshold=4.0):
        """Inlier counts of pose hypotheses (R [nh,3,3], t [nh,3]) under cv::solvePnPRansac's error (frontend.cpp:911-923);
        returns (counts[nh], best index, mask[n] of the best)."""
        pts3d = np.ascontiguousarray(pts3d, np.float32).reshape(-1, 3); pts2d = np.ascontiguousarray(pts2d, np.float32).reshape(-1, 2)
        R = np.ascontiguousarray(R, np.float64).reshape(-1, 9); t = np.ascontiguousarray(t, np.float64).reshape(-1, 3)
        Rt = np.ascontiguousarray(np.concatenate([R, t], 1))
        counts = np.zeros(len(Rt), np.int32); mask = np.zeros(max(len(pts3d), 1), np.uint8); best = ct.c_int32()
        self._check(self.L.orbx_pnp_score(self._h, _p(pts3d), _p(pts2d), len(pts3d), _p(Rt), len(Rt), ct.c_double(fx), ct.c_double(fy), ct.c_double(cx), ct.c_double(cy),
                                          ct.c_double(threshold), _p(counts), ct.byref(best), _p(mask)))
        return counts, best.value, mask[:len(pts3d)]

    def fmat_score(self, pts1, pts2, F, threshold=2.0):
        """Inlier counts of fundamental-matrix hypotheses F [nh, 3, 3] under OpenCV's RANSAC error (frontend.cpp:1134-1154);
        returns (counts[nh], best index, mask[n] of the best)."""
        pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
        F = np.ascontiguousarray(F, np.float64).reshape(-1, 9)
        counts = np.zeros(len(F), np.int32)
        mask = np.zeros(max(len(pts1), 1), np.uint8)
        best = ct.c_int32()
        self._check(self.L.orbx_fmat_score(self._h, _p(pts1), _p(pts2), len(pts1), _p(F), len(F), ct.c_double(threshold), _p(counts), ct.byref(best), _p(mask)))
        return counts, best.value, mask[:len(pts1)].copy()

    def fmat_ransac(self, pts1, pts2, iters=1000, threshold=2.0, seed=1):
        """Device RANSAC for the fundamental matrix (8-point hypotheses + OpenCV's scoring): returns (F [3,3], mask [n], inliers)."""
        pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
        F = np.zeros(9, np.float64)
        mask = np.zeros(len(pts1), np.uint8)
        n = ct.c_int32()
        self._check(self.L.orbx_fmat_ransac(self._h, _p(pts1), _p(pts2), len(pts1), int(iters), ct.c_double(threshold), ct.c_uint32(seed), _p(F), _p(mask), ct.byref(n)))
        return F.reshape(3, 3), mask, n.value

    def extract_batch_device(self, d_gray, nframes, w, h, step, frame_stride, d_kps, d_desc, cap, d_counts,
                             d_depth=None, dstep=0, dframe_stride=0):
        """Raw device-pointer variant (ints), asynchronous on the handle's stream."""
        self._check(self.L.orbx_extract_batch_device(self._h, d_gray, nframes, w, h, step, frame_stride,
                                                     d_depth, dstep, dframe_stride, d_kps, d_desc, cap, d_counts))

    # -- stream step (hot part of Frontend::syncCallback) --
    def track_reset(self):
        self.L.orbx_track_reset(self._h)

    def track_batch(self, frames, depth=None, cap=2048, max_dist=50.0, frame_boxes=None, drop_class_mask=0):
        """frames [n,h,w] u8 (+ depth [n,h,w] u16, + per-frame YOLO boxes): per-frame filtered keypoints/descriptors and matches vs the previous frame."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        nf, h, w = frames.shape
        self._last_w, self._last_h = w, h
        kps = np.zeros((nf, cap), KP_DTYPE)
        desc = np.zeros((nf, cap, 32), np.uint8)
        counts = np.zeros(nf, np.int32)
        matches = np.zeros((nf, cap), DM_DTYPE)
        mcounts = np.zeros(nf, np.int32)
        dptr, dstep = None, 0
        if depth is not None:
            depth = np.ascontiguousarray(depth, dtype=np.uint16)
            dptr, dstep = _p(depth), depth.strides[1]
        if frame_boxes is None:
            self._check(self.L.orbx_track_batch(self._h, _p(frames), nf, w, h, frames.strides[1], dptr, dstep,
                                                _p(kps), _p(desc), cap, _p(counts), _p(matches), _p(mcounts), ct.c_float(max_dist)))
        else:
            boxes, off = self.pack_frame_boxes(frame_boxes)
            self._check(self.L.orbx_track_batch_boxes(self._h, _p(frames), nf, w, h, frames.strides[1], dptr, dstep,
                                                      _p(boxes), _p(off), ct.c_uint64(drop_class_mask),
                                                      _p(kps), _p(desc), cap, _p(counts), _p(matches), _p(mcounts), ct.c_float(max_dist)))
        return kps, desc, counts, matches, mcounts

    def track_batch_submit(self, frames, depth, out, max_dist=50.0):
        """Asynchronous track_batch: `frames` [n,h,w] u8 / `depth` [n,h,w] u16 (pinned for real overlap) and `out` = dict of
        pre-allocated arrays kps [n,cap] KP_DTYPE, desc [n,cap,32], counts [n], matches [n,cap] DM_DTYPE, mcounts [n].
        Returns a ticket for batch_wait(); at most two batches may be in flight."""
        nf, h, w = frames.shape
        cap = out["kps"].shape[1]
        t = ct.c_int32()
        self._check(self.L.orbx_track_batch_submit(self._h, _p(frames), nf, w, h, frames.strides[1],
                                                   _p(depth) if depth is not None else None, depth.strides[1] if depth is not None else 0,
                                                   _p(out["kps"]), _p(out["desc"]), cap, _p(out["counts"]), _p(out["matches"]), _p(out["mcounts"]),
                                                   ct.c_float(max_dist), ct.byref(t)))
        return t.value

    def batch_wait(self, ticket):
        self._check(self.L.orbx_batch_wait(self._h, ticket))

    def profile_enable(self, on=True):
        self.L.orbx_profile_enable(self._h, 1 if on else 0)

    def profile_read(self):
        n = self.L.orbx_profile_kernels()
        ms = np.zeros(n, np.float64)
        cnt = np.zeros(n, np.int64)
        self._check(self.L.orbx_profile_read(self._h, _p(ms), _p(cnt)))
        return {self.L.orbx_profile_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    # -- stage access --
    def pyramid_level(self, level, frame=0):
        w, h = self.level_size(self._last_w, self._last_h, level)
        out = np.zeros((h, w), np.uint8)
        self._check(self.L.orbx_get_pyramid_level(self._h, frame, level, _p(out), out.strides[0]))
        return out

    def blurred_level(self, level, frame=0):
        w, h = self.level_size(self._last_w, self._last_h, level)
        out = np.zeros((h, w), np.uint8)
        self._check(self.L.orbx_get_blurred_level(self._h, frame, level, _p(out), out.strides[0]))
        return out

    def fast_scores(self, level, frame=0):
        """iniThFAST score map of a level after a dense FAST run: [h - 38, w - 38], level pixel (19 + c, 19 + r) at [r, c]."""
        w, h = self.level_size(self._last_w, self._last_h, level)
        out = np.zeros((max(h - 38, 1), max(w - 38, 1)), np.uint8)
        self._check(self.L.orbx_get_fast_scores(self._h, frame, level, _p(out), out.strides[0]))
        return out

    def fast_edge_corners(self, level, frame=0, cap=1 << 20):
        out = np.zeros((cap, 3), np.int32)
        n = ct.c_int32()
        self._check(self.L.orbx_get_fast_edge_corners(self._h, frame, level, _p(out), cap, ct.byref(n)))
        return out[:n.value].copy()

    def candidates(self, level, frame=0, cap=1 << 18):
        out = np.zeros((cap, 3), np.int32)
        n = ct.c_int32()
        self._check(self.L.orbx_get_candidates(self._h, frame, level, _p(out), cap, ct.byref(n)))
        return out[:n.value].copy()

    def level_counts(self, frame=0):
        out = np.zeros(self.nlevels, np.int32)
        self._check(self.L.orbx_get_level_counts(self._h, frame, _p(out)))
        return out

    def harris_responses(self, level, xy, frame=0, block=7, k=0.04):
        """cv::ORB's Harris response at integer (x, y) points of a pyramid level of the last extracted frame."""
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2)
        out = np.zeros(len(xy), np.float32)
        self._check(self.L.orbx_harris_responses(self._h, frame, level, _p(xy), len(xy), block, ct.c_float(k), _p(out)))
        return out

    def set_size(self, w, h):
        self._last_w, self._last_h = w, h

    # -- self tests --
    def test_trig(self, x):
        x = np.ascontiguousarray(x, np.float32)
        c, s = np.zeros_like(x), np.zeros_like(x)
        self._check(self.L.orbx_test_trig(self._h, _p(x), len(x), _p(c), _p(s)))
        return c, s

    def test_atan2(self, y, x):
        y = np.ascontiguousarray(y, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        o = np.zeros_like(x)
        self._check(self.L.orbx_test_atan2(self._h, _p(y), _p(x), len(x), _p(o)))
        return o

    def test_trig_checksum(self, first_bits, last_bits):
        a, b = ct.c_uint64(), ct.c_uint64()
        self._check(self.L.orbx_test_trig_checksum(self._h, first_bits, last_bits, ct.byref(a), ct.byref(b)))
        return a.value, b.value

    def test_quadtree(self, xys, box_w, box_h, wcell, hcell, ncols, N):
        xys = np.ascontiguousarray(xys, np.int32).reshape(-1, 3)
        out = np.zeros((len(xys) + 16, 3), np.int32)
        n = ct.c_int32()
        self._check(self.L.orbx_test_quadtree(self._h, _p(xys), len(xys), box_w, box_h, wcell, hcell, ncols, N,
                                              _p(out), len(out), ct.byref(n)))
        return out[:n.value].copy()

    def bench_popc(self):
        v = ct.c_double()
        self._check(self.L.orbx_bench_popc(self._h, ct.byref(v)))
        return v.value

    # -- matching (cv::BFMatcher(NORM_HAMMING)) --
    def match(self, query, train, k=1, max_dist=0.0, ratio=0.0):
        query = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        train = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        out = np.zeros(max(len(query) * k, 1), DM_DTYPE)
        n = ct.c_int32()
        self._check(self.L.orbx_match(self._h, _p(query), len(query), _p(train), len(train), k,
                                      ct.c_float(max_dist), ct.c_float(ratio), _p(out), ct.byref(n)))
        return out[:n.value].copy()


class BFMatcher:
    """cv::BFMatcher(NORM_HAMMING) call shape over an extractor handle (reference frontend.cpp:220, 294)."""

    def __init__(self, extractor):
        self.ex = extractor

    def match(self, query, train):
        return self.ex.match(query, train, k=1)

    def knnMatch(self, query, train, k=2):
        assert k == 2
        return self.ex.match(query, train, k=2).reshape(-1, 2)


class Comm:
    """The library's own NCCL communicator for the sharded landmark database (orbx_comm).  `dist` = an initialised torch.distributed
    process group used ONLY to carry rank 0's 128-byte unique id to the other ranks (any transport would do)."""

    def __init__(self, extractor, dist=None, nranks=None, rank=None, unique_id=None):
        self.ex, self.L = extractor, extractor.L
        if dist is not None:
            import torch
            nranks, rank = dist.get_world_size(), dist.get_rank()
            buf = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                raw = (ct.c_uint8 * 128)()
                st = self.L.orbx_comm_get_unique_id(raw)
                if st != OK:
                    raise OrbxError(st, (self.L.orbx_comm_last_error() or b"").decode())
                buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
            if dist.get_backend() == "nccl":
                buf = buf.cuda()
            dist.broadcast(buf, 0)
            unique_id = bytes(buf.cpu().numpy().tobytes())
        self.nranks, self.rank = int(nranks), int(rank)
        self._c = ct.c_void_p()
        raw = (ct.c_uint8 * 128).from_buffer_copy(unique_id)
        extractor._check(self.L.orbx_comm_create(extractor.handle, self.nranks, self.rank, raw, ct.byref(self._c)))

    @property
    def peer_memory(self):
        """True when the mailboxes of all ranks are mapped into each other (CUDA IPC over NVLink)"""
        return bool(self.L.orbx_comm_peer_memory(self._c))

    def set_transport(self, nccl):
        self.ex._check(self.L.orbx_comm_set_transport(self._c, 1 if nccl else 0))

    def close(self):
        if self._c.value:
            self.L.orbx_comm_destroy(self._c)
            self._c = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LandmarkDB:
    """Row-sharded landmark descriptor database (Backend::associateObservation, backend.cpp:1064-1083)."""

    def __init__(self, extractor, capacity_rows, first_index=0):
        self.ex = extractor
        self.L = extractor.L
        self._db = ct.c_void_p()
        extractor._check(self.L.orbx_db_create(extractor.handle, capacity_rows, first_index, ct.byref(self._db)))

    def close(self):
        if self._db.value:
            self.L.orbx_db_destroy(self._db)
            self._db = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def rows(self):
        return int(self.L.orbx_db_rows(self._db))

    def append(self, rows):
        rows = np.ascontiguousarray(rows, np.uint8).reshape(-1, 32)
        self.ex._check(self.L.orbx_db_append(self._db, _p(rows), len(rows)))

    def append_device(self, d_rows, nrows):
        self.ex._check(self.L.orbx_db_append_device(self._db, d_rows, nrows))

    def query_top2(self, query):
        query = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        out = np.zeros(len(query), TOP2_DTYPE)
        self.ex._check(self.L.orbx_db_query_top2(self._db, _p(query), len(query), _p(out)))
        return out

    def query_top2_device(self, d_query, nq, d_out):
        self.ex._check(self.L.orbx_db_query_top2_device(self._db, d_query, nq, d_out))

    def query_top2_sharded_device(self, comm, d_query, nq, d_out):
        """collective: per-shard top-2 -> ncclAllGather -> merge on the handle's stream (orbx_db_query_top2_sharded_device)"""
        self.ex._check(self.L.orbx_db_query_top2_sharded_device(self._db, comm._c, d_query, nq, d_out))

    def associate_sharded_device(self, comm, d_query, d_query_px, nq, pose, d_out, max_desc_dist=50.0, max_reproj_err=5.0):
        self.ex._check(self.L.orbx_db_associate_sharded_device(self._db, comm._c, d_query, d_query_px, nq, _p(pose), ct.c_float(max_desc_dist),
                                                               ct.c_double(max_reproj_err), d_out))

    def set_positions(self, xyz, first_row=0):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        self.ex._check(self.L.orbx_db_set_positions(self._db, first_row, len(xyz), _p(xyz)))

    @staticmethod
    def pose(R, t, fx, fy, cx, cy):
        p = np.zeros(1, POSE_DTYPE)
        p["R"][0] = np.asarray(R, np.float64).reshape(9)
        p["t"][0] = np.asarray(t, np.float64).reshape(3)
        p["fx"], p["fy"], p["cx"], p["cy"] = fx, fy, cx, cy
        return p

    def associate(self, query, query_px, pose, max_desc_dist=50.0, max_reproj_err=5.0):
        """Backend::associateObservation for a batch of observations: ASSOC_DTYPE per observation (landmark = global row or -1)."""
        query = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        query_px = np.ascontiguousarray(query_px, np.float32).reshape(-1, 2)
        out = np.zeros(len(query), ASSOC_DTYPE)
        self.ex._check(self.L.orbx_db_associate(self._db, _p(query), _p(query_px), len(query), _p(pose), ct.c_float(max_desc_dist),
                                                ct.c_double(max_reproj_err), _p(out)))
        return out

    def query_radius(self, query, max_dist=50.0, cap=1 << 20):
        query = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        out = np.zeros(cap, DM_DTYPE)
        n = ct.c_int32()
        self.ex._check(self.L.orbx_db_query_radius(self._db, _p(query), len(query), ct.c_float(max_dist), _p(out), cap, ct.byref(n)))
        return out[:n.value].copy()
