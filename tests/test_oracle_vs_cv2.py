"""CPU: live cross-check of the C oracle against python cv2 (when cv2 is importable).

Layer L-A (oracle/cv_oracle.py) is a literal restatement of ORB_SLAM3::ORBextractor that calls the SAME OpenCV
primitives the reference calls; layer L-B (oracle/orb_oracle.c) restates those primitives in plain C.  Here
L-B must reproduce L-A bit for bit on seeded inputs, including odd sizes (hypothesis).  The committed golden
vectors (test_oracle_golden.py) cover the case where cv2 is absent.
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from hypothesis import given, settings, strategies as st   # noqa: E402


def _table(k):
    return np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"], k["octave"], k["class_id"]], 1).astype(np.float64)


@pytest.mark.parametrize("w,h,seed", [(640, 480, 3), (741, 417, 11), (333, 257, 5)])
def test_extractor_bit_exact_vs_cv2(oracle, w, h, seed):
    import cv_oracle as cvo
    g = oracle.synth_gray(seed, 0, w, h)
    tr = {}
    tab, desc = cvo.ORBextractorCV()(g, trace=tr)
    ref = oracle.COracle().extract(g, trace=True)
    for l in range(8):
        assert np.array_equal(ref["pyramid"][l], tr["pyramid"][l]), "pyramid level %d" % l
        assert np.array_equal(ref["blurred"][l], tr["blurred"][l]), "blurred level %d" % l
        c = ref["cands"][l]
        assert list(zip(c["x"].tolist(), c["y"].tolist(), c["score"].tolist())) == tr["cands"][l], "FAST cell candidates, reference order"
    assert np.array_equal(_table(ref["kps"]), tab)
    assert np.array_equal(ref["desc"], desc)


def test_1280x720_frame_vs_cv2(oracle):
    import cv_oracle as cvo
    g = oracle.synth_gray(20261018, 0, 1280, 720)
    tab, desc = cvo.ORBextractorCV()(g)
    ref = oracle.COracle().extract(g)
    assert len(tab) >= 1000
    assert np.array_equal(_table(ref["kps"]), tab) and np.array_equal(ref["desc"], desc)


@settings(max_examples=25, deadline=None)
@given(sw=st.integers(40, 400), sh=st.integers(40, 300), seed=st.integers(0, 1 << 30))
def test_resize_blur_fast_any_size(oracle, sw, sh, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    if seed & 1:
        img = cv2.GaussianBlur(img, (5, 5), 1.5)
    dw, dh = int(np.rint(np.float32(sw) * np.float32(1 / 1.2))), int(np.rint(np.float32(sh) * np.float32(1 / 1.2)))
    assert np.array_equal(oracle.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
    assert np.array_equal(oracle.gaussian_blur7(img), cv2.GaussianBlur(img.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101))
    for th in (20, 7):
        kp = cv2.FastFeatureDetector_create(th, True).detect(img)
        c = oracle.fast_roi(img, th)
        assert [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kp] == list(zip(c["x"].tolist(), c["y"].tolist(), c["score"].tolist()))


def test_fast_atan2_vs_cv2(oracle):
    rng = np.random.default_rng(3)
    ys = rng.integers(-2900000, 2900000, 20000).astype(np.float32)
    xs = rng.integers(-2900000, 2900000, 20000).astype(np.float32)
    for y, x in zip(ys[:5000], xs[:5000]):
        assert np.float32(oracle.fast_atan2(float(y), float(x))) == np.float32(cv2.fastAtan2(float(y), float(x)))


def test_matcher_vs_cv2(oracle):
    import cv_oracle as cvo
    rng = np.random.default_rng(11)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    t = rng.integers(0, 4, (500, 32), dtype=np.uint8)          # few distinct bit patterns -> many distance ties
    q[:100] = t[rng.integers(0, 500, 100)]
    m = oracle.match(q, t)
    assert [(int(a), int(b), float(c)) for a, b, c in zip(m["queryIdx"], m["trainIdx"], m["distance"])] == cvo.bf_match(q, t)
    k2 = oracle.knn2(q, t)
    want = cvo.bf_knn2(q, t)
    got = [[(int(r["queryIdx"]), int(r["trainIdx"]), float(r["distance"])) for r in row] for row in k2]
    assert got == want


def test_quadtree_vs_list_restatement(oracle):
    """DistributeOctTree: the C oracle against the Python std::list restatement driven by the real std::sort."""
    import cv_oracle as cvo
    rng = np.random.default_rng(5)
    for n, W, H, N in [(50, 600, 300, 30), (3000, 1248, 688, 217), (800, 325, 209, 60), (5, 100, 100, 10), (1500, 400, 400, 1500)]:
        xs, ys = rng.integers(0, W, n), rng.integers(0, H, n)
        sc = rng.integers(7, 120, n)
        keys = sorted(set(zip(xs.tolist(), ys.tolist())))
        keys = [(x, y, int(s)) for (x, y), s in zip(keys, sc)]
        want = cvo.distribute_octtree(keys, 16, 16 + W, 16, 16 + H, N)
        c = np.zeros(len(keys), oracle.CAND_DTYPE)
        c["x"], c["y"], c["score"] = [k[0] for k in keys], [k[1] for k in keys], [k[2] for k in keys]
        got = oracle.distribute_octtree(c, 16, 16 + W, 16, 16 + H, N)
        assert list(zip(got["x"].tolist(), got["y"].tolist(), got["score"].tolist())) == [tuple(k) for k in want], (n, W, H, N)


def test_reproject_and_associate_vs_cv2(oracle):
    """Backend::reprojectPoint / associateObservation (reference backend.cpp:1064-1173): the C oracle against a literal cv2/numpy statement
    (cv2.gemm for R.t()*(X - t), float32 pixel, double norm)."""
    rng = np.random.default_rng(17)
    R, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    t = rng.standard_normal(3)
    fx, fy, cx, cy = 615.3, 615.9, 640.2, 360.4
    n, nq = 400, 60
    rows = oracle.synth_descriptors(7, 0, n)
    pc = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1, 1, n), rng.uniform(-1.0, 5.0, n)], 1)
    pos = (pc @ R.T + t).astype(np.float32)

    def reproj(p):
        pcam = cv2.gemm(R, p.astype(np.float64).reshape(3, 1) - t.reshape(3, 1), 1.0, None, 0.0, flags=cv2.GEMM_1_T)
        x, y, z = pcam[0, 0], pcam[1, 0], pcam[2, 0]
        if z <= 0:
            return np.array([-1, -1], np.float32)
        return np.array([np.float32(fx * x / z + cx), np.float32(fy * y / z + cy)], np.float32)

    for j in range(n):
        assert np.array_equal(oracle.reproject(pos[j], R, t, fx, fy, cx, cy).view(np.uint32), reproj(pos[j]).view(np.uint32))
    src = rng.integers(0, n, nq)
    q = rows[src].copy()
    q[::3, 5] ^= 0xFF
    qpx = (np.stack([reproj(pos[j]) for j in src]) + rng.normal(0, 3, (nq, 2))).astype(np.float32)
    idx, err, dist = oracle.associate(q, qpx, rows, pos, R, t, fx, fy, cx, cy)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    for i in range(nq):
        best, best_err = -1, np.finfo(np.float64).max
        for j in range(n):
            m = bf.match(q[i:i + 1], rows[j:j + 1])                       # the reference's 1x1 match per landmark (backend.cpp:1072)
            if m and m[0].distance < 50.0:
                d = qpx[i] - reproj(pos[j])                               # Point2f difference in float
                e = np.sqrt(np.float64(d[0]) * np.float64(d[0]) + np.float64(d[1]) * np.float64(d[1]))
                if e < 5.0 and e < best_err:
                    best, best_err = j, e
        assert idx[i] == best and (best < 0 or err[i] == best_err)


def test_bgr2gray_vs_cv2(oracle):
    """cv::cvtColor(BGR2GRAY) (reference frontend.cpp:1084): identity on equal channels, integer formula otherwise."""
    rng = np.random.default_rng(2)
    bgr = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert np.array_equal(oracle.bgr2gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    g = rng.integers(0, 256, (50, 64), dtype=np.uint8)
    assert np.array_equal(oracle.bgr2gray(np.stack([g, g, g], 2)), g)


def test_pack_keyframe_vs_cv2(oracle):
    """Frontend::publishKeyframe packing loop (reference frontend.cpp:731-776) against a literal numpy/cv2 statement."""
    w, h = 320, 240
    g = oracle.synth_gray(6, 0, w, h)
    depth = oracle.synth_depth(6, 0, w, h)
    r = oracle.COracle().extract(g)
    rng = np.random.default_rng(8)
    R, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    t = rng.standard_normal(3)
    fx, fy, cx, cy = np.float32(305.4), np.float32(306.1), np.float32(160.3), np.float32(119.7)
    rec = oracle.pack_keyframe(r["kps"], r["desc"], depth, fx, fy, cx, cy, R, t)
    want = []
    for i, kp in enumerate(r["kps"]):
        x, y = int(np.floor(abs(kp["x"]) + 0.5) * np.sign(kp["x"])), int(np.floor(abs(kp["y"]) + 0.5) * np.sign(kp["y"]))
        d = np.float32(depth[y, x]) * np.float32(0.001)
        X = np.float32(np.float32(kp["x"] - cx) * d) / fx
        Y = np.float32(np.float32(kp["y"] - cy) * d) / fy
        if float(d) > 0.3 and float(d) < 3.0:
            pw = cv2.gemm(R, np.array([[X], [Y], [d]], np.float64), 1.0, t.reshape(3, 1), 1.0).ravel()
            want.append((i, pw, float(kp["x"]), float(kp["y"]), r["desc"][i]))
    assert len(rec) == len(want) and 0 < len(rec) < len(r["kps"])
    for a, (i, pw, px, py, dsc) in zip(rec, want):
        assert a["landmark_id"] == i and np.array_equal(a["position"].view(np.uint64), pw.view(np.uint64))
        assert a["pixel_x"] == px and a["pixel_y"] == py and np.array_equal(a["descriptor"], dsc)


def test_harris_response_vs_cv_orb(oracle):
    """HARRIS_SCORE: the oracle's Harris response equals cv::ORB's keypoint.response on a single-level cv2.ORB (north_star: <= 1e-4 rel)."""
    g = oracle.synth_gray(3, 0, 640, 480)
    kps = cv2.ORB_create(nfeatures=500, nlevels=1).detect(g)
    assert len(kps) > 100
    for kp in kps:
        got = oracle.harris_response(g, int(round(kp.pt[0])), int(round(kp.pt[1])))
        assert abs(got - kp.response) <= 1e-4 * abs(kp.response)
