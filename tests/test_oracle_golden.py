"""CPU: the dependency-free C oracle (oracle/orb_oracle.c) against the committed golden vectors.

The vectors under tests/golden/ were produced by oracle/gen_golden.py with python cv2 4.13.0 driving the
same OpenCV primitives the reference calls (resize INTER_LINEAR, FAST, GaussianBlur, fastAtan2, BFMatcher)
through a literal restatement of ORB_SLAM3::ORBextractor.  This is what pins the oracle; it needs neither
cv2 nor /root/reference at run time.
"""
import glob
import os
import zlib

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def load_image(oracle, g):
    if "image" in g.files:
        return g["image"]
    return oracle.synth_gray(int(g["seed"]), int(g["frame"]), int(g["width"]), int(g["height"]))


EXTRACT_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "extract_*.npz")))


def test_golden_present():
    assert len(EXTRACT_CASES) >= 4
    for f in ("primitives.npz", "match.npz"):
        assert os.path.exists(os.path.join(GOLD, f))


@pytest.mark.parametrize("case", EXTRACT_CASES)
def test_extract_matches_golden(oracle, case):
    g = np.load(os.path.join(GOLD, case))
    img = load_image(oracle, g)
    assert crc(img) == g["image_crc"], "synthetic input generator drifted from the fixture"
    ref = oracle.COracle().extract(img, trace=True)
    for l in range(8):
        assert ref["pyramid"][l].shape == (g["level_h"][l], g["level_w"][l])
        assert crc(ref["pyramid"][l]) == g["pyr_crc"][l], "pyramid level %d" % l
        assert crc(ref["blurred"][l]) == g["blur_crc"][l], "blurred level %d" % l
        c = ref["cands"][l]
        tab = np.array(sorted(zip(c["x"].tolist(), c["y"].tolist(), c["score"].tolist())), np.int32).reshape(-1, 3)
        assert len(tab) == g["cand_counts"][l], "FAST candidate count level %d" % l
        assert crc(tab) == g["cand_crc"][l], "FAST candidates level %d" % l
    assert ref["nkeys"] == g["level_counts"].tolist()
    assert np.array_equal(ref["kps"].view(np.uint8), g["kps"].view(np.uint8)), "keypoints (order, coordinates, angle, response)"
    assert np.array_equal(ref["desc"], g["desc"]), "descriptors"


def test_primitives_match_golden(oracle):
    g = np.load(os.path.join(GOLD, "primitives.npz"))
    noise, smooth = g["noise"], g["smooth"]
    assert np.array_equal(oracle.resize_linear(noise, 109, 81), g["resize_noise_109x81"])
    assert np.array_equal(oracle.resize_linear(smooth, 133, 100), g["resize_smooth_133x100"])
    assert np.array_equal(oracle.gaussian_blur7(noise), g["blur_noise"])
    assert np.array_equal(oracle.gaussian_blur7(smooth), g["blur_smooth"])
    for th in (20, 7):
        for nm, im in (("noise", noise), ("smooth", smooth)):
            c = oracle.fast_roi(im, th)
            tab = np.stack([c["x"], c["y"], c["score"]], 1).astype(np.int32).reshape(-1, 3)
            assert np.array_equal(tab, g["fast%d_%s" % (th, nm)]), "FAST th=%d %s (raster order, scores)" % (th, nm)
    got = np.array([oracle.fast_atan2(float(y), float(x)) for y, x in zip(g["atan2_y"], g["atan2_x"])], np.float32)
    assert np.array_equal(got.view(np.uint32), g["atan2"].view(np.uint32))


def test_match_matches_golden(oracle):
    g = np.load(os.path.join(GOLD, "match.npz"))
    q, t = g["q"], g["t"]
    m = oracle.match(q, t)
    assert np.array_equal(np.stack([m["queryIdx"], m["trainIdx"], m["distance"]], 1).astype(np.float32), g["match"])
    assert (m["imgIdx"] == 0).all()
    k2 = oracle.knn2(q, t)
    got = np.stack([k2["queryIdx"], k2["trainIdx"], k2["distance"]], 2).astype(np.float32)
    assert np.array_equal(got, g["knn2"])
    # lowest-trainIdx tie-break: q[5] duplicates t[3], t[10], t[50]
    assert m[5]["trainIdx"] == 3 and m[5]["distance"] == 0
    assert k2[5, 0]["trainIdx"] == 3 and k2[5, 1]["trainIdx"] == 10
    # frame-to-frame matches of the two 320x240 golden frames
    d0 = np.load(os.path.join(GOLD, "extract_synth_320x240_s1_f0.npz"))["desc"]
    d1 = np.load(os.path.join(GOLD, "extract_synth_320x240_s1_f1.npz"))["desc"]
    mf = oracle.match(d1, d0)
    assert np.array_equal(np.stack([mf["queryIdx"], mf["trainIdx"], mf["distance"]], 1).astype(np.float32), g["frame_match"])


def test_depth_rule_is_the_integer_rule(oracle):
    """Frontend::isValidDepth (reference frontend.cpp:457-473) passes exactly u16 mm in [300, 2999]."""
    mm = np.arange(65536, dtype=np.uint16)
    d = mm.astype(np.float32) * np.float32(0.001)
    ok_float = ~((d < np.float32(0.3)) | (d > np.float32(3.0)))
    assert np.array_equal(ok_float, (mm >= 300) & (mm <= 2999))
    # the oracle's filter on a 256x256 depth map holding every u16 value, one keypoint per pixel of a few rows
    depth = mm.reshape(256, 256)
    kps = np.zeros(256 * 4, oracle.KP_DTYPE)
    ys = np.repeat(np.array([0, 1, 11, 255]), 256)
    kps["x"] = np.tile(np.arange(256), 4).astype(np.float32)
    kps["y"] = ys.astype(np.float32)
    desc = np.arange(len(kps) * 32, dtype=np.uint32).astype(np.uint8).reshape(-1, 32)
    ok, od, oi = oracle.filter_depth(kps, desc, depth)
    want = np.nonzero((depth[ys, np.tile(np.arange(256), 4)] >= 300) & (depth[ys, np.tile(np.arange(256), 4)] <= 2999))[0]
    assert np.array_equal(oi, want.astype(np.int32))
    assert np.array_equal(od, desc[want]) and np.array_equal(ok.view(np.uint8), kps[want].view(np.uint8))


def test_depth_rounding_and_bounds(oracle):
    """std::round (half away from zero) of the pixel and the 0 <= x < cols, 0 <= y < rows guard (frontend.cpp:511-517)."""
    depth = np.full((10, 10), 1000, np.uint16)
    depth[3, 5] = 0
    kps = np.zeros(5, oracle.KP_DTYPE)
    kps["x"] = [4.5, 4.49, 9.5, 2.0, -0.4]
    kps["y"] = [2.5, 3.0, 1.0, 9.6, 0.0]
    desc = np.zeros((5, 32), np.uint8)
    _, _, oi = oracle.filter_depth(kps, desc, depth)
    # (4.5,2.5)->(5,3): depth 0 -> dropped; (4.49,3.0)->(4,3) ok; (9.5,..)->x=10 out of bounds; y=9.6->10 out; (-0.4,0)->(0,0) ok
    assert oi.tolist() == [1, 4]


def test_categorize_first_containing_box(oracle):
    """Backend::categorizeObservation (reference backend.cpp:1011-1029): first box containing the pixel, inclusive bounds."""
    boxes = np.zeros(3, oracle.BOX_DTYPE)
    boxes["cx"], boxes["cy"], boxes["w"], boxes["h"], boxes["class_id"] = [50, 60, 200], [50, 50, 200], [20, 60, 10], [20, 60, 10], [1, 2, 3]
    assert oracle.categorize(50.0, 50.0, boxes) == 1          # inside box 0 and box 1: first wins
    assert oracle.categorize(40.0, 40.0, boxes) == 1          # on the corner of box 0: inclusive
    assert oracle.categorize(39.9, 40.0, boxes) == 2          # just outside box 0, inside box 1
    assert oracle.categorize(300.0, 300.0, boxes) == -1       # "unlabeled"
    assert oracle.categorize(205.0, 195.0, boxes) == 3


def test_introsort_emulation_matches_libstdcxx(oracle):
    """The quadtree's std::sort on (count, UL.x) leaves ties in libstdc++ introsort order (ORBextractor.cpp:700);
    the C restatement of that sort is checked against the real std::sort (oracle/stdsort_shim.cpp)."""
    import ctypes as ct
    shim = ct.CDLL(os.path.join(os.path.dirname(GOLD), "..", "oracle", "libstdsort_shim.so"))
    rng = np.random.default_rng(7)
    for n in (1, 2, 15, 16, 17, 33, 100, 257, 1000):
        for _ in range(5):
            cnt = rng.integers(1, 6, n).astype(np.int32)
            ulx = (rng.integers(0, 8, n) * 16).astype(np.int32)
            want = (ct.c_int * n)(*range(n))
            shim.real_std_sort((ct.c_int * n)(*cnt.tolist()), (ct.c_int * n)(*ulx.tolist()), want, n)
            assert oracle.introsort_pairs(cnt, ulx) == list(want), n


def _cull_literal(response, match_query, max_new, min_response):
    """Reference frontend.cpp:1168-1218 stated literally: std::set of matched query indices, matched keypoints in match order, unmatched
    (response, index) pairs in index order sorted by the REAL std::sort with `a.first > b.first`, taken while added < max_new and
    response >= min_response."""
    import ctypes as ct
    shim = ct.CDLL(os.path.join(os.path.dirname(GOLD), "..", "oracle", "libstdsort_shim.so"))
    matched = set(int(q) for q in match_query)
    out = [int(q) for q in match_query]
    un = [(np.float32(response[i]), i) for i in range(len(response)) if i not in matched]
    n = len(un)
    r = (ct.c_float * max(n, 1))(*[float(a) for a, _ in un])
    ix = (ct.c_int * max(n, 1))(*[b for _, b in un])
    shim.real_std_sort_response_desc(r, ix, n)
    added = 0
    for k in range(n):
        if added >= max_new or r[k] < min_response:
            break
        out.append(int(ix[k])); added += 1
    return out


def test_cull_rule_matches_libstdcxx(oracle):
    """Feature culling for the backend (reference frontend.cpp:1168-1218): the C oracle against the literal statement above on
    tie-heavy responses (FAST scores are small integers, so the std::sort tie order decides WHICH features survive the 200 cut)."""
    rng = np.random.default_rng(23)
    for n, nm, lo, hi in [(0, 0, 7, 60), (1, 0, 7, 60), (1, 1, 7, 60), (17, 3, 40, 70), (300, 100, 45, 56), (1000, 593, 7, 120),
                          (1000, 0, 50, 52), (1000, 1000, 7, 120), (800, 10, 7, 49), (4096, 77, 48, 53)]:
        resp = rng.integers(lo, hi, n).astype(np.float32)
        q = rng.permutation(n)[:nm].astype(np.int32)
        for max_new, min_resp in ((200, 50.0), (5, 0.0), (0, 50.0), (10000, 51.5)):
            got = oracle.cull_keyframe(resp, q, max_new, min_resp).tolist()
            assert got == _cull_literal(resp, q, max_new, min_resp), (n, nm, max_new, min_resp)
    with pytest.raises(ValueError):
        oracle.cull_keyframe(np.ones(4, np.float32), np.array([4], np.int32))


def test_extract_edge_cases(oracle):
    orc = oracle.COracle()
    # featureless frame: no keypoints, not an error
    flat = np.full((240, 320), 77, np.uint8)
    r = orc.extract(flat)
    assert len(r["kps"]) == 0 and r["desc"].shape == (0, 32)
    # empty image: the reference's operator() returns -1 (ORBextractor.cpp:1090-1091)
    import ctypes as ct
    rc = oracle.lib().orc_extract(ct.byref(orc.ex), None, 0, 0, ct.c_size_t(0), None, None, 0, None)
    assert rc == -1
    # keypoints stay >= 19 px from every border at level scale
    g = oracle.synth_gray(2, 0, 320, 240)
    r = orc.extract(g)
    k = r["kps"]
    s = orc.scale[k["octave"]]
    x, y = k["x"] / s, k["y"] / s
    for l in range(8):
        lw, lh = orc.level_size(320, 240, l)
        sel = k["octave"] == l
        if sel.any():
            assert x[sel].min() >= 19 - 1e-3 and y[sel].min() >= 19 - 1e-3
            assert x[sel].max() < lw - 19 + 1e-3 and y[sel].max() < lh - 19 + 1e-3
    assert (k["class_id"] == -1).all() and (k["response"] > 0).all()
