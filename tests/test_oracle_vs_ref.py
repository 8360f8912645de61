"""CPU: the C oracle (oracle/orb_oracle.c) against the REFERENCE ITSELF.

oracle/_ref/libref_orbextractor.so is /root/reference/dynamic_visual_slam/src/ORBextractor.cpp compiled unmodified (recipe:
oracle/Makefile target `_ref`; header shim oracle/ref_shim/, OpenCV primitives forwarded to the C primitives that
tests/test_oracle_vs_cv2.py pins against real cv2).  Its std::list / std::sort / DivideNode / per-cell FAST loop / operator()
assembly are the reference's own machine code, so equality here pins the restatement — and, through it, the golden vectors and
the CUDA path — to the reference rather than to the builder's reading of it.
"""
import glob
import os
import zlib

import numpy as np
import pytest

import ref_oracle as ro

pytestmark = pytest.mark.skipif(not ro.available(), reason="oracle/_ref not built and /root/reference absent")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def circles_image():
    """the reference's own fixture (test/test_dbow2_integration.cpp:14-17), as stored in the golden file"""
    return np.load(os.path.join(GOLD, "extract_circles_640x480.npz"))["image"]


@pytest.fixture(scope="module")
def ref(built):
    return ro.RefExtractor()


def _same(a, b):
    return len(a["kps"]) == len(b["kps"]) and np.array_equal(a["kps"].view(np.uint8), b["kps"].view(np.uint8)) and np.array_equal(a["desc"], b["desc"])


def test_ref_is_the_reference_source(ref):
    assert ro.lib().ref_source_path().decode().endswith("dynamic_visual_slam/src/ORBextractor.cpp")
    sums = os.path.join(os.path.dirname(ro.SO), "SOURCES.sha256")
    assert os.path.exists(sums) and "ORBextractor.cpp" in open(sums).read()


def test_ctor_tables(ref, oracle):
    """ORBextractor.cpp:409-469 through the reference's getters"""
    for args in [(1000, 1.2, 8, 20, 7), (2000, 1.2, 8, 20, 7), (500, 1.5, 4, 20, 7), (1500, 1.1, 12, 15, 5), (300, 2.0, 3, 20, 7)]:
        r, o = ro.RefExtractor(*args), oracle.COracle(*args)
        t = r.tables()
        n = args[2]
        assert t["nfeat"].tolist() == o.nfeat and t["umax"].tolist() == o.umax
        for name in ("scale", "inv_scale", "sigma2", "inv_sigma2"):
            assert np.array_equal(t[name].view(np.uint32), np.array(getattr(o.ex, name)[:n], np.float32).view(np.uint32)), name


@pytest.mark.parametrize("w,h,seed,frame", [(1280, 720, 3, 0), (1280, 720, 1234, 5), (640, 480, 5, 0), (640, 480, 5, 1), (741, 417, 7, 0),
                                            (320, 240, 1, 0), (417, 301, 4, 0), (1281, 721, 9, 0), (257, 193, 2, 0), (160, 120, 6, 0), (100, 100, 6, 0)])
def test_operator_call_equals_reference(ref, oracle, w, h, seed, frame):
    g = oracle.synth_gray(seed, frame, w, h)
    r = ref.extract(g)
    o = oracle.COracle().extract(g)
    assert r["ret"] == len(r["kps"])
    assert _same(r, o), "keypoints (order, coordinates, angle, response, octave, size) or descriptors differ from the reference"


def test_three_circle_fixture_equals_reference(ref, oracle):
    g = circles_image()
    assert _same(ref.extract(g), oracle.COracle().extract(g))


def test_other_parameters_equal_reference(oracle):
    g = oracle.synth_gray(21, 0, 800, 600)
    for args in [(2000, 1.2, 8, 20, 7), (500, 1.5, 4, 20, 7), (1500, 1.1, 10, 15, 5), (250, 1.2, 8, 40, 12)]:
        assert _same(ro.RefExtractor(*args).extract(g), oracle.COracle(*args).extract(g)), args


@pytest.mark.parametrize("w,h,seed", [(1280, 720, 3), (640, 480, 5), (741, 417, 7)])
def test_stages_equal_reference(ref, oracle, w, h, seed):
    """pyramid (public mvImagePyramid), the candidate list handed to DistributeOctTree (every cv::FAST call of the per-cell loop
    observed, :781-872), the retained keypoints per level in list order with IC_Angle — all against the restatement's trace."""
    g = oracle.synth_gray(seed, 0, w, h)
    tr = ref.stage_trace(g)
    o = oracle.COracle().extract(g, trace=True)
    for l in range(8):
        assert np.array_equal(tr["pyramid"][l], o["pyramid"][l]), "pyramid level %d" % l
        rc, oc = tr["cands"][l], o["cands"][l]
        assert len(rc) == len(oc) and np.array_equal(rc.view(np.int32), oc.view(np.int32)), "FAST candidates level %d (push order)" % l
        assert len(tr["keys"][l]) == o["nkeys"][l], "retained count level %d" % l
    # retained keypoints of level l before `pt *= scale`: octave, size, angle, response must equal the final output's
    off = 0
    for l in range(8):
        k = tr["keys"][l]
        f = o["kps"][off:off + len(k)]
        off += len(k)
        for name in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(k[name], f[name]), (l, name)
        s = np.float32(oracle.COracle().scale[l])
        assert np.array_equal((k["x"] * s if l else k["x"]).astype(np.float32), f["x"]) and np.array_equal((k["y"] * s if l else k["y"]).astype(np.float32), f["y"])
    assert (tr["calls_ini"] >= tr["calls_min"]).all() and tr["calls_ini"].sum() > 0


def test_pyramid_border_ring_is_reflect101(ref, oracle):
    """the 19-px ring (:1184-1190) is never read on this path and is not materialised by the build; pin what it is"""
    g = oracle.synth_gray(8, 0, 320, 240)
    ref.extract(g)
    for l in (0, 3):
        inner, pad = ref.level(l), ref.level(l, padded=True)
        assert np.array_equal(pad, np.pad(inner, 19, mode="reflect"))


def test_min_threshold_retry_cells(ref, oracle):
    """a low-contrast image forces the th=20 → th=7 retry in most cells"""
    g = (oracle.synth_gray(13, 0, 640, 480).astype(np.int32) // 6 + 100).astype(np.uint8)
    tr = ref.stage_trace(g)
    assert tr["calls_min"].sum() > 50
    assert _same(ref.extract(g), oracle.COracle().extract(g))


def test_empty_and_featureless(ref, oracle):
    assert ref.extract(np.zeros((0, 0), np.uint8))["ret"] == -1                 # :1090-1091
    flat = np.full((240, 320), 128, np.uint8)
    r = ref.extract(flat)
    assert r["ret"] == 0 and len(r["kps"]) == 0
    assert len(oracle.COracle().extract(flat)["kps"]) == 0


def test_defined_domain_equals_reference(ref, oracle):
    """Small / elongated frames: the reference throws (a 32-row level or a negative root count makes vpIniNodes.resize throw, :559-566;
    a zero-pixel level makes cv::resize throw) or indexes an empty vector (nIni == 0, :586).  orc_geometry_status names those sizes;
    everywhere else the outputs are equal.  Status-3 sizes are not executed (the reference dereferences a null pointer there)."""
    orc = oracle.COracle()
    seen = {0: 0, 1: 0, 2: 0, 3: 0}
    for w in (1, 3, 20, 33, 64, 67, 96, 100, 115, 130, 160, 240, 300):
        for h in (1, 20, 32, 40, 64, 80, 100, 115, 200, 240):
            st = orc.geometry_status(w, h)
            seen[st] += 1
            if st == 3:
                continue
            g = oracle.synth_gray(6, 0, w, h)
            r = ref.extract(g)
            if st == 0:
                assert _same(r, orc.extract(g)), (w, h)
            else:
                assert r["ret"] == -3, (w, h, st)
                with pytest.raises(oracle.GeometryError):
                    orc.extract(g)
    assert all(v > 0 for v in seen.values()), seen
    assert orc.geometry_status(115, 300) == 3 and orc.geometry_status(96, 80) == 2 and orc.geometry_status(1280, 720) == 0


def test_quadtree_tie_heavy_equals_reference(ref, oracle):
    """DistributeOctTree (:555-779) alone, on candidate sets built so that many nodes tie in the comparator (equal count, equal UL.x):
    the order libstdc++'s introsort leaves them in decides which nodes are split last, hence the output."""
    rng = np.random.default_rng(77)
    cases = [(3000, 1248, 688, 217), (1800, 1035, 568, 181), (900, 325, 209, 60), (5000, 1248, 688, 500), (400, 600, 300, 151), (2500, 857, 468, 1000),
             (64, 300, 300, 60), (5, 100, 100, 60), (1, 100, 100, 60), (0, 100, 100, 60)]
    for n, W, H, N in cases:
        for rep in range(4):
            gx, gy = rng.integers(0, W // 8, n) * 8 + rng.integers(0, 2, n), rng.integers(0, H // 8, n) * 8 + rng.integers(0, 2, n)
            pts = sorted(set(zip(np.minimum(gx, W - 1).tolist(), np.minimum(gy, H - 1).tolist())))
            sc = rng.integers(7, 60, len(pts))
            order = rng.permutation(len(pts))
            c = np.zeros(len(pts), oracle.CAND_DTYPE)
            c["x"], c["y"], c["score"] = [pts[i][0] for i in order], [pts[i][1] for i in order], [int(sc[i]) for i in order]
            want = ref.distribute_octtree(c, 16, 16 + W, 16, 16 + H, N)
            got = oracle.distribute_octtree(c, 16, 16 + W, 16, 16 + H, N)
            assert got.tolist() == want.tolist(), (n, W, H, N, rep)


def test_lapping_area_split(ref, oracle):
    """operator()'s mono/stereo split (:1152-1161): keypoints with lap0 <= x <= lap1 fill the arrays from the back, the rest from the
    front; the return value is the mono count.  Pins the rule the C++ adapter (host/ORBextractor.hpp) reproduces."""
    g = oracle.synth_gray(3, 0, 640, 480)
    base = ref.extract(g)
    for lap in [(100, 300), (0, 639), (640, 700), (250, 250)]:
        r = ref.extract(g, lapping=lap)
        k = base["kps"]
        inside = (k["x"] >= lap[0]) & (k["x"] <= lap[1])
        mono, stereo = np.nonzero(~inside)[0], np.nonzero(inside)[0]
        order = np.concatenate([mono, stereo[::-1]]).astype(np.int64)
        assert r["ret"] == len(mono)
        assert np.array_equal(r["kps"].view(np.uint8), k[order].view(np.uint8)) and np.array_equal(r["desc"], base["desc"][order])


def test_goldens_are_reference_output(ref, oracle):
    """every committed extraction vector equals what the compiled reference produces today"""
    cases = sorted(glob.glob(os.path.join(GOLD, "extract_*.npz")))
    assert len(cases) >= 4
    for path in cases:
        gd = np.load(path)
        img = gd["image"] if "image" in gd.files else oracle.synth_gray(int(gd["seed"]), int(gd["frame"]), int(gd["width"]), int(gd["height"]))
        r = ref.extract(img)
        assert np.array_equal(r["kps"].view(np.uint8), gd["kps"].view(np.uint8)) and np.array_equal(r["desc"], gd["desc"]), path
    sums = np.load(os.path.join(GOLD, "ref_stream_1280x720.npz"))
    for i, f in enumerate(sums["frames"].tolist()):
        r = ref.extract(oracle.synth_gray(int(sums["seed"]), f, 1280, 720))
        assert (len(r["kps"]), crc(r["kps"]), crc(r["desc"])) == (int(sums["count"][i]), int(sums["kps_crc"][i]), int(sums["desc_crc"][i])), f


def test_dead_per_cell_topn_variant_runs(ref, oracle):
    """ComputeKeyPointsOld (:898-1075, retainBest :1049/:1067) is dead in the reference (call commented out at :1101); it compiles
    and runs here, each level within its quota"""
    g = oracle.synth_gray(3, 0, 640, 480)
    lv = ref.keypoints_old(g)
    nf = ref.tables()["nfeat"]
    assert sum(len(k) for k in lv) > 300
    for l, k in enumerate(lv):
        assert len(k) <= nf[l] and (k["octave"] == l).all()


def test_batch_entry_equals_single(oracle):
    frames = np.stack([oracle.synth_gray(1234, f, 320, 240) for f in range(6)])
    kps, desc, cnt = ro.extract_batch(frames, cap=2048, nthreads=3)
    okps, odesc, ocnt = oracle.COracle().extract_batch(frames, cap=2048)
    assert cnt.tolist() == ocnt.tolist()
    for f in range(6):
        assert np.array_equal(kps[f, :cnt[f]].view(np.uint8), okps[f, :cnt[f]].view(np.uint8)) and np.array_equal(desc[f, :cnt[f]], odesc[f, :cnt[f]])
