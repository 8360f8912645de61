"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, stage by stage.

Bars (BASELINE.json north_star): pyramid levels, FAST corners, smoothed images, descriptors and Hamming
matches bit-exact; retained keypoint sets identical (we also require identical ORDER, i.e. the
reference's list order, which DMatch indices depend on).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [(1280, 720, 7), (640, 480, 3), (741, 417, 11), (200, 150, 9)]


@pytest.fixture(scope="module")
def ex(built):
    import orbx
    e = orbx.ORBextractor(max_width=1280, max_height=720, max_batch=4)
    yield e
    e.close()


def _kp_table(k):
    return np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"], k["octave"].astype(np.float32),
                     k["class_id"].astype(np.float32)], 1)


def test_reference_binary_present(refx):
    """the compiled reference (oracle/_ref/libref_orbextractor.so) must have travelled to the GPU box: the parity tests below compare
    the CUDA path with it directly, not only with the C restatement"""
    assert refx is not None


@pytest.mark.parametrize("w,h,seed", SIZES)
def test_stages_bit_exact(ex, oracle, refx, w, h, seed):
    g = oracle.synth_gray(seed, 0, w, h)
    orc = oracle.COracle()
    ref = orc.extract(g, trace=True)
    kps, desc = ex(g)
    # a1: ctor tables
    assert list(ex.features_per_level()) == orc.nfeat
    assert np.array_equal(ex.GetScaleFactors(), orc.scale)
    for l in range(8):
        # a2: pyramid levels bit-exact
        assert np.array_equal(ex.pyramid_level(l), ref["pyramid"][l]), "pyramid level %d" % l
        # a7: smoothed levels bit-exact
        assert np.array_equal(ex.blurred_level(l), ref["blurred"][l]), "blurred level %d" % l
        # a3: FAST corners (set equality; the device list is unordered by design)
        c = ex.candidates(l)
        r = ref["cands"][l]
        rc = np.stack([r["x"], r["y"], r["score"]], 1)
        assert len(c) == len(rc), ("candidate count level %d" % l, len(c), len(rc))
        cs = c[np.lexsort((c[:, 0], c[:, 1]))]
        rs = rc[np.lexsort((rc[:, 0], rc[:, 1]))]
        assert np.array_equal(cs, rs), "FAST candidates level %d" % l
    # a4: retained per level
    assert list(ex.level_counts()) == ref["nkeys"]
    # a5/a6/a8/a9: keypoints (order, coordinates, size, angle, response, octave) and descriptors bit-exact
    assert len(kps) == len(ref["kps"])
    assert np.array_equal(kps.view(np.uint8), ref["kps"].view(np.uint8)), "keypoints"
    assert np.array_equal(desc, ref["desc"]), "descriptors"
    # the same against the reference's own compiled ORBextractor.cpp: operator() output and its public mvImagePyramid
    r = refx.extract(g)
    assert r["ret"] == len(kps) and np.array_equal(kps.view(np.uint8), r["kps"].view(np.uint8)) and np.array_equal(desc, r["desc"]), "vs compiled reference"
    for l in range(8):
        assert np.array_equal(ex.pyramid_level(l), refx.level(l)), "mvImagePyramid[%d] of the compiled reference" % l


def test_separate_blur_path_and_border_keypoints(ex, oracle):
    """ORBX_OPT_FUSED_BLUR = 0 (blur every level with k_blur7, descriptors read the blurred pyramid) gives the same bits as the default
    fused descriptor kernel; the frame is built so that corners sit right at the 19-px limit, where the fused kernel's 43 x 43 window
    leaves the level and has to mirror rows and columns (BORDER_REFLECT_101)."""
    w, h = 640, 480
    rng = np.random.default_rng(12)
    g = oracle.synth_gray(8, 0, w, h).copy()
    band = rng.integers(0, 256, (h, w), dtype=np.uint8)
    g[:40], g[-40:], g[:, :40], g[:, -40:] = band[:40], band[-40:], band[:, :40], band[:, -40:]       # noise frame: corners everywhere near the edges
    ref = oracle.COracle().extract(g, trace=True)
    k = ref["kps"]
    sc = oracle.COracle().scale
    lx, ly = np.rint(k["x"] / sc[k["octave"]]), np.rint(k["y"] / sc[k["octave"]])                       # level coordinates
    lw = np.array([ref["pyramid"][o].shape[1] for o in k["octave"]]); lh = np.array([ref["pyramid"][o].shape[0] for o in k["octave"]])
    assert (lx <= 20).any() and (ly <= 20).any() and (lx >= lw - 21).any() and (ly >= lh - 21).any()       # windows leaving the level on all four sides
    for fused in (True, False, True):
        ex.set_fused_blur(fused)
        kps, desc = ex(g)
        assert np.array_equal(kps.view(np.uint8), k.view(np.uint8)) and np.array_equal(desc, ref["desc"]), "fused=%s" % fused
        for l in (0, 3, 7):
            assert np.array_equal(ex.blurred_level(l), ref["blurred"][l])


def test_full_hd_and_other_parameters(built, oracle):
    """1920x1080 and non-default extractor parameters (nfeatures, scale factor, levels, thresholds) stay bit-exact."""
    import orbx
    for (w, h, kw) in [(1920, 1080, dict()), (960, 540, dict(nfeatures=2000, scaleFactor=1.3, nlevels=6, iniThFAST=30, minThFAST=10)),
                       (800, 600, dict(nfeatures=500, scaleFactor=1.1, nlevels=10, iniThFAST=12, minThFAST=5)),
                       # scale 1.9: the widest byte span the word-based horizontal pass of the resize takes; 2.5: its byte-wise path
                       (1280, 720, dict(nfeatures=800, scaleFactor=1.9, nlevels=4)), (1280, 720, dict(nfeatures=800, scaleFactor=2.5, nlevels=3))]:
        e = orbx.ORBextractor(max_width=w, max_height=h, **kw)
        try:
            g = oracle.synth_gray(31, 1, w, h)
            ref = oracle.COracle(**kw).extract(g, trace=True)
            kps, desc = e(g, cap=8192)
            nl = kw.get("nlevels", 8)
            for l in range(nl):
                assert np.array_equal(e.pyramid_level(l), ref["pyramid"][l]), ("pyramid", w, h, l)
                assert np.array_equal(e.blurred_level(l), ref["blurred"][l]), ("blurred", w, h, l)
            assert list(e.level_counts()) == ref["nkeys"]
            assert np.array_equal(kps.view(np.uint8), ref["kps"].view(np.uint8)) and np.array_equal(desc, ref["desc"]), (w, h, kw)
        finally:
            e.close()


@pytest.mark.parametrize("w,h", [(257, 193), (511, 383), (513, 385), (1027, 771), (389, 263), (1281, 721), (833, 479), (1279, 719), (100, 100), (160, 120)])
def test_awkward_sizes(built, oracle, refx, w, h):
    """Sizes around the tile boundaries of the kernels (256-px blur tiles, 192-px resize tiles, 16-byte TMA columns)."""
    import orbx
    e = orbx.ORBextractor(max_width=w, max_height=h)
    try:
        g = oracle.synth_gray(w * 7 + h, 0, w, h)
        ref = oracle.COracle().extract(g, trace=True)
        kps, desc = e(g, cap=8192)
        for l in range(8):
            assert np.array_equal(e.pyramid_level(l), ref["pyramid"][l]), ("pyramid", l)
            assert np.array_equal(e.blurred_level(l), ref["blurred"][l]), ("blurred", l)
            c = e.candidates(l)
            r = ref["cands"][l]
            assert sorted(map(tuple, c.tolist())) == sorted(zip(r["x"].tolist(), r["y"].tolist(), r["score"].tolist())), ("FAST", l)
        assert np.array_equal(kps.view(np.uint8), ref["kps"].view(np.uint8)) and np.array_equal(desc, ref["desc"])
        r = refx.extract(g)
        assert np.array_equal(kps.view(np.uint8), r["kps"].view(np.uint8)) and np.array_equal(desc, r["desc"]), "vs compiled reference"
    finally:
        e.close()


def test_frame_sizes_where_the_reference_throws_are_refused(built, oracle):
    """Small or very elongated frames make the reference throw (std::length_error from vpIniNodes.resize, ORBextractor.cpp:559-566;
    cv::resize on a vanished level) or index an empty vector (:586) — pinned with the compiled reference in
    tests/test_oracle_vs_ref.py::test_defined_domain_equals_reference.  The library refuses exactly those sizes."""
    import orbx
    orc = oracle.COracle()
    e = orbx.ORBextractor(max_width=320, max_height=320)
    try:
        seen = set()
        for w in (1, 3, 20, 33, 64, 67, 96, 100, 115, 130, 160, 240, 300):
            for h in (1, 20, 32, 40, 64, 80, 100, 115, 200, 240):
                st = orc.geometry_status(w, h)
                seen.add(st)
                g = oracle.synth_gray(6, 0, w, h)
                if st == 0:
                    kps, desc = e(g)
                    want = orc.extract(g)
                    assert np.array_equal(kps.view(np.uint8), want["kps"].view(np.uint8)) and np.array_equal(desc, want["desc"]), (w, h)
                else:
                    with pytest.raises(orbx.OrbxError) as err:
                        e(g)
                    assert err.value.status == orbx.E_UNSUPPORTED, (w, h, st)
        assert seen == {0, 1, 2, 3}
    finally:
        e.close()


def test_match_bit_exact(ex, oracle):
    w, h = 1280, 720
    g0, g1 = oracle.synth_gray(7, 0, w, h), oracle.synth_gray(7, 1, w, h)
    _, d0 = ex(g0)
    _, d1 = ex(g1)
    m = ex.match(d1, d0, k=1)
    mo = oracle.match(d1, d0)
    assert np.array_equal(m.view(np.uint8), mo.view(np.uint8))
    # frontend's distance < 50 loop (frontend.cpp:1126-1132)
    mf = ex.match(d1, d0, k=1, max_dist=50.0)
    assert np.array_equal(mf.view(np.uint8), mo[mo["distance"] < 50.0].view(np.uint8))
    # knnMatch k=2 and ratio 0.75
    k2 = ex.match(d1, d0, k=2).reshape(-1, 2)
    ko = oracle.knn2(d1, d0)
    assert np.array_equal(k2.view(np.uint8), ko.view(np.uint8))
    good = ex.match(d1, d0, k=2, ratio=0.75)
    keep = ko[:, 0]["distance"] < np.float32(0.75) * ko[:, 1]["distance"]
    assert np.array_equal(good.view(np.uint8), ko[keep, 0].view(np.uint8))


def test_both_match_engines_ragged_sizes(built, oracle):
    """ORBX_OPT_MATCH_MMA: the tensor-memory matcher (tcgen05, 3), the mma.sync int8 GEMM matcher (2) and the LOP3/POPC matcher (0) against
    BFMatcher's restatement on ragged problem sizes (fewer rows than one MMA tile, sizes straddling the 16 / 128-query and 64 / 128-row staging
    units, several splits, duplicates -> lowest-index ties, all-ones / all-zeros descriptors: the extremes of the signed key arithmetic)"""
    import orbx
    rng = np.random.default_rng(11)
    e = orbx.ORBextractor(max_width=320, max_height=240)
    try:
        for nq, nt in [(1, 1), (1, 7), (3, 9), (16, 8), (17, 63), (33, 65), (129, 130), (500, 1000), (1000, 777), (64, 4096), (257, 40000), (1500, 120000)]:          # the last: tens of 128-row tiles per CTA in the tensor-memory kernel
            q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
            t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
            if nt > 4:
                t[nt - 1] = t[1]; q[0] = t[1]                       # exact duplicate at both ends of the train set: tie at distance 0
                t[nt // 2] = t[1]
            if nt > 8 and nq > 2:
                t[3] = 255; t[4] = 0; q[1] = 255; q[2] = 0           # popcounts 256 and 0 on both sides
            want1, want2 = oracle.match(q, t), oracle.knn2(q, t)
            for mma in (3, 2, 0):                                    # 3 / 2 = force the tensor-core kernels whatever the problem size
                e.set_match_mma(mma)
                m = e.match(q, t, k=1)
                assert np.array_equal(m.view(np.uint8), want1.view(np.uint8)), (nq, nt, mma)
                k2 = e.match(q, t, k=2).reshape(-1, 2)
                assert np.array_equal(k2["trainIdx"], want2["trainIdx"]) and np.array_equal(k2["distance"], want2["distance"]), (nq, nt, mma)
                good = e.match(q, t, k=1, max_dist=110.0)
                assert np.array_equal(good.view(np.uint8), want1[want1["distance"] < 110.0].view(np.uint8)), (nq, nt, mma)
    finally:
        e.close()


def test_depth_filter(ex, oracle):
    w, h = 1280, 720
    g = oracle.synth_gray(9, 2, w, h)
    d = oracle.synth_depth(9, 2, w, h)
    kps, desc = ex(g)
    fk, fd = ex(g, depth=d)
    ok, od, _ = oracle.filter_depth(kps, desc, d)
    assert 0 < len(ok) < len(kps)
    assert np.array_equal(fk.view(np.uint8), ok.view(np.uint8))
    assert np.array_equal(fd, od)


def test_device_trig_and_atan2_pinned(ex, oracle):
    """The only floating-point steps of the path must round like the reference's x86-64 build: the CUDA restatement of glibc's
    cosf / sinf against this machine's libm over EVERY fp32 angle in [0, 360] degrees (1.1e9 values, by checksum of the result bit
    patterns, and element-wise on a sample), and cv::fastAtan2 against the oracle (itself pinned to cv2 and the golden vectors)."""
    first, last = 0x00000000, int(np.float32(360.0).view(np.uint32))
    assert ex.test_trig_checksum(first, last) == oracle.trig_checksum(first, last)
    rng = np.random.default_rng(6)
    x = np.concatenate([rng.uniform(0, 6.2832, 20000), [0.0, 1e-30, 1e-5, np.pi / 4, np.pi / 2, np.pi, 2 * np.pi, 6.283185]]).astype(np.float32)
    c, s_ = ex.test_trig(x)
    wc = np.array([oracle.cosf(float(v)) for v in x], np.float32)
    ws = np.array([oracle.sinf(float(v)) for v in x], np.float32)
    assert np.array_equal(c.view(np.uint32), wc.view(np.uint32)) and np.array_equal(s_.view(np.uint32), ws.view(np.uint32))
    ys = np.concatenate([rng.integers(-2900000, 2900000, 20000), [0, 0, 1, -1, 5, -5, 0, 7]]).astype(np.float32)
    xs = np.concatenate([rng.integers(-2900000, 2900000, 20000), [0, 1, 0, 0, 5, 5, -3, -7]]).astype(np.float32)
    got = ex.test_atan2(ys, xs)
    want = np.array([oracle.fast_atan2(float(a), float(b)) for a, b in zip(ys, xs)], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_semantic_box_filter(ex, oracle):
    """BASELINE configs[4] / reference backend.cpp:1011-1029, 746-751: keypoints whose pixel falls in a detection box of a dropped class
    are removed (first containing box decides, edges inclusive, fp64 compares); combined with the depth filter, order preserved."""
    w, h = 1280, 720
    g = oracle.synth_gray(9, 3, w, h)
    d = oracle.synth_depth(9, 3, w, h)
    kps, desc = ex(g)
    boxes = np.zeros(5, oracle.BOX_DTYPE)
    boxes["cx"], boxes["cy"] = [300.0, 640.0, 700.0, 1000.5, 100.0], [200.0, 360.0, 400.0, 600.25, 650.0]
    boxes["w"], boxes["h"] = [220.0, 400.0, 500.0, 301.0, 150.0], [180.0, 300.0, 380.0, 200.5, 120.0]
    boxes["class_id"] = [0, 2, 0, 5, 63]                                  # box 2 overlaps box 1: the first containing box wins
    # a keypoint exactly on an edge of box 0 must count as inside
    kx, ky = float(kps["x"][0]), float(kps["y"][0])
    boxes["cx"][0], boxes["w"][0] = kx - 110.0, 220.0                      # right edge == kx
    boxes["cy"][0], boxes["h"][0] = ky, 180.0

    def want(mask, depth):
        k, dsc = (kps, desc) if depth is None else oracle.filter_depth(kps, desc, depth)[:2]
        keep = []
        for i in range(len(k)):
            c = oracle.categorize(float(k["x"][i]), float(k["y"][i]), boxes)
            keep.append(not (0 <= c < 64 and (mask >> c) & 1))
        keep = np.array(keep, bool)
        return k[keep], dsc[keep]

    assert oracle.categorize(kx, ky, boxes) == 0
    for mask in (1 << 0, 1 << 2, (1 << 0) | (1 << 5) | (1 << 63), 0):
        for depth in (None, d):
            gk, gd = ex(g, depth=depth, boxes=boxes, drop_class_mask=mask)
            wk, wd = want(mask, depth)
            assert np.array_equal(gk.view(np.uint8), wk.view(np.uint8)) and np.array_equal(gd, wd), (mask, depth is not None)
            if mask:
                assert 0 < len(gk) < len(kps)
    gk, gd = ex(g, boxes=boxes[:0], drop_class_mask=1)                     # no boxes: nothing is labeled, nothing dropped
    assert np.array_equal(gk.view(np.uint8), kps.view(np.uint8))


def test_bgr_ingest(ex, oracle):
    """cvtColor(BGR2GRAY) on the device + extraction == the reference's host-side conversion followed by extraction."""
    import torch
    import orbx
    w, h = 741, 417
    g = oracle.synth_gray(13, 0, w, h)
    rng = np.random.default_rng(4)
    bgr = np.clip(np.stack([g, g, g], 2).astype(np.int32) + rng.integers(-12, 13, (h, w, 3)), 0, 255).astype(np.uint8)
    gray = oracle.bgr2gray(bgr)
    ref = oracle.COracle().extract(gray)
    k, d = ex.extract_bgr(bgr)
    assert np.array_equal(k.view(np.uint8), ref["kps"].view(np.uint8)) and np.array_equal(d, ref["desc"])
    assert np.array_equal(ex.pyramid_level(0), gray)
    # the batched device conversion
    dev = torch.device("cuda", 0)
    src = torch.from_numpy(np.stack([bgr, bgr[::-1].copy()])).to(dev)
    gp = (w + 3) & ~3
    dst = torch.zeros((2, h, gp), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ex._check(ex.L.orbx_bgr2gray_device(ex.handle, src.data_ptr(), 2, w, h, 3 * w, 3 * w * h, dst.data_ptr(), gp, gp * h))
    ex.sync()
    out = dst.cpu().numpy()
    assert np.array_equal(out[0, :, :w], gray) and np.array_equal(out[1, :, :w], gray[::-1])


def test_harris_responses(ex, oracle):
    """Harris score (HARRIS_SCORE of ORBextractor.hpp:48 / cv::ORB's HarrisResponses) at the retained keypoints of every level: within 1e-4
    relative of the oracle (north_star's bar; the arithmetic is integer sums + 7 fp32 operations, so it is in fact bit-identical)."""
    w, h = 1280, 720
    g = oracle.synth_gray(7, 0, w, h)
    orc = oracle.COracle()
    ref = orc.extract(g, trace=True)
    kps, _ = ex(g)
    for l in range(8):
        sel = kps[kps["octave"] == l]
        xy = np.stack([np.rint(sel["x"] / orc.scale[l]), np.rint(sel["y"] / orc.scale[l])], 1).astype(np.int32)
        got = ex.harris_responses(l, xy)
        want = np.array([oracle.harris_response(ref["pyramid"][l], x, y) for x, y in xy], np.float32)
        assert len(got) > 30 and np.all(np.abs(got - want) <= 1e-4 * np.abs(want))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_corner_dense_image_capacity(built, oracle):
    """White noise makes almost every pixel a FAST corner: the default candidate lists overflow and the call says so
    (ORBX_E_CAPACITY, nothing truncated silently); with cand_divisor=1 the same frame is bit-exact again."""
    import orbx
    w, h = 640, 480
    g = np.random.default_rng(5).integers(0, 256, (h, w), dtype=np.uint8)
    ref = oracle.COracle().extract(g, trace=True)
    assert sum(len(c) for c in ref["cands"]) > 640 * 480 // 16
    e = orbx.ORBextractor(max_width=w, max_height=h)
    try:
        with pytest.raises(orbx.OrbxError) as err:
            e(g)
        assert err.value.status == orbx.E_CAPACITY
        k2, _ = e(oracle.synth_gray(3, 0, w, h))                      # the handle stays usable
        assert len(k2) > 500
    finally:
        e.close()
    e = orbx.ORBextractor(max_width=w, max_height=h, cand_divisor=1)
    try:
        kps, desc = e(g, cap=8192)
        assert np.array_equal(kps.view(np.uint8), ref["kps"].view(np.uint8)) and np.array_equal(desc, ref["desc"])
        for l in range(8):
            assert len(e.candidates(l)) == len(ref["cands"][l])
    finally:
        e.close()


def test_quadtree_alone_tie_heavy(ex, oracle):
    """DistributeOctTree on its own (orbx_test_quadtree) against the oracle on candidate sets built to produce many comparator ties
    (equal counts and equal UL.x), which is where the block-parallel restatement of std::sort must land exactly like libstdc++."""
    rng = np.random.default_rng(77)
    cases = [(3000, 1248, 688, 217), (1800, 1035, 568, 181), (900, 325, 209, 60), (5000, 1248, 688, 500), (400, 600, 300, 151), (2500, 857, 468, 1000)]
    for n, W, H, N in cases:
        for rep in range(3):
            # points on a coarse lattice with jitter: many nodes end up with the same count
            gx, gy = rng.integers(0, W // 8, n) * 8 + rng.integers(0, 2, n), rng.integers(0, H // 8, n) * 8 + rng.integers(0, 2, n)
            pts = sorted(set(zip(np.minimum(gx, W - 1).tolist(), np.minimum(gy, H - 1).tolist())))
            sc = rng.integers(7, 60, len(pts))
            ncols = max(1, W // 35)
            wcell, hcell = int(np.ceil(W / ncols)), int(np.ceil(H / max(1, H // 35)))
            # the reference feeds the tree in cell-row-major, then raster order: reproduce that order for the oracle
            order = sorted(range(len(pts)), key=lambda i: ((pts[i][1] - 3) // hcell if pts[i][1] >= 3 else 0, (pts[i][0] - 3) // wcell if pts[i][0] >= 3 else 0, pts[i][1], pts[i][0]))
            c = np.zeros(len(pts), oracle.CAND_DTYPE)
            c["x"], c["y"], c["score"] = [pts[i][0] for i in order], [pts[i][1] for i in order], [int(sc[i]) for i in order]
            want = oracle.distribute_octtree(c, 16, 16 + W, 16, 16 + H, N)
            xys = np.stack([c["x"], c["y"], c["score"]], 1).astype(np.int32)
            got = ex.test_quadtree(xys[rng.permutation(len(xys))], W, H, wcell, hcell, ncols, N)      # device input order is arbitrary
            assert got.tolist() == np.stack([want["x"], want["y"], want["score"]], 1).tolist(), (n, W, H, N, rep)


def test_large_batch_uses_narrow_quadtree_blocks(built, oracle):
    """More than one (frame, level) CTA per SM switches the quadtree to 256-thread blocks: same results."""
    import torch
    import orbx
    w, h, nb, CAP = 320, 240, 24, 1024
    frames = np.stack([oracle.synth_gray(40, f, w, h) for f in range(nb)])
    ex2 = orbx.ORBextractor(max_width=w, max_height=h, max_batch=nb, max_keypoints=CAP)
    try:
        dev = torch.device("cuda", 0)
        g = torch.from_numpy(frames).to(dev)
        kps = torch.zeros((nb, CAP, 28), dtype=torch.uint8, device=dev)
        desc = torch.zeros((nb, CAP, 32), dtype=torch.uint8, device=dev)
        cnt = torch.zeros(nb, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        ex2.extract_batch_device(g.data_ptr(), nb, w, h, w, w * h, kps.data_ptr(), desc.data_ptr(), CAP, cnt.data_ptr())
        ex2.sync()
        kk, dd, cc = kps.cpu().numpy().view(orbx.KP_DTYPE).reshape(nb, CAP), desc.cpu().numpy(), cnt.cpu().numpy()
        orc = oracle.COracle()
        for f in (0, 7, 23):
            ref = orc.extract(frames[f])
            assert cc[f] == len(ref["kps"]) and np.array_equal(kk[f, :cc[f]].view(np.uint8), ref["kps"].view(np.uint8)) and np.array_equal(dd[f, :cc[f]], ref["desc"])
    finally:
        ex2.close()
