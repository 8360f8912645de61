"""GPU parity of the frontend's feature culling for the backend (reference frontend.cpp:1168-1218) against the
oracle: matched keypoints in match order, then the best unmatched ones by response in the reference's std::sort order."""
import ctypes as ct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex(built):
    import orbx
    e = orbx.ORBextractor(max_width=640, max_height=480, max_batch=1)
    yield e
    e.close()


def _want(oracle, kps, desc, q, max_new=200, min_response=50.0):
    idx = oracle.cull_keyframe(kps["response"], q, max_new, min_response)
    return kps[idx], desc[idx], idx


def test_cull_on_extracted_frames(ex, oracle):
    """The call sequence of Frontend::syncCallback on two frames: extract + filterDepth, match vs the previous frame, distance filter,
    (the RANSAC mask is the caller's: here every second good match), then the culling rule."""
    w, h = 640, 480
    orc = oracle.COracle()
    r0, r1 = orc.extract(oracle.synth_gray(9, 0, w, h)), orc.extract(oracle.synth_gray(9, 1, w, h))
    k1, d1, _ = oracle.filter_depth(r1["kps"], r1["desc"], oracle.synth_depth(9, 1, w, h))
    k0, d0, _ = oracle.filter_depth(r0["kps"], r0["desc"], oracle.synth_depth(9, 0, w, h))
    m = oracle.match(d1, d0)
    good = m[m["distance"] < 50.0]["queryIdx"]
    for q in (good, good[::2], good[:0]):
        gk, gd, gi = ex.cull_keyframe(k1, d1, q)
        wk, wd, wi = _want(oracle, k1, d1, q)
        assert np.array_equal(gi, wi) and np.array_equal(gk.view(np.uint8), wk.view(np.uint8)) and np.array_equal(gd, wd)
        assert len(gi) > len(q)


def test_cull_tie_heavy_and_edges(ex, oracle):
    """Responses drawn from a handful of integers around the 50 cut: the std::sort tie order decides which features make the 200."""
    rng = np.random.default_rng(41)
    for n, nm, lo, hi in [(1, 0, 60, 61), (1, 1, 60, 61), (16, 5, 45, 56), (17, 0, 45, 56), (33, 33, 45, 56), (300, 100, 45, 56), (1000, 593, 7, 120),
                          (1000, 0, 50, 52), (1280, 400, 49, 52), (4096, 77, 48, 53), (8192, 1000, 7, 255)]:
        kps = np.zeros(n, oracle.KP_DTYPE)
        kps["x"], kps["y"] = rng.uniform(16, 1264, n).astype(np.float32), rng.uniform(16, 704, n).astype(np.float32)
        kps["response"] = rng.integers(lo, hi, n).astype(np.float32)
        kps["octave"] = rng.integers(0, 8, n)
        desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        q = rng.permutation(n)[:nm].astype(np.int32)
        for max_new, min_resp in ((200, 50.0), (5, 0.0), (0, 50.0), (10000, 51.5)):
            gk, gd, gi = ex.cull_keyframe(kps, desc, q, max_new, min_resp)
            wk, wd, wi = _want(oracle, kps, desc, q, max_new, min_resp)
            assert np.array_equal(gi, wi), (n, nm, max_new, min_resp)
            assert np.array_equal(gk.view(np.uint8), wk.view(np.uint8)) and np.array_equal(gd, wd)
    # empty input, bad index, too small an output, too many keypoints
    import orbx
    gk, gd, gi = ex.cull_keyframe(np.zeros(0, oracle.KP_DTYPE), np.zeros((0, 32), np.uint8), np.zeros(0, np.int32))
    assert len(gk) == 0 and len(gi) == 0
    kps = np.zeros(10, oracle.KP_DTYPE); kps["response"] = 80
    desc = np.zeros((10, 32), np.uint8)
    with pytest.raises(orbx.OrbxError) as e:
        ex.cull_keyframe(kps, desc, np.array([10], np.int32))
    assert e.value.status == orbx.E_INVALID
    with pytest.raises(orbx.OrbxError) as e:
        ex.cull_keyframe(kps, desc, np.array([1, 2], np.int32), cap=5)
    assert e.value.status == orbx.E_CAPACITY
    with pytest.raises(orbx.OrbxError) as e:
        ex.cull_keyframe(np.zeros(8193, oracle.KP_DTYPE), np.zeros((8193, 32), np.uint8), np.zeros(0, np.int32))
    assert e.value.status == orbx.E_CAPACITY
    gk, _, gi = ex.cull_keyframe(kps, desc, np.array([3], np.int32))         # the handle stays usable
    assert gi.tolist() == oracle.cull_keyframe(kps["response"], np.array([3], np.int32)).tolist()


def test_cull_device_variant(ex, oracle):
    """orbx_cull_keyframe_device on device-resident lists (asynchronous on the handle's stream), and a bad index surfacing at the sync."""
    import torch
    import orbx
    rng = np.random.default_rng(3)
    n, nm = 900, 420
    kps = np.zeros(n, oracle.KP_DTYPE)
    kps["response"] = rng.integers(30, 70, n).astype(np.float32)
    kps["x"] = np.arange(n, dtype=np.float32)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    q = rng.permutation(n)[:nm].astype(np.int32)
    dev = torch.device("cuda", 0)
    cap = nm + 200
    d_k = torch.from_numpy(kps.view(np.uint8).reshape(n, 28)).to(dev)
    d_d = torch.from_numpy(desc).to(dev)
    d_q = torch.from_numpy(q).to(dev)
    o_k = torch.zeros((cap, 28), dtype=torch.uint8, device=dev)
    o_d = torch.zeros((cap, 32), dtype=torch.uint8, device=dev)
    o_i = torch.zeros(cap, dtype=torch.int32, device=dev)
    o_n = torch.zeros(1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    L = ex.L
    ex._check(L.orbx_cull_keyframe_device(ex.handle, d_k.data_ptr(), d_d.data_ptr(), n, d_q.data_ptr(), nm, 200, ct.c_float(50.0),
                                          o_k.data_ptr(), o_d.data_ptr(), o_i.data_ptr(), cap, o_n.data_ptr()))
    ex.sync()
    m = int(o_n.cpu()[0])
    wk, wd, wi = _want(oracle, kps, desc, q)
    assert m == len(wi) and np.array_equal(o_i.cpu().numpy()[:m], wi)
    assert np.array_equal(o_k.cpu().numpy()[:m].reshape(-1), wk.view(np.uint8)) and np.array_equal(o_d.cpu().numpy()[:m], wd)
    bad = torch.tensor([5, 900], dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ex._check(L.orbx_cull_keyframe_device(ex.handle, d_k.data_ptr(), d_d.data_ptr(), n, bad.data_ptr(), 2, 200, ct.c_float(50.0),
                                          o_k.data_ptr(), o_d.data_ptr(), o_i.data_ptr(), cap, o_n.data_ptr()))
    with pytest.raises(orbx.OrbxError) as e:
        ex.sync()
    assert e.value.status == orbx.E_INVALID
    ex.sync()                                                                 # the flag is cleared, the handle stays usable
