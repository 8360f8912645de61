"""GPU: the entry points of include/orbx.h that the other suites do not reach — device-memory helpers, the asynchronous extraction
batches, matching of arbitrary frame pairs on device-resident descriptor sets, and the device-pointer variants of the landmark
association — each against the oracle or against its host-pointer twin."""
import ctypes as ct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, N, SEED = 640, 480, 5, 33


@pytest.fixture(scope="module")
def frames_ref(oracle):
    orc = oracle.COracle()
    frames = np.stack([oracle.synth_gray(SEED, f, W, H) for f in range(N)])
    return frames, [orc.extract(frames[f]) for f in range(N)]


def test_version_and_device_memory_helpers(built):
    import orbx
    ex = orbx.ORBextractor(max_width=W, max_height=H)
    try:
        L = ex.L
        L.orbx_version.restype = ct.c_char_p
        assert L.orbx_version().decode().startswith("orbx")
        L.orbx_alloc_device.restype = ct.c_void_p
        L.orbx_alloc_device.argtypes = [ct.c_void_p, ct.c_size_t]
        L.orbx_free_device.argtypes = [ct.c_void_p, ct.c_void_p]
        src = np.arange(100000, dtype=np.uint32)
        dst = np.zeros_like(src)
        d = L.orbx_alloc_device(ex.handle, src.nbytes)
        assert d
        ex._check(L.orbx_copy_to_device(ex.handle, ct.c_void_p(d), src.ctypes.data_as(ct.c_void_p), ct.c_size_t(src.nbytes)))
        ex._check(L.orbx_copy_to_host(ex.handle, dst.ctypes.data_as(ct.c_void_p), ct.c_void_p(d), ct.c_size_t(src.nbytes)))
        assert np.array_equal(src, dst)
        L.orbx_free_device(ex.handle, ct.c_void_p(d))
    finally:
        ex.close()


def test_extract_batch_submit_two_in_flight(built, frames_ref):
    """orbx_extract_batch_submit / orbx_batch_wait: batch k+1 is submitted before batch k is collected."""
    import orbx
    frames, ref = frames_ref
    B, CAP = 2, 2048
    pg = orbx.PinnedArray(frames.shape, np.uint8)
    pg.array[...] = frames
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=B)
    outs = [dict(kps=orbx.PinnedArray((B, CAP), orbx.KP_DTYPE), desc=orbx.PinnedArray((B, CAP, 32), np.uint8), counts=orbx.PinnedArray((B,), np.int32))
            for _ in range(2)]
    try:
        L = ex.L

        def submit(a, b, o):
            t = ct.c_int32()
            ex._check(L.orbx_extract_batch_submit(ex.handle, pg.array[a:b].ctypes.data_as(ct.c_void_p), b - a, W, H, ct.c_size_t(W), None, ct.c_size_t(0),
                                                  ct.c_void_p(o["kps"].ptr), ct.c_void_p(o["desc"].ptr), CAP, ct.c_void_p(o["counts"].ptr), ct.byref(t)))
            return t.value

        def check(o, a, b):
            for i in range(b - a):
                r = ref[a + i]
                n = int(o["counts"].array[i])
                assert n == len(r["kps"])
                assert np.array_equal(o["kps"].array[i, :n].view(np.uint8), r["kps"].view(np.uint8)) and np.array_equal(o["desc"].array[i, :n], r["desc"])

        spans = [(f0, min(N, f0 + B)) for f0 in range(0, N, B)]
        tickets = []
        for k, (a, b) in enumerate(spans):
            tickets.append(submit(a, b, outs[k & 1]))
            if k >= 1:
                ex.batch_wait(tickets[k - 1])
                check(outs[(k - 1) & 1], *spans[k - 1])
        ex.batch_wait(tickets[-1])
        check(outs[(len(spans) - 1) & 1], *spans[-1])
    finally:
        ex.close()
        pg.close()
        for o in outs:
            for v in o.values():
                v.close()


def test_match_pairs_device(built, oracle, frames_ref):
    """orbx_match_pairs_device on the descriptor sets orbx_extract_batch_device left in HBM: arbitrary (query frame, train frame) pairs,
    k = 1 with and without the distance threshold, k = 2 with the ratio test."""
    import torch
    import orbx
    frames, ref = frames_ref
    dev = torch.device("cuda", 0)
    CAP = 1536
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=N, max_keypoints=CAP)
    try:
        g = torch.from_numpy(frames).to(dev)
        kps = torch.zeros((N, CAP, 28), dtype=torch.uint8, device=dev)
        desc = torch.zeros((N, CAP, 32), dtype=torch.uint8, device=dev)
        cnt = torch.zeros(N, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        ex.extract_batch_device(g.data_ptr(), N, W, H, W, W * H, kps.data_ptr(), desc.data_ptr(), CAP, cnt.data_ptr())
        ex.sync()
        qf = np.array([1, 2, 4, 0, 3], np.int32)
        tf = np.array([0, 1, 0, 4, 3], np.int32)
        for k, max_dist, ratio in ((1, 0.0, 0.0), (1, 50.0, 0.0), (2, 0.0, 0.75)):
            out = torch.zeros((len(qf), CAP * k, 16), dtype=torch.uint8, device=dev)
            nout = torch.zeros(len(qf), dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            ex._check(ex.L.orbx_match_pairs_device(ex.handle, desc.data_ptr(), cnt.data_ptr(), CAP, qf.ctypes.data_as(ct.c_void_p), tf.ctypes.data_as(ct.c_void_p),
                                                   len(qf), k, ct.c_float(max_dist), ct.c_float(ratio), out.data_ptr(), nout.data_ptr()))
            ex.sync()
            o = out.cpu().numpy().view(orbx.DM_DTYPE).reshape(len(qf), CAP * k)
            no = nout.cpu().numpy()
            for p in range(len(qf)):
                q, t = ref[qf[p]]["desc"], ref[tf[p]]["desc"]
                if k == 1:
                    want = oracle.match(q, t)
                    if max_dist > 0:
                        want = want[want["distance"] < max_dist]
                else:
                    ko = oracle.knn2(q, t)
                    want = ko[ko[:, 0]["distance"] < np.float32(ratio) * ko[:, 1]["distance"], 0]
                assert no[p] == len(want), (k, max_dist, p, no[p], len(want))
                assert np.array_equal(o[p, :no[p]].view(np.uint8), want.view(np.uint8)), (k, max_dist, p)
    finally:
        ex.close()


def test_db_device_pointer_variants(built, oracle):
    """orbx_db_set_positions_device / orbx_db_associate_device give what their host-pointer twins give (those are checked against the oracle
    in test_gpu_assoc.py)."""
    import torch
    import orbx
    rng = np.random.default_rng(4)
    n, nq = 20000, 300
    rows = oracle.synth_descriptors(55, 0, n)
    R, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    t = rng.standard_normal(3) * 0.2
    fx, fy, cx, cy = 615.3, 615.9, 640.2, 360.4
    pc = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1.2, 1.2, n), rng.uniform(0.4, 6.0, n)], 1)
    pos = (pc @ R.T + t).astype(np.float32)
    src = rng.integers(0, n, nq)
    q = rows[src].copy()
    q[::4, 7] ^= 0x3C
    qpx = (np.stack([oracle.reproject(pos[j], R, t, fx, fy, cx, cy) for j in src]) + rng.normal(0, 2.0, (nq, 2))).astype(np.float32)
    dev = torch.device("cuda", 0)
    ex = orbx.ORBextractor(max_width=W, max_height=H)
    db = orbx.LandmarkDB(ex, n)
    try:
        db.append(rows)
        db.set_positions(pos)
        pose = db.pose(R, t, fx, fy, cx, cy)
        want = db.associate(q, qpx, pose)
        assert (want["landmark"] >= 0).any()
        db2 = orbx.LandmarkDB(ex, n)
        try:
            db2.append(rows)
            d_pos, d_q, d_px = torch.from_numpy(pos).to(dev), torch.from_numpy(q).to(dev), torch.from_numpy(qpx).to(dev)
            d_out = torch.zeros((nq, 16), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            ex._check(ex.L.orbx_db_set_positions_device(db2._db, ct.c_int64(0), ct.c_int64(n), ct.c_void_p(d_pos.data_ptr())))
            ex._check(ex.L.orbx_db_associate_device(db2._db, ct.c_void_p(d_q.data_ptr()), ct.c_void_p(d_px.data_ptr()), nq, pose.ctypes.data_as(ct.c_void_p),
                                                    ct.c_float(50.0), ct.c_double(5.0), ct.c_void_p(d_out.data_ptr())))
            ex.sync()
            got = d_out.cpu().numpy().view(orbx.ASSOC_DTYPE).reshape(nq)
            assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
        finally:
            db2.close()
    finally:
        db.close()
        ex.close()


def test_two_handles_share_a_gpu_and_argument_validation(built, oracle):
    """(a) dynamic shared-memory opt-ins are per function and device — a second handle with smaller needs must not lower what
    the first one relies on; (b) an nfeatures whose quadtree node table cannot fit shared memory is refused at create with a message;
    (c) a bad depth step is ORBX_E_INVALID, not a device fault."""
    import orbx
    rng = np.random.default_rng(3)
    a = orbx.ORBextractor(max_width=W, max_height=H)
    b = orbx.ORBextractor(max_width=W, max_height=H)
    try:
        n_big, n_small = 8000, 64
        kb = np.zeros(n_big, orbx.KP_DTYPE); kb["response"] = rng.integers(0, 200, n_big).astype(np.float32)
        db = rng.integers(0, 256, (n_big, 32), dtype=np.uint8)
        want_big = oracle.cull_keyframe(kb["response"], np.arange(0, n_big, 7, dtype=np.int32))
        _, _, ia = a.cull_keyframe(kb, db, np.arange(0, n_big, 7, dtype=np.int32))          # handle A: large opt-in
        assert np.array_equal(ia, want_big)
        _, _, ib = b.cull_keyframe(kb[:n_small], db[:n_small], np.zeros(0, np.int32))       # handle B: small launch of the same kernel
        assert np.array_equal(ib, oracle.cull_keyframe(kb["response"][:n_small], np.zeros(0, np.int32)))
        _, _, ia2 = a.cull_keyframe(kb, db, np.arange(0, n_big, 7, dtype=np.int32))         # handle A again: must still launch
        assert np.array_equal(ia2, want_big)
        g = oracle.synth_gray(3, 0, W, H)
        depth = oracle.synth_depth(3, 0, W, H)
        kps = np.zeros(4096, orbx.KP_DTYPE); desc = np.zeros((4096, 32), np.uint8); n = ct.c_int32()
        st = a.L.orbx_extract_filtered(a.handle, g.ctypes.data_as(ct.c_void_p), W, H, W, depth.ctypes.data_as(ct.c_void_p), 2 * W - 1, None, 0, ct.c_uint64(0),
                                       kps.ctypes.data_as(ct.c_void_p), desc.ctypes.data_as(ct.c_void_p), 4096, ct.byref(n))
        assert st == orbx.E_INVALID                                                          # odd depth step
        st = a.L.orbx_extract_filtered(a.handle, g.ctypes.data_as(ct.c_void_p), W, H, W, depth.ctypes.data_as(ct.c_void_p), W, None, 0, ct.c_uint64(0),
                                       kps.ctypes.data_as(ct.c_void_p), desc.ctypes.data_as(ct.c_void_p), 4096, ct.byref(n))
        assert st == orbx.E_INVALID                                                          # depth rows shorter than 2 * width
        k_ok, _ = a(g, depth=depth)                                                          # the handle stays usable
        assert len(k_ok) > 100
    finally:
        a.close(); b.close()
    with pytest.raises(orbx.OrbxError) as err:
        orbx.ORBextractor(nfeatures=60000, max_width=W, max_height=H)
    assert err.value.status == orbx.E_UNSUPPORTED and "shared-memory" in str(err.value)
