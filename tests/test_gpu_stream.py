"""GPU parity of the batch / stream entry points (host buffers through the 3-stream pipeline, pinned zero-copy depth,
device buffers) against the oracle running the reference's per-frame sequence:
extract -> filterDepth -> match(filtered, prev_filtered) + distance < 50 -> prev = filtered  (reference frontend.cpp:1094-1132, 1258-1259)."""
import ctypes as ct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, N, SEED = 640, 480, 11, 21


@pytest.fixture(scope="module")
def stream_ref(oracle):
    orc = oracle.COracle()
    frames = np.stack([oracle.synth_gray(SEED, f, W, H) for f in range(N)])
    depths = np.stack([oracle.synth_depth(SEED, f, W, H) for f in range(N)])
    ref, prev = [], None
    for f in range(N):
        r = orc.extract(frames[f])
        fk, fd, _ = oracle.filter_depth(r["kps"], r["desc"], depths[f])
        m = oracle.match(fd, prev) if prev is not None and len(prev) and len(fd) else np.zeros(0, oracle.DM_DTYPE)
        ref.append(dict(raw_kps=r["kps"], raw_desc=r["desc"], kps=fk, desc=fd, good=m[m["distance"] < 50.0]))
        prev = fd
    return frames, depths, ref


def _check_track(out, ref, n0=0):
    kps, desc, counts, matches, mcounts = out
    for i in range(len(counts)):
        r = ref[n0 + i]
        assert counts[i] == len(r["kps"]), ("count frame %d" % (n0 + i), counts[i], len(r["kps"]))
        assert np.array_equal(kps[i, :counts[i]].view(np.uint8), r["kps"].view(np.uint8)), "keypoints frame %d" % (n0 + i)
        assert np.array_equal(desc[i, :counts[i]], r["desc"]), "descriptors frame %d" % (n0 + i)
        assert mcounts[i] == len(r["good"]), ("matches frame %d" % (n0 + i), mcounts[i], len(r["good"]))
        assert np.array_equal(matches[i, :mcounts[i]].view(np.uint8), r["good"].view(np.uint8)), "matches frame %d" % (n0 + i)


@pytest.mark.parametrize("max_batch,chunk", [(4, 0), (8, 3), (16, 0), (1, 0)])
def test_track_batch_host_pageable(built, stream_ref, max_batch, chunk):
    import orbx
    frames, depths, ref = stream_ref
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=max_batch, host_chunk=chunk)
    try:
        # two calls: the second continues from the carried previous-frame descriptors
        _check_track(ex.track_batch(frames[:6], depths[:6]), ref, 0)
        _check_track(ex.track_batch(frames[6:], depths[6:]), ref, 6)
        # after a reset the first frame has no predecessor again (reference first frame, frontend.cpp:1277-1317)
        ex.track_reset()
        out = ex.track_batch(frames[3:5], depths[3:5])
        assert out[4][0] == 0 and out[4][1] == len(ref[4]["good"])
    finally:
        ex.close()


def test_track_batch_host_pinned_zero_copy_depth(built, stream_ref):
    import orbx
    frames, depths, ref = stream_ref
    pg, pd = orbx.PinnedArray(frames.shape, np.uint8), orbx.PinnedArray(depths.shape, np.uint16)
    pg.array[...] = frames
    pd.array[...] = depths
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=4)
    try:
        _check_track(ex.track_batch(pg.array, pd.array), ref, 0)
        # single-frame call with pinned depth (gathered in place)
        fk, fd = ex(pg.array[2], depth=pd.array[2])
        assert np.array_equal(fk.view(np.uint8), ref[2]["kps"].view(np.uint8)) and np.array_equal(fd, ref[2]["desc"])
    finally:
        ex.close()
        pg.close(); pd.close()


def test_track_batch_async_two_in_flight(built, stream_ref):
    """submit(k+1) before wait(k): results identical to the synchronous sequence, tickets collected in order."""
    import orbx
    frames, depths, ref = stream_ref
    B, CAP = 3, 2048
    pg, pd = orbx.PinnedArray(frames.shape, np.uint8), orbx.PinnedArray(depths.shape, np.uint16)
    pg.array[...] = frames
    pd.array[...] = depths
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=B)
    outs = [dict(kps=np.zeros((B, CAP), orbx.KP_DTYPE), desc=np.zeros((B, CAP, 32), np.uint8), counts=np.zeros(B, np.int32),
                 matches=np.zeros((B, CAP), orbx.DM_DTYPE), mcounts=np.zeros(B, np.int32)) for _ in range(2)]
    try:
        spans = [(f0, min(N, f0 + B)) for f0 in range(0, N, B)]
        tickets = []
        for k, (a, b) in enumerate(spans):
            tickets.append(ex.track_batch_submit(pg.array[a:b], pd.array[a:b], outs[k & 1]))
            if k >= 1:                                               # collect batch k-1 while batch k is in flight
                ex.batch_wait(tickets[k - 1])
                a0, b0 = spans[k - 1]
                o = outs[(k - 1) & 1]
                _check_track((o["kps"][:b0 - a0], o["desc"][:b0 - a0], o["counts"][:b0 - a0], o["matches"][:b0 - a0], o["mcounts"][:b0 - a0]), ref, a0)
        with pytest.raises(orbx.OrbxError):                          # a synchronous call while a ticket is outstanding is refused
            ex(frames[0])
        ex.batch_wait(tickets[-1])
        a0, b0 = spans[-1]
        o = outs[(len(spans) - 1) & 1]
        _check_track((o["kps"][:b0 - a0], o["desc"][:b0 - a0], o["counts"][:b0 - a0], o["matches"][:b0 - a0], o["mcounts"][:b0 - a0]), ref, a0)
        with pytest.raises(orbx.OrbxError):
            ex.batch_wait(12345)
        k0, _ = ex(frames[0])                                        # back to synchronous use
        assert len(k0) == len(ref[0]["raw_kps"])
    finally:
        ex.close()
        pg.close(); pd.close()


def test_extract_batch_host(built, stream_ref):
    import orbx
    frames, depths, ref = stream_ref
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=4)
    try:
        kps, desc, counts = ex.extract_batch(frames)
        for i in range(N):
            assert counts[i] == len(ref[i]["raw_kps"])
            assert np.array_equal(kps[i, :counts[i]].view(np.uint8), ref[i]["raw_kps"].view(np.uint8))
            assert np.array_equal(desc[i, :counts[i]], ref[i]["raw_desc"])
        kps, desc, counts = ex.extract_batch(frames, depth=depths)
        for i in range(N):
            assert counts[i] == len(ref[i]["kps"]) and np.array_equal(desc[i, :counts[i]], ref[i]["desc"])
        # capacity error is reported, not truncated silently
        with pytest.raises(orbx.OrbxError) as e:
            ex.extract_batch(frames[:2], cap=100)
        assert e.value.status == orbx.E_CAPACITY
    finally:
        ex.close()


def test_track_batch_device(built, stream_ref):
    import torch
    import orbx
    frames, depths, ref = stream_ref
    dev = torch.device("cuda", 0)
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=N, max_keypoints=1536)
    try:
        CAP = 1536
        g = torch.from_numpy(frames).to(dev)
        d = torch.from_numpy(depths.view(np.int16)).to(dev)
        kps = torch.zeros((N, CAP, 28), dtype=torch.uint8, device=dev)
        desc = torch.zeros((N, CAP, 32), dtype=torch.uint8, device=dev)
        cnt = torch.zeros(N, dtype=torch.int32, device=dev)
        m = torch.zeros((N, CAP, 16), dtype=torch.uint8, device=dev)
        mc = torch.zeros(N, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        ex._check(ex.L.orbx_track_batch_device(ex.handle, g.data_ptr(), N, W, H, W, W * H, d.data_ptr(), 2 * W, 2 * W * H,
                                               kps.data_ptr(), desc.data_ptr(), CAP, cnt.data_ptr(), m.data_ptr(), mc.data_ptr(), ct.c_float(50.0)))
        ex.sync()
        out = (kps.cpu().numpy().view(orbx.KP_DTYPE).reshape(N, CAP), desc.cpu().numpy(), cnt.cpu().numpy(),
               m.cpu().numpy().view(orbx.DM_DTYPE).reshape(N, CAP), mc.cpu().numpy())
        _check_track(out, ref, 0)
    finally:
        ex.close()


def test_empty_and_featureless_frames(built):
    import orbx
    ex = orbx.ORBextractor(max_width=320, max_height=240)
    try:
        assert ex(np.zeros((0, 0), np.uint8)) == -1                       # reference returns -1 (ORBextractor.cpp:1090-1091)
        k, d = ex(np.full((240, 320), 77, np.uint8))
        assert len(k) == 0 and d.shape == (0, 32)
        with pytest.raises(TypeError):
            ex(np.zeros((240, 320), np.float32))
        with pytest.raises(orbx.OrbxError):                                # larger than the arenas sized at create
            ex(np.zeros((480, 640), np.uint8))
        assert len(ex.match(np.zeros((0, 32), np.uint8), np.zeros((5, 32), np.uint8))) == 0
        m = ex.match(np.zeros((3, 32), np.uint8), np.zeros((0, 32), np.uint8))
        assert len(m) == 0
    finally:
        ex.close()


def test_pack_keyframe(built, stream_ref, oracle):
    """Keyframe.msg landmark / observation records (reference frontend.cpp:731-776): host call and batched device call vs the oracle."""
    import torch
    import orbx
    frames, depths, ref = stream_ref
    rng = np.random.default_rng(12)
    R, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    t = rng.standard_normal(3)
    fx, fy, cx, cy = 615.3, 615.9, 320.2, 240.4
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=4)
    try:
        for f in (0, 5):
            kps, desc = ref[f]["raw_kps"], ref[f]["raw_desc"]             # unfiltered list: the depth gate does the filtering
            want = oracle.pack_keyframe(kps, desc, depths[f], fx, fy, cx, cy, R, t)
            got = ex.pack_keyframe(kps, desc, depths[f], fx, fy, cx, cy, R, t)
            assert 0 < len(want) < len(kps)
            assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), "keyframe records frame %d" % f
        # batched, device-resident: straight from the extractor's outputs
        dev = torch.device("cuda", 0)
        nb, CAP = 3, 1536
        g = torch.from_numpy(frames[:nb]).to(dev)
        d = torch.from_numpy(depths[:nb].view(np.int16)).to(dev)
        kps = torch.zeros((nb, CAP, 28), dtype=torch.uint8, device=dev)
        desc = torch.zeros((nb, CAP, 32), dtype=torch.uint8, device=dev)
        cnt = torch.zeros(nb, dtype=torch.int32, device=dev)
        rec = torch.zeros((nb, CAP, 80), dtype=torch.uint8, device=dev)
        rcnt = torch.zeros(nb, dtype=torch.int32, device=dev)
        K = np.zeros(1, orbx.KFPARAMS_DTYPE)
        K["R"][0], K["t"][0] = R.reshape(9), t
        K["fx"], K["fy"], K["cx"], K["cy"] = fx, fy, cx, cy
        torch.cuda.synchronize()
        ex.extract_batch_device(g.data_ptr(), nb, W, H, W, W * H, kps.data_ptr(), desc.data_ptr(), CAP, cnt.data_ptr())
        ex._check(ex.L.orbx_pack_keyframe_device(ex.handle, nb, kps.data_ptr(), desc.data_ptr(), cnt.data_ptr(), CAP, d.data_ptr(), W, H, 2 * W, 2 * W * H,
                                                 K.ctypes.data_as(ct.c_void_p), rec.data_ptr(), rcnt.data_ptr(), CAP))
        ex.sync()
        rc, rr = rcnt.cpu().numpy(), rec.cpu().numpy().view(orbx.KF_DTYPE).reshape(nb, CAP)
        for f in range(nb):
            want = oracle.pack_keyframe(ref[f]["raw_kps"], ref[f]["raw_desc"], depths[f], fx, fy, cx, cy, R, t)
            assert rc[f] == len(want) and np.array_equal(rr[f, :rc[f]].view(np.uint8), want.view(np.uint8))
    finally:
        ex.close()


def test_stream_blocks_with_preamble(built, stream_ref):
    """A stream cut into rank blocks (sharding.stream_block): with the one-frame preamble every block reproduces the matches of
    the unsharded stream, including the pair that straddles the block boundary."""
    import orbx
    from orbx import sharding
    frames, depths, ref = stream_ref
    world = 3
    for rank in range(world):
        first, count, pre = sharding.stream_block(N, world, rank)
        ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=4)
        try:
            if pre is not None:
                ex.track_batch(frames[pre:pre + 1], depths[pre:pre + 1])        # outputs discarded: only the carried descriptors matter
            _check_track(ex.track_batch(frames[first:first + count], depths[first:first + count]), ref, first)
        finally:
            ex.close()
