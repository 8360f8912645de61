"""GPU parity on the BASELINE.json configurations AT THEIR STATED SIZES and through the code paths bench.py times.

  configs[1]/[2]  orbx_track_batch_device / orbx_extract_batch_device, 1280x720, batch 128, the production schedule (not ORBX_OPT_SERIAL):
                  every frame of the batch against the oracle, against the compiled reference (oracle/_ref) and against the committed
                  checksums of the compiled reference's output (tests/golden/ref_stream_1280x720.npz);
  configs[4]      per-frame YOLO box lists at batch, device and host entry points, incl. the stream (track) variant.
"""
import ctypes as ct
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, SEED, CAP = 1280, 720, 20261018, 1280          # bench.py's stream
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def _device_stream(ex, first, n):
    import torch
    dev = torch.device("cuda", 0)
    gray = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    depth = torch.empty((n, H, W), dtype=torch.int16, device=dev)
    ex._check(ex.L.orbx_synth_gray_device(ex.handle, SEED, first, n, W, H, gray.data_ptr(), W, W * H))
    ex._check(ex.L.orbx_synth_depth_device(ex.handle, SEED, first, n, W, H, depth.data_ptr(), 2 * W, 2 * W * H))
    ex.sync()
    return gray, depth


def _outputs(n):
    import torch
    dev = torch.device("cuda", 0)
    return dict(kps=torch.zeros((n, CAP, 28), dtype=torch.uint8, device=dev), desc=torch.zeros((n, CAP, 32), dtype=torch.uint8, device=dev),
                cnt=torch.zeros(n, dtype=torch.int32, device=dev), m=torch.zeros((n, CAP, 16), dtype=torch.uint8, device=dev),
                mc=torch.zeros(n, dtype=torch.int32, device=dev))


def _cpu_stream(oracle, first, n, with_boxes=False):
    """the reference's per-frame sequence on the CPU: extract -> filterDepth (-> box filter) -> match vs previous filtered + distance < 50"""
    frames = np.stack([oracle.synth_gray(SEED, first + f, W, H) for f in range(n)])
    kps, desc, cnt = oracle.COracle().extract_batch(frames, cap=2048)
    out, prev = [], None
    for f in range(n):
        k, d = kps[f, :cnt[f]], desc[f, :cnt[f]]
        fk, fd, _ = oracle.filter_depth(k, d, oracle.synth_depth(SEED, first + f, W, H))
        if with_boxes:
            fk, fd = oracle.filter_boxes(fk, fd, oracle.synth_boxes(SEED, first + f, W, H), 1)
        m = oracle.match(fd, prev) if prev is not None and len(prev) and len(fd) else np.zeros(0, oracle.DM_DTYPE)
        out.append(dict(raw_kps=k.copy(), raw_desc=d.copy(), kps=fk, desc=fd, good=m[m["distance"] < 50.0]))
        prev = fd
    return frames, out


def test_bench_path_batch128_every_frame(built, oracle, refx):
    """the call, the size, the batch and the schedule that bench.py times: orbx_track_batch_device, 1280x720, 128 frames, two
    consecutive steps (the second matches its frame 0 against the carried last frame of the first)"""
    import orbx
    import torch
    B = 128
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=B, max_keypoints=CAP)
    try:
        _, ref = _cpu_stream(oracle, 0, 2 * B)
        for step in range(2):
            gray, depth = _device_stream(ex, step * B, B)
            o = _outputs(B)
            ex._check(ex.L.orbx_track_batch_device(ex.handle, gray.data_ptr(), B, W, H, W, W * H, depth.data_ptr(), 2 * W, 2 * W * H,
                                                   o["kps"].data_ptr(), o["desc"].data_ptr(), CAP, o["cnt"].data_ptr(),
                                                   o["m"].data_ptr(), o["mc"].data_ptr(), ct.c_float(50.0)))
            ex.sync()
            kk, dd = o["kps"].cpu().numpy().view(orbx.KP_DTYPE).reshape(B, CAP), o["desc"].cpu().numpy()
            cc, mm, mc = o["cnt"].cpu().numpy(), o["m"].cpu().numpy().view(orbx.DM_DTYPE).reshape(B, CAP), o["mc"].cpu().numpy()
            for f in range(B):
                r = ref[step * B + f]
                assert cc[f] == len(r["kps"]), ("count", step, f)
                assert np.array_equal(kk[f, :cc[f]].view(np.uint8), r["kps"].view(np.uint8)), ("keypoints", step, f)
                assert np.array_equal(dd[f, :cc[f]], r["desc"]), ("descriptors", step, f)
                assert mc[f] == len(r["good"]) and np.array_equal(mm[f, :mc[f]].view(np.uint8), r["good"].view(np.uint8)), ("matches", step, f)
            assert mc[0] == (0 if step == 0 else len(ref[B]["good"]))
            del gray, depth, o
            torch.cuda.empty_cache()
    finally:
        ex.close()


def test_overlap_option_same_results(built, oracle):
    """ORBX_OPT_OVERLAP = 1 (two staggered half-batches on two streams; off by default because it measured slower): identical outputs,
    incl. the match of the second half's first frame against the first half's last and the carried frame across calls"""
    import orbx
    B = 40
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=B, max_keypoints=CAP)
    try:
        ex.set_overlap(True)
        _, ref = _cpu_stream(oracle, 0, 2 * B)
        for step in range(2):
            gray, depth = _device_stream(ex, step * B, B)
            o = _outputs(B)
            ex._check(ex.L.orbx_track_batch_device(ex.handle, gray.data_ptr(), B, W, H, W, W * H, depth.data_ptr(), 2 * W, 2 * W * H,
                                                   o["kps"].data_ptr(), o["desc"].data_ptr(), CAP, o["cnt"].data_ptr(),
                                                   o["m"].data_ptr(), o["mc"].data_ptr(), ct.c_float(50.0)))
            ex.sync()
            kk, dd = o["kps"].cpu().numpy().view(orbx.KP_DTYPE).reshape(B, CAP), o["desc"].cpu().numpy()
            cc, mm, mc = o["cnt"].cpu().numpy(), o["m"].cpu().numpy().view(orbx.DM_DTYPE).reshape(B, CAP), o["mc"].cpu().numpy()
            for f in range(B):
                r = ref[step * B + f]
                assert cc[f] == len(r["kps"]) and np.array_equal(kk[f, :cc[f]].view(np.uint8), r["kps"].view(np.uint8)) and np.array_equal(dd[f, :cc[f]], r["desc"]), (step, f)
                if step or f:
                    assert mc[f] == len(r["good"]) and np.array_equal(mm[f, :mc[f]].view(np.uint8), r["good"].view(np.uint8)), (step, f)
    finally:
        ex.close()


def test_extract_batch128_vs_compiled_reference_and_golden_checksums(built, oracle, refx):
    """configs[2]'s unit of work (frame-parallel extraction, no depth) at batch 128: the reference's own compiled ORBextractor.cpp on a
    sample of frames, and the committed checksums of its output on frames 0..31 and 127"""
    import orbx
    B = 128
    gold = np.load(os.path.join(GOLD, "ref_stream_1280x720.npz"))
    assert int(gold["seed"]) == SEED
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=B, max_keypoints=CAP)
    try:
        gray, _ = _device_stream(ex, 0, B)
        o = _outputs(B)
        ex.extract_batch_device(gray.data_ptr(), B, W, H, W, W * H, o["kps"].data_ptr(), o["desc"].data_ptr(), CAP, o["cnt"].data_ptr())
        ex.sync()
        kk, dd, cc = o["kps"].cpu().numpy().view(orbx.KP_DTYPE).reshape(B, CAP), o["desc"].cpu().numpy(), o["cnt"].cpu().numpy()
        for i, f in enumerate(gold["frames"].tolist()):
            if f >= B:
                continue
            assert (int(cc[f]), crc(kk[f, :cc[f]]), crc(dd[f, :cc[f]])) == (int(gold["count"][i]), int(gold["kps_crc"][i]), int(gold["desc_crc"][i])), f
        for f in (0, 1, 63, 64, 100, 127):
            r = refx.extract(oracle.synth_gray(SEED, f, W, H))
            assert r["ret"] == cc[f] and np.array_equal(kk[f, :cc[f]].view(np.uint8), r["kps"].view(np.uint8)) and np.array_equal(dd[f, :cc[f]], r["desc"]), f
    finally:
        ex.close()


def test_configs4_boxes_at_batch_device_and_host(built, oracle):
    """configs[4]: per-frame YOLO boxes (4 per frame, class 0 = person dropped, box 2 of another class) in the batch calls — device
    entry, host entry (chunked pipeline, chunk boundaries inside the batch) and the stream variant with matching"""
    import orbx
    import torch
    B = 24
    _, ref = _cpu_stream(oracle, 0, B, with_boxes=True)
    frame_boxes = [oracle.synth_boxes(SEED, f, W, H) for f in range(B)]
    dropped = sum(len(oracle.filter_depth(r["raw_kps"], r["raw_desc"], oracle.synth_depth(SEED, f, W, H))[0]) - len(r["kps"]) for f, r in enumerate(ref))
    assert dropped > 50 * B // 10, "the synthetic boxes must actually remove keypoints"
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=B, max_keypoints=CAP, host_chunk=5)
    try:
        # --- device entry, stream variant ---
        gray, depth = _device_stream(ex, 0, B)
        boxes, off = ex.pack_frame_boxes(frame_boxes)
        d_boxes = torch.from_numpy(boxes.view(np.uint8).reshape(-1)).to("cuda:0")
        d_off = torch.from_numpy(off).to("cuda:0")
        o = _outputs(B)
        ex._check(ex.L.orbx_track_batch_boxes_device(ex.handle, gray.data_ptr(), B, W, H, W, W * H, depth.data_ptr(), 2 * W, 2 * W * H,
                                                     d_boxes.data_ptr(), d_off.data_ptr(), len(boxes), ct.c_uint64(1),
                                                     o["kps"].data_ptr(), o["desc"].data_ptr(), CAP, o["cnt"].data_ptr(),
                                                     o["m"].data_ptr(), o["mc"].data_ptr(), ct.c_float(50.0)))
        ex.sync()
        kk, dd = o["kps"].cpu().numpy().view(orbx.KP_DTYPE).reshape(B, CAP), o["desc"].cpu().numpy()
        cc, mm, mc = o["cnt"].cpu().numpy(), o["m"].cpu().numpy().view(orbx.DM_DTYPE).reshape(B, CAP), o["mc"].cpu().numpy()
        for f in range(B):
            r = ref[f]
            assert cc[f] == len(r["kps"]) and np.array_equal(kk[f, :cc[f]].view(np.uint8), r["kps"].view(np.uint8)) and np.array_equal(dd[f, :cc[f]], r["desc"]), f
            assert mc[f] == len(r["good"]) and np.array_equal(mm[f, :mc[f]].view(np.uint8), r["good"].view(np.uint8)), f
        # --- device entry, extraction only (no depth): boxes alone ---
        o2 = _outputs(B)
        ex._check(ex.L.orbx_extract_batch_boxes_device(ex.handle, gray.data_ptr(), B, W, H, W, W * H, None, 0, 0,
                                                       d_boxes.data_ptr(), d_off.data_ptr(), len(boxes), ct.c_uint64(1),
                                                       o2["kps"].data_ptr(), o2["desc"].data_ptr(), CAP, o2["cnt"].data_ptr()))
        ex.sync()
        kk2, dd2, cc2 = o2["kps"].cpu().numpy().view(orbx.KP_DTYPE).reshape(B, CAP), o2["desc"].cpu().numpy(), o2["cnt"].cpu().numpy()
        for f in (0, 7, B - 1):
            wk, wd = oracle.filter_boxes(ref[f]["raw_kps"], ref[f]["raw_desc"], frame_boxes[f], 1)
            assert cc2[f] == len(wk) and np.array_equal(kk2[f, :cc2[f]].view(np.uint8), wk.view(np.uint8)) and np.array_equal(dd2[f, :cc2[f]], wd), f
        # --- host entry: chunks of 5 frames, so box ranges are re-based per chunk ---
        frames = gray.cpu().numpy()
        depths = depth.cpu().numpy().view(np.uint16)
        ex.track_reset()
        kps, desc, counts, matches, mcounts = ex.track_batch(frames, depths, cap=CAP, frame_boxes=frame_boxes, drop_class_mask=1)
        for f in range(B):
            r = ref[f]
            assert counts[f] == len(r["kps"]) and np.array_equal(kps[f, :counts[f]].view(np.uint8), r["kps"].view(np.uint8)) and np.array_equal(desc[f, :counts[f]], r["desc"]), f
            assert mcounts[f] == len(r["good"]) and np.array_equal(matches[f, :mcounts[f]].view(np.uint8), r["good"].view(np.uint8)), f
        # a mask that drops nothing, and ragged lists (frames without boxes)
        ragged = [frame_boxes[f] if f % 3 == 0 else frame_boxes[f][:0] for f in range(B)]
        kps, desc, counts = ex.extract_batch(frames, depth=None, cap=CAP, frame_boxes=ragged, drop_class_mask=1)
        for f in (0, 1, 2, 3, B - 1):
            wk, wd = (oracle.filter_boxes(ref[f]["raw_kps"], ref[f]["raw_desc"], ragged[f], 1) if len(ragged[f]) else (ref[f]["raw_kps"], ref[f]["raw_desc"]))
            assert counts[f] == len(wk) and np.array_equal(kps[f, :counts[f]].view(np.uint8), wk.view(np.uint8)) and np.array_equal(desc[f, :counts[f]], wd), f
        kps, desc, counts = ex.extract_batch(frames[:4], depth=None, cap=CAP, frame_boxes=frame_boxes[:4], drop_class_mask=2)
        for f in range(4):
            wk, _ = oracle.filter_boxes(ref[f]["raw_kps"], ref[f]["raw_desc"], frame_boxes[f], 2)
            assert counts[f] == len(wk)
        with pytest.raises(orbx.OrbxError):                       # decreasing offsets are refused, not read
            bad = off.copy(); bad[3] = bad[2] - 1
            ex._check(ex.L.orbx_extract_batch_boxes(ex.handle, frames.ctypes.data_as(ct.c_void_p), B, W, H, W, None, 0,
                                                    boxes.ctypes.data_as(ct.c_void_p), bad.ctypes.data_as(ct.c_void_p), ct.c_uint64(1),
                                                    kps.ctypes.data_as(ct.c_void_p), desc.ctypes.data_as(ct.c_void_p), CAP, counts.ctypes.data_as(ct.c_void_p)))
    finally:
        ex.close()


def test_filter_order_option_same_results(built, oracle):
    """ORBX_OPT_FILTER_FIRST: the default (depth / box filter on the selected positions, then descriptors of the survivors only) and the
    reference's order (describe every selected keypoint, then drop rows) give the same bytes as the CPU sequence — depth alone, depth +
    boxes, through the device and the host entry (pinned depth gathered in place), and agree on the capacity error"""
    import orbx
    import torch
    B = 12
    _, ref = _cpu_stream(oracle, 0, B, with_boxes=True)
    _, ref_d = _cpu_stream(oracle, 0, B)
    frame_boxes = [oracle.synth_boxes(SEED, f, W, H) for f in range(B)]
    ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=B, max_keypoints=CAP, host_chunk=5)
    try:
        gray, depth = _device_stream(ex, 0, B)
        boxes, off = ex.pack_frame_boxes(frame_boxes)
        d_boxes = torch.from_numpy(boxes.view(np.uint8).reshape(-1)).to("cuda:0")
        d_off = torch.from_numpy(off).to("cuda:0")
        frames, depths = gray.cpu().numpy(), depth.cpu().numpy().view(np.uint16)
        got = {}
        for first in (1, 0):
            ex.set_filter_first(bool(first))
            ex.track_reset()
            o = _outputs(B)
            ex._check(ex.L.orbx_track_batch_boxes_device(ex.handle, gray.data_ptr(), B, W, H, W, W * H, depth.data_ptr(), 2 * W, 2 * W * H,
                                                         d_boxes.data_ptr(), d_off.data_ptr(), len(boxes), ct.c_uint64(1),
                                                         o["kps"].data_ptr(), o["desc"].data_ptr(), CAP, o["cnt"].data_ptr(),
                                                         o["m"].data_ptr(), o["mc"].data_ptr(), ct.c_float(50.0)))
            ex.sync()
            kk, dd = o["kps"].cpu().numpy().view(orbx.KP_DTYPE).reshape(B, CAP), o["desc"].cpu().numpy()
            cc, mm, mc = o["cnt"].cpu().numpy(), o["m"].cpu().numpy().view(orbx.DM_DTYPE).reshape(B, CAP), o["mc"].cpu().numpy()
            for f in range(B):
                r = ref[f]
                assert cc[f] == len(r["kps"]) and np.array_equal(kk[f, :cc[f]].view(np.uint8), r["kps"].view(np.uint8)) and np.array_equal(dd[f, :cc[f]], r["desc"]), (first, f)
                assert mc[f] == len(r["good"]) and np.array_equal(mm[f, :mc[f]].view(np.uint8), r["good"].view(np.uint8)), (first, f)
            # host entry, depth alone (pinned depth is gathered in place: one PCIe read per selected keypoint in either order)
            ex.track_reset()
            kps, desc, counts, matches, mcounts = ex.track_batch(frames, depths, cap=CAP)
            for f in range(B):
                r = ref_d[f]
                assert counts[f] == len(r["kps"]) and np.array_equal(kps[f, :counts[f]].view(np.uint8), r["kps"].view(np.uint8)) and np.array_equal(desc[f, :counts[f]], r["desc"]), (first, f)
                assert mcounts[f] == len(r["good"]) and np.array_equal(matches[f, :mcounts[f]].view(np.uint8), r["good"].view(np.uint8)), (first, f)
            # single-frame host call
            k1, d1 = ex(frames[3], depth=depths[3])
            assert np.array_equal(k1.view(np.uint8), ref_d[3]["kps"].view(np.uint8)) and np.array_equal(d1, ref_d[3]["desc"]), first
            # an output capacity below the filtered count: the same error in either order
            small = min(len(r["kps"]) for r in ref_d) - 1
            with pytest.raises(orbx.OrbxError) as e:
                ex.extract_batch(frames[:2], depth=depths[:2], cap=small)
            got[first] = [str(e.value)]
            # the device entry with an output capacity below the filtered count: the kernels raise the capacity flag, counts read 0
            o = _outputs(2)
            ex._check(ex.L.orbx_extract_batch_device(ex.handle, gray.data_ptr(), 2, W, H, W, W * H, depth.data_ptr(), 2 * W, 2 * W * H,
                                                     o["kps"].data_ptr(), o["desc"].data_ptr(), small, o["cnt"].data_ptr()))
            with pytest.raises(orbx.OrbxError) as e:
                ex.sync()
            got[first].append(str(e.value))
            assert o["cnt"].cpu().numpy().tolist() == [0, 0]
        assert got[0] == got[1], got
    finally:
        ex.close()


def test_capacity_flag_of_one_call_does_not_leak_into_the_next(built, oracle):
    """a device capacity flag raised by a multi-chunk host batch must not surface in a later, valid call"""
    import orbx
    rng = np.random.default_rng(5)
    w, h = 640, 480
    noisy = np.stack([rng.integers(0, 256, (h, w), dtype=np.uint8) for _ in range(6)])       # corner-dense: overflows the candidate lists
    calm = np.stack([oracle.synth_gray(3, f, w, h) for f in range(2)])
    ex = orbx.ORBextractor(max_width=w, max_height=h, max_batch=4, host_chunk=2)
    try:
        with pytest.raises(orbx.OrbxError) as err:
            ex.extract_batch(noisy)
        assert err.value.status == orbx.E_CAPACITY
        kps, desc, counts = ex.extract_batch(calm[:1])            # one chunk, slot 0 only: must be clean
        want = oracle.COracle().extract(calm[0])
        assert counts[0] == len(want["kps"]) and np.array_equal(desc[0, :counts[0]], want["desc"])
        kps, desc, counts = ex.extract_batch(calm)
        assert counts[1] == len(oracle.COracle().extract(calm[1])["kps"])
    finally:
        ex.close()
