"""The two FAST formulations (k_fast.cu: warp per cell; k_fast_dense.cu: whole-level tiles + per-corner NMS + minThFAST retry launch)
against each other and against the oracle's cell loop (oracle/orb_oracle.c: orc_fast_cells, reference ORBextractor.cpp:785-872),
on the frame contents that stress what differs between them: cells that need the retry (flat and low-contrast regions), tiles with more
pre-test survivors than the queue holds (noise), corners on cell and tile borders (checkerboards of several periods), partial edge tiles
(awkward sizes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(oracle, w, h, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    out = [oracle.synth_gray(seed, 0, w, h), oracle.synth_gray(seed, 1, w, h)]
    out.append(np.full((h, w), 97, np.uint8))                                                   # nothing anywhere: every cell retries, finds nothing
    out.append(rng.integers(0, 256, (h, w), dtype=np.uint8))                                    # noise: queue rounds, list flushes
    low = (128 + 6 * np.sin(xx / 3.1) * np.cos(yy / 2.7) + rng.integers(-5, 6, (h, w))).astype(np.uint8)
    out.append(low)                                                                             # low contrast: corners only at minThFAST
    half = out[0].copy(); half[:, w // 2:] = low[:, w // 2:]; half[h // 2:, : w // 3] = 40
    out.append(half)                                                                            # mixed: served, retried and empty cells side by side
    for period in (5, 35):
        cb = (((xx // period) + (yy // period)) & 1).astype(np.uint8) * 200 + 20
        out.append(cb)                                                                          # corners on a lattice that walks over cell / tile borders
    return np.stack(out)


def _extract_device(ex, frames, cap):
    """one device-buffer call over the whole batch (the host-buffer call pipelines it in chunks: stage access sees the last chunk only)"""
    import torch
    import orbx
    n, h, w = frames.shape
    step = (w + 15) & ~15                                               # device frames: 16-byte aligned rows and frame stride
    fstride = (step * h + 15) & ~15
    g = torch.zeros(n * fstride, dtype=torch.uint8, device="cuda")
    for f in range(n):
        g[f * fstride:f * fstride + step * h].view(h, step)[:, :w] = torch.from_numpy(frames[f]).cuda()
    kps = torch.zeros((n, cap, orbx.KP_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    desc = torch.zeros((n, cap, 32), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    ex._last_w, ex._last_h = w, h
    ex.extract_batch_device(g.data_ptr(), n, w, h, step, fstride, kps.data_ptr(), desc.data_ptr(), cap, cnt.data_ptr())
    ex.sync()
    return kps.cpu().numpy().view(orbx.KP_DTYPE).reshape(n, cap), desc.cpu().numpy(), cnt.cpu().numpy()


def _cand_sets(ex, nframes):
    res = []
    for f in range(nframes):
        res.append([sorted(map(tuple, ex.candidates(l, frame=f).tolist())) for l in range(8)])
    return res


@pytest.mark.parametrize("w,h", [(1280, 720), (640, 480), (355, 291), (1027, 771), (160, 120), (2049, 400)])
def test_dense_and_cell_formulations_agree_with_each_other_and_the_oracle(built, oracle, w, h):
    import orbx
    frames = _frames(oracle, w, h, w * 3 + h)
    n = len(frames)
    ex = orbx.ORBextractor(max_width=w, max_height=h, max_batch=n, max_keypoints=4096, cand_divisor=2)   # the noise frame: ~10 % of its pixels are keypoints
    try:
        outs, cands = {}, {}
        for mode in (0, 2, 3):                                       # warp per cell; dense with the NMS as a second kernel; dense with the NMS inside the tile kernel
            ex.set_fast_dense(mode)
            outs[mode] = _extract_device(ex, frames, 4096)
            cands[mode] = _cand_sets(ex, n)
        for f in range(n):
            for l in range(8):
                assert cands[0][f][l] == cands[2][f][l], ("FAST candidate sets differ between the formulations", f, l, len(cands[0][f][l]), len(cands[2][f][l]))
                assert cands[0][f][l] == cands[3][f][l], ("FAST candidate sets differ (NMS inside the tile kernel)", f, l, len(cands[0][f][l]), len(cands[3][f][l]))
        k0, d0, c0 = outs[0]
        k2, d2, c2 = outs[2]
        assert np.array_equal(c0, c2)
        for f in range(n):
            assert np.array_equal(k0[f, :c0[f]].view(np.uint8), k2[f, :c2[f]].view(np.uint8)) and np.array_equal(d0[f, :c0[f]], d2[f, :c2[f]]), f
        orc = oracle.COracle()
        for f in range(n):
            ref = orc.extract(frames[f], trace=True)
            for l in range(8):
                r = ref["cands"][l]
                assert cands[2][f][l] == sorted(zip(r["x"].tolist(), r["y"].tolist(), r["score"].tolist())), ("dense FAST vs oracle", f, l)
            assert c2[f] == len(ref["kps"]) and np.array_equal(k2[f, :c2[f]].view(np.uint8), ref["kps"].view(np.uint8)) and np.array_equal(d2[f, :c2[f]], ref["desc"]), f
    finally:
        ex.close()


def test_dense_with_other_thresholds_and_single_frame_calls(built, oracle):
    """iniThFAST / minThFAST other than 20 / 7 (incl. equal ones and 0), and the dense path forced on single-frame calls"""
    import orbx
    w, h = 640, 480
    g = _frames(oracle, w, h, 11)[5]
    for ini, mn in ((20, 7), (40, 40), (12, 3), (7, 20), (1, 0)):
        ex = orbx.ORBextractor(max_width=w, max_height=h, max_batch=8, iniThFAST=ini, minThFAST=mn, max_keypoints=4096, cand_divisor=2)   # thresholds near 0: most pixels are corners
        try:
            ref = oracle.COracle(iniThFAST=ini, minThFAST=mn).extract(g, trace=True)
            for mode in (0, 2, 3):
                ex.set_fast_dense(mode)
                kps, desc = ex(g, cap=4096)
                for l in range(8):
                    r = ref["cands"][l]
                    assert sorted(map(tuple, ex.candidates(l).tolist())) == sorted(zip(r["x"].tolist(), r["y"].tolist(), r["score"].tolist())), (ini, mn, mode, l)
                assert np.array_equal(kps.view(np.uint8), ref["kps"].view(np.uint8)) and np.array_equal(desc, ref["desc"]), (ini, mn, mode)
        finally:
            ex.close()
