"""Where does the end-to-end rate go?  Host-buffer batches (two in flight) with and without the depth filter / the matcher.
usage: python tools/e2e_probe.py"""
import ctypes as ct, sys, time
import numpy as np
sys.path.insert(0, "dynamic-visual-slam_b200/python")
import orbx
import torch

W, H, B, CAP, K = 1280, 720, 128, 1280, 30
ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=B, max_keypoints=CAP)
L, hnd = ex.L, ex.handle
dev = torch.device("cuda", 0)
g = torch.empty((B, H, W), dtype=torch.uint8, device=dev); d = torch.empty((B, H, W), dtype=torch.int16, device=dev)
ex._check(L.orbx_synth_gray_device(hnd, 5, 0, B, W, H, g.data_ptr(), W, W * H))
ex._check(L.orbx_synth_depth_device(hnd, 5, 0, B, W, H, d.data_ptr(), 2 * W, 2 * W * H))
ex.sync()
pg, pd = orbx.PinnedArray((B, H, W), np.uint8), orbx.PinnedArray((B, H, W), np.uint16)
ex._check(L.orbx_copy_to_host(hnd, ct.c_void_p(pg.ptr), g.data_ptr(), B * W * H))
ex._check(L.orbx_copy_to_host(hnd, ct.c_void_p(pd.ptr), d.data_ptr(), B * W * H * 2))
outs = [dict(k=orbx.PinnedArray((B, CAP), orbx.KP_DTYPE), d=orbx.PinnedArray((B, CAP, 32), np.uint8), c=orbx.PinnedArray((B,), np.int32),
             m=orbx.PinnedArray((B, CAP), orbx.DM_DTYPE), mc=orbx.PinnedArray((B,), np.int32)) for _ in range(2)]


def run(mode, n):
    prev = None
    for k in range(n):
        o = outs[k & 1]
        t = ct.c_int32()
        depth = ct.c_void_p(pd.ptr) if "depth" in mode else None
        if "track" in mode:
            ex._check(L.orbx_track_batch_submit(hnd, ct.c_void_p(pg.ptr), B, W, H, W, depth, 2 * W, ct.c_void_p(o["k"].ptr), ct.c_void_p(o["d"].ptr), CAP,
                                                ct.c_void_p(o["c"].ptr), ct.c_void_p(o["m"].ptr), ct.c_void_p(o["mc"].ptr), ct.c_float(50.0), ct.byref(t)))
        else:
            ex._check(L.orbx_extract_batch_submit(hnd, ct.c_void_p(pg.ptr), B, W, H, W, depth, 2 * W, ct.c_void_p(o["k"].ptr), ct.c_void_p(o["d"].ptr), CAP,
                                                  ct.c_void_p(o["c"].ptr), ct.byref(t)))
        if prev is not None:
            ex._check(L.orbx_batch_wait(hnd, prev))
        prev = t.value
    ex._check(L.orbx_batch_wait(hnd, prev))


for mode in ("extract", "extract+depth", "track", "track+depth"):
    run(mode, 3)
    t0 = time.perf_counter(); run(mode, K); dt = time.perf_counter() - t0
    print("%-14s %.0f frames/s  %.3f ms/step  (gray H2D alone at that rate: %.1f GB/s)" % (mode, B * K / dt, dt / K * 1e3, B * W * H * K / dt / 1e9))
