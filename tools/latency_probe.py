import sys, time, ctypes as ct, numpy as np, torch
sys.path.insert(0, 'dynamic-visual-slam_b200/python')
import orbx
W, H, CAP = 1280, 720, 1280
dev = torch.device("cuda", 0)
ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=1, max_keypoints=CAP)
L, hnd = ex.L, ex.handle
if 'describe-all' in sys.argv: ex.set_filter_first(False)          # ORBX_OPT_FILTER_FIRST = 0 (the reference's order)
gray = torch.empty((1, H, W), dtype=torch.uint8, device=dev); depth = torch.empty((1, H, W), dtype=torch.int16, device=dev)
ex._check(L.orbx_synth_gray_device(hnd, 7, 0, 1, W, H, gray.data_ptr(), W, W * H))
ex._check(L.orbx_synth_depth_device(hnd, 7, 0, 1, W, H, depth.data_ptr(), 2 * W, 2 * W * H))
kps = torch.empty((1, CAP, 28), dtype=torch.uint8, device=dev); desc = torch.empty((1, CAP, 32), dtype=torch.uint8, device=dev)
counts = torch.zeros(1, dtype=torch.int32, device=dev); m = torch.empty((1, CAP, 16), dtype=torch.uint8, device=dev); mc = torch.zeros(1, dtype=torch.int32, device=dev)
def step():
    ex._check(L.orbx_track_batch_device(hnd, gray.data_ptr(), 1, W, H, W, W * H, depth.data_ptr(), 2 * W, 2 * W * H, kps.data_ptr(), desc.data_ptr(), CAP, counts.data_ptr(), m.data_ptr(), mc.data_ptr(), ct.c_float(50.0)))
for _ in range(20): step()
ex.sync()
# CPU time to enqueue one frame (GPU idle at start of each: sync between)
cpu = []
tot = []
for _ in range(100):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); ex.sync(); t2 = time.perf_counter()
    cpu.append((t1 - t0) * 1e6); tot.append((t2 - t0) * 1e6)
print("enqueue CPU time p50 %.1f us, call-to-sync p50 %.1f us" % (np.percentile(cpu, 50), np.percentile(tot, 50)))
t0 = time.perf_counter()
for _ in range(200): step()
ex.sync()
print("back-to-back per frame %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))

# the blocking host call (pinned host buffers in, host arrays out): what bench.py reports as `latency`
orc_frames = gray.cpu().numpy(); orc_depth = depth.cpu().numpy().view(np.uint16)
ex.track_reset()
for _ in range(20): ex.track_batch(orc_frames, orc_depth, cap=CAP)
ts = []
for _ in range(300):
    t0 = time.perf_counter(); ex.track_batch(orc_frames, orc_depth, cap=CAP); ts.append((time.perf_counter() - t0) * 1e6)
print("host call, batch 1: p50 %.1f us, p95 %.1f us" % (np.percentile(ts, 50), np.percentile(ts, 95)))
