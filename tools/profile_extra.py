#!/usr/bin/env python3
"""Workload for the ncu captures of the kernels outside the per-step chain: the landmark-database query
(k_match_mma / k_match_partial at 2048 x 1M), the reprojection-gated association (k_assoc_partial), the whole-level Gaussian (k_blur7,
ORBX_OPT_FUSED_BLUR = 0), the backend culling (k_cull), profile C's kernels and the F-matrix scoring.  Run under
    ncu --set full --clock-control none -k regex:'k_match_mma|k_match_partial|k_assoc_partial|k_blur7|k_cull|k_cfast|k_cretain|k_cblur|k_fmat|k_resize_exact|k_describe_c' ...
"""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-visual-slam_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orbx, c_oracle as co
dev = torch.device("cuda", 0)
W, H, B = 1280, 720, 32
ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=B, max_keypoints=1280)
L, h = ex.L, ex.handle
# landmark database: 2048 x 1M top-2 with both engines, then the reprojection-gated association
NQ, ROWS = 2048, 1 << 20
db = orbx.LandmarkDB(ex, ROWS)
rows = torch.empty((ROWS, 32), dtype=torch.uint8, device=dev)
ex._check(L.orbx_synth_descriptors_device(h, 1234, 0, ROWS, rows.data_ptr())); db.append_device(rows.data_ptr(), ROWS)
q = rows[::512][:NQ].contiguous().clone(); q[:, 3] ^= 0x5A
out = torch.empty((NQ, 4), dtype=torch.int32, device=dev)
for mma in (1, 0):
    ex.set_match_mma(mma)
    for _ in range(2): db.query_top2_device(q.data_ptr(), NQ, out.data_ptr())
ex.set_match_mma(1); ex.sync()
rng = np.random.default_rng(1)
pos = rng.uniform(-3, 3, (ROWS, 3)).astype(np.float32) + np.array([0, 0, 6], np.float32)
db.set_positions(pos)
pose = orbx.LandmarkDB.pose(np.eye(3), np.zeros(3), 600.0, 600.0, 640.0, 360.0)
qpx = rng.uniform(0, 1280, (NQ, 2)).astype(np.float32)
for _ in range(2): db.associate(q.cpu().numpy(), qpx, pose)
# whole-level Gaussian + describe on the blurred pyramid
gray = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
ex._check(L.orbx_synth_gray_device(h, 20261018, 0, B, W, H, gray.data_ptr(), W, W * H))
kps = torch.empty((B, 1280, 28), dtype=torch.uint8, device=dev); desc = torch.empty((B, 1280, 32), dtype=torch.uint8, device=dev); cnt = torch.zeros(B, dtype=torch.int32, device=dev)
ex.set_fused_blur(False)
for _ in range(2): ex.extract_batch_device(gray.data_ptr(), B, W, H, W, W * H, kps.data_ptr(), desc.data_ptr(), 1280, cnt.data_ptr())
ex.set_fused_blur(True); ex.sync()
# backend culling on one frame
n = int(cnt[0].item()); k0 = kps[0, :n].cpu().numpy().view(orbx.KP_DTYPE).reshape(-1); d0 = desc[0, :n].cpu().numpy()
for _ in range(2): ex.cull_keyframe(k0, d0, np.arange(0, n, 3, dtype=np.int32))
# F-matrix hypothesis scoring: 1000 x 800
p1 = rng.uniform(0, 1280, (800, 2)).astype(np.float32); p2 = p1 + rng.normal(0, 1, (800, 2)).astype(np.float32)
for _ in range(2): ex.fmat_score(p1, p2, rng.normal(0, 1, (1000, 9)), 2.0)
# profile C on one 1280 x 720 frame
ec = orbx.ORBextractor(max_width=W, max_height=H, profile="cvorb")
g0 = co.synth_gray(20261018, 0, W, H)
for _ in range(2): ec(g0)
print("extra workload done")
