#!/usr/bin/env python3
"""Workload for the ncu captures of the kernels added late in round 2: the tensor-core association (k_assoc_mma), the dense FAST formulation
(k_fast_dense + k_fast_nms + k_fast_retry_list + k_fast_cells in retry mode, and the variant with the NMS inside the tile kernel) and the
device RANSAC (k_fmat_hypotheses).  Run under
    ncu --set full --clock-control none -k regex:'k_assoc_mma|k_fast_dense|k_fast_nms|k_fast_retry_list|k_fast_cells|k_fmat_hypotheses' ...
"""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-visual-slam_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orbx
dev = torch.device("cuda", 0)
W, H, B = 1280, 720, 32
ex = orbx.ORBextractor(max_width=W, max_height=H, max_batch=B, max_keypoints=1280)
L, h = ex.L, ex.handle
NQ, ROWS = 2048, 1 << 20
db = orbx.LandmarkDB(ex, ROWS)
rows = torch.empty((ROWS, 32), dtype=torch.uint8, device=dev)
ex._check(L.orbx_synth_descriptors_device(h, 1234, 0, ROWS, rows.data_ptr())); db.append_device(rows.data_ptr(), ROWS)
q = rows[::512][:NQ].contiguous().clone(); q[:, 3] ^= 0x5A
rng = np.random.default_rng(1)
pos = rng.uniform(-3, 3, (ROWS, 3)).astype(np.float32) + np.array([0, 0, 6], np.float32)
db.set_positions(pos)
pose = orbx.LandmarkDB.pose(np.eye(3), np.zeros(3), 600.0, 600.0, 640.0, 360.0)
qpx = rng.uniform(0, 1280, (NQ, 2)).astype(np.float32)
for _ in range(2): db.associate(q.cpu().numpy(), qpx, pose)                      # k_assoc_mma (2048 x 1M >= 8 M pairs)
gray = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
ex._check(L.orbx_synth_gray_device(h, 20261018, 0, B, W, H, gray.data_ptr(), W, W * H))
kps = torch.empty((B, 1280, 28), dtype=torch.uint8, device=dev); desc = torch.empty((B, 1280, 32), dtype=torch.uint8, device=dev); cnt = torch.zeros(B, dtype=torch.int32, device=dev)
for mode in (1, 3):                                                              # NMS as a second kernel / inside the tile kernel
    ex.set_fast_dense(mode)
    for _ in range(2): ex.extract_batch_device(gray.data_ptr(), B, W, H, W, W * H, kps.data_ptr(), desc.data_ptr(), 1280, cnt.data_ptr())
    ex.sync()
ex.set_fast_dense(0)
p1 = rng.uniform(0, 1280, (800, 2)).astype(np.float32); p2 = p1 + rng.normal(0, 1, (800, 2)).astype(np.float32)
for _ in range(2): ex.fmat_ransac(p1, p2, iters=1000, threshold=2.0, seed=3)
print("extra2 workload done")
