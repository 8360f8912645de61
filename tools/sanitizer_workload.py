import sys, numpy as np
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'dynamic-visual-slam_b200/python')
import c_oracle as co, orbx
w, h = 640, 480
ex = orbx.ORBextractor(max_width=w, max_height=h, max_batch=4)
frames = np.stack([co.synth_gray(5, f, w, h) for f in range(5)])
depths = np.stack([co.synth_depth(5, f, w, h) for f in range(5)])
out = ex.track_batch(frames, depths)
k, d = ex(frames[0], depth=depths[0])
r = co.COracle().extract(frames[0])
fk, fd, _ = co.filter_depth(r["kps"], r["desc"], depths[0])
assert np.array_equal(d, fd)
m = ex.match(d, fd, k=2, ratio=0.75)
g = np.random.default_rng(1).integers(0, 256, (200, 150), dtype=np.uint8)
e2 = orbx.ORBextractor(max_width=150, max_height=200, cand_divisor=1)
k2, d2 = e2(g, cap=8192)
db = orbx.LandmarkDB(ex, 4096)
rows = co.synth_descriptors(1, 0, 3000); db.append(rows); db.set_positions(np.random.default_rng(2).standard_normal((3000, 3)).astype(np.float32) + [0, 0, 3])
a = db.associate(rows[:50], np.full((50, 2), 300, np.float32), orbx.LandmarkDB.pose(np.eye(3), np.zeros(3), 600, 600, 320, 240))
# a batch large enough for the strided descriptor grid (>= 16 frames), filter first and the reference's order, depth + per-frame boxes
e3 = orbx.ORBextractor(max_width=w, max_height=h, max_batch=16)
f16 = np.stack([co.synth_gray(5, f, w, h) for f in range(16)]); d16 = np.stack([co.synth_depth(5, f, w, h) for f in range(16)])
b16 = [co.synth_boxes(5, f, w, h) for f in range(16)]
o1 = e3.track_batch(f16, d16, frame_boxes=b16, drop_class_mask=1)
e3.set_filter_first(False); e3.track_reset()
o0 = e3.track_batch(f16, d16, frame_boxes=b16, drop_class_mask=1)
assert np.array_equal(o1[2], o0[2]) and np.array_equal(o1[4], o0[4]), "counts differ between the two filter orders"
for f in range(16):                                   # rows past a frame's count are not defined: compare the valid ones
    n, nm = int(o1[2][f]), int(o1[4][f])
    assert np.array_equal(o1[0][f, :n].view(np.uint8), o0[0][f, :n].view(np.uint8)) and np.array_equal(o1[1][f, :n], o0[1][f, :n]), ("keypoints / descriptors", f)
    assert np.array_equal(o1[3][f, :nm].view(np.uint8), o0[3][f, :nm].view(np.uint8)), ("matches", f)
print("sanitizer workload ok", len(k), len(k2), out[2].tolist(), int((a["landmark"] >= 0).sum()))
