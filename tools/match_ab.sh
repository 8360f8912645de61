#!/bin/bash
# A/B of the two matchers (ORBX_OPT_MATCH_MMA) on the frame-matching step and the 2048 x 1M association (run on the GPU box)
python - <<'PY'
import sys, ctypes as ct, numpy as np, torch
sys.path.insert(0, "dynamic-visual-slam_b200/python")
import orbx
ex = orbx.ORBextractor(max_width=1280, max_height=720, max_batch=1, max_keypoints=1280)
L, h = ex.L, ex.handle
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ex.stream, device=dev)
NQ, ROWS = 2048, 1 << 20
db = orbx.LandmarkDB(ex, ROWS)
rows = torch.empty((ROWS, 32), dtype=torch.uint8, device=dev)
ex._check(L.orbx_synth_descriptors_device(h, 1234, 0, ROWS, rows.data_ptr())); db.append_device(rows.data_ptr(), ROWS)
q = rows[::512][:NQ].contiguous().clone(); q[:, 3] ^= 0x5A
out = torch.empty((NQ, 4), dtype=torch.int32, device=dev)
res = {}
for mma in (0, 1, 0, 1):
    ex.set_match_mma(bool(mma))
    for _ in range(3): db.query_top2_device(q.data_ptr(), NQ, out.data_ptr())
    ex.sync(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): db.query_top2_device(q.data_ptr(), NQ, out.data_ptr())
    e1.record(stream); ex.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res.setdefault(mma, out.clone())
    print("association 2048 x 1M  mma=%d  %.3f ms  %.0f G pairs/s  same=%s" % (mma, ms, NQ * ROWS / ms / 1e6, bool((res[0] == out).all()) if 0 in res else None))
PY
