#!/usr/bin/env python3
"""Randomised parity sweep on the GPU box: random frame sizes / nfeatures / scale factors / thresholds / contents, the CUDA path (both FAST
formulations) against the C oracle and, where built, the reference's compiled ORBextractor.cpp; then the three matcher engines (POPC, mma.sync,
tcgen05) on random problem sizes against BFMatcher's restatement.  usage: fuzz_parity.py [cases] [seed]"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dynamic-visual-slam_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orbx, c_oracle as co
try:
    import ref_oracle as ro
    have_ref = ro.available()
except Exception:
    have_ref = False
ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for case in range(ncases):
    while True:
        w, h = int(rng.integers(100, 1500)), int(rng.integers(100, 1100))
        nf = int(rng.choice([200, 500, 1000, 1500, 2500]))
        sf = float(rng.choice([1.1, 1.2, 1.25, 1.3, 1.44]))
        nl = int(rng.choice([4, 6, 8]))
        ini, mn = [(20, 7), (20, 7), (30, 10), (12, 5), (40, 40)][int(rng.integers(0, 5))]
        orc = co.COracle(nfeatures=nf, scaleFactor=sf, nlevels=nl, iniThFAST=ini, minThFAST=mn)
        if orc.geometry_status(w, h) == 0:
            break
    kind = int(rng.integers(0, 4))
    g = co.synth_gray(int(rng.integers(1, 1 << 30)), int(rng.integers(0, 1000)), w, h)
    if kind == 1:
        g = (g.astype(np.int32) // 4 + 100).astype(np.uint8)                                   # low contrast: retry cells
    elif kind == 2:
        g[:, : w // 2] = rng.integers(0, 256, (h, w // 2), dtype=np.uint8)                      # half noise
    elif kind == 3:
        yy, xx = np.mgrid[0:h, 0:w]; p = int(rng.integers(3, 40))
        g = ((((xx // p) + (yy // p)) & 1) * int(rng.integers(30, 220)) + 20).astype(np.uint8)
    ref = orc.extract(g, trace=True)
    ex = orbx.ORBextractor(nfeatures=nf, scaleFactor=sf, nlevels=nl, iniThFAST=ini, minThFAST=mn, max_width=w, max_height=h, max_batch=8,
                           max_keypoints=max(4096, 2 * nf), cand_divisor=2)
    try:
        for mode in (0, 2):
            ex.set_fast_dense(mode)
            kps, desc = ex(g, cap=max(4096, 2 * nf))
            ok = len(kps) == len(ref["kps"]) and np.array_equal(kps.view(np.uint8), ref["kps"].view(np.uint8)) and np.array_equal(desc, ref["desc"])
            for l in range(nl):
                r = ref["cands"][l]
                ok = ok and sorted(map(tuple, ex.candidates(l).tolist())) == sorted(zip(r["x"].tolist(), r["y"].tolist(), r["score"].tolist()))
            if not ok:
                bad += 1
                print("MISMATCH case", case, dict(w=w, h=h, nf=nf, sf=sf, nl=nl, ini=ini, mn=mn, kind=kind, mode=mode), len(kps), len(ref["kps"]), flush=True)
        if have_ref:
            rx = ro.RefExtractor(nfeatures=nf, scaleFactor=sf, nlevels=nl, iniThFAST=ini, minThFAST=mn)
            r2 = rx.extract(g)
            if not (np.array_equal(r2["kps"].view(np.uint8), ref["kps"].view(np.uint8)) and np.array_equal(r2["desc"], ref["desc"])):
                bad += 1
                print("ORACLE != COMPILED REFERENCE case", case, dict(w=w, h=h, nf=nf, sf=sf, nl=nl, ini=ini, mn=mn, kind=kind), flush=True)
    finally:
        ex.close()
# ---- matcher engines (POPC, mma.sync, tcgen05) on random problem sizes: k = 1, k = 2, threshold ----
em = orbx.ORBextractor(max_width=320, max_height=240)
for case in range(ncases):
    nq, nt = int(rng.integers(1, 2500)), int(rng.integers(1, 6000))
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    if nt > 3:                                                       # near neighbours and exact duplicates (ties -> lowest index)
        for i in range(0, nq, 3):
            q[i] = t[int(rng.integers(0, nt))]; q[i, int(rng.integers(0, 32))] ^= np.uint8(1 << int(rng.integers(0, 8)))
        t[nt - 1] = t[0]
    want1, want2 = co.match(q, t), co.knn2(q, t)
    for eng in (0, 2, 3):
        em.set_match_mma(eng)
        m = em.match(q, t, k=1)
        k2 = em.match(q, t, k=2).reshape(-1, 2)
        good = em.match(q, t, k=1, max_dist=60.0)
        ok = (np.array_equal(m.view(np.uint8), want1.view(np.uint8)) and np.array_equal(k2["trainIdx"], want2["trainIdx"]) and np.array_equal(k2["distance"], want2["distance"])
              and np.array_equal(good.view(np.uint8), want1[want1["distance"] < 60.0].view(np.uint8)))
        if not ok:
            bad += 1
            print("MATCH MISMATCH case", case, dict(nq=nq, nt=nt, engine=eng), flush=True)
em.close()
print("fuzz: %d cases, %d mismatches, compiled reference %s" % (ncases, bad, "checked" if have_ref else "absent"))
sys.exit(1 if bad else 0)
