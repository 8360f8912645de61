#!/usr/bin/env python3
"""gen_golden.py — writes the golden vectors under tests/golden/ (TEST INFRASTRUCTURE).

Runs in the build container only: it needs /root/reference (compiled unmodified into oracle/_ref, see oracle/Makefile) and
python cv2; every extraction vector is asserted identical between the two before it is written.  It drives python cv2 (4.13.0) — the same OpenCV primitives the
reference calls (cv::resize INTER_LINEAR, cv::FAST, cv::GaussianBlur, cv::fastAtan2,
cv::BFMatcher(NORM_HAMMING)) — through oracle/cv_oracle.py, the literal restatement of the reference's
ORB_SLAM3::ORBextractor (reference dynamic_visual_slam/src/ORBextractor.cpp).  The GPU box has neither
/root/reference nor a guarantee of cv2, so the outputs are committed as small .npz fixtures and the
dependency-free C oracle (orb_oracle.c) and the CUDA path are checked against them.

    python oracle/gen_golden.py            # regenerates tests/golden/*.npz

Inputs are either the seeded integer-only synthetic generator (orc_synth_gray, identical bytes on host
and device) or the reference's own test fixture: a black 640x480 image with three filled white circles
(reference test/test_dbow2_integration.cpp:14-17), drawn here with cv2.circle exactly as the test does.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import c_oracle as co          # noqa: E402  (only for the synthetic input generator)
import cv_oracle as cvo        # noqa: E402
import ref_oracle as ro        # noqa: E402  (the reference's own ORBextractor.cpp, compiled unmodified: oracle/_ref)
import cv2                     # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def kp_struct(tab):
    k = np.zeros(len(tab), co.KP_DTYPE)
    for i, name in enumerate(["x", "y", "size", "angle", "response"]):
        k[name] = tab[:, i].astype(np.float32)
    k["octave"] = tab[:, 5].astype(np.int32)
    k["class_id"] = tab[:, 6].astype(np.int32)
    return k


def circles_image():
    img = np.zeros((480, 640), np.uint8)
    cv2.circle(img, (100, 100), 50, 255, -1)
    cv2.circle(img, (300, 200), 30, 255, -1)
    cv2.circle(img, (500, 300), 40, 255, -1)
    return img


def extract_case(name, img, **meta):
    ex = cvo.ORBextractorCV()
    tr = {}
    tab, desc = ex(img, trace=tr)
    # every vector written below is ALSO what the compiled reference (oracle/_ref) returns for this input: the two independent
    # routes — real cv2 primitives under a Python restatement, and the reference's own C++ over restated primitives — must agree
    r = ro.RefExtractor().extract(img)
    assert np.array_equal(r["kps"].view(np.uint8), kp_struct(tab).view(np.uint8)) and np.array_equal(r["desc"], desc), name
    cands = [np.array(sorted(c), np.int32).reshape(-1, 3) for c in tr["cands"]]
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        image_crc=crc(img), width=img.shape[1], height=img.shape[0],
        kps=kp_struct(tab), desc=desc,
        level_counts=np.array([len(s) for s in tr["selected"]], np.int32),
        cand_counts=np.array([len(c) for c in cands], np.int32),
        cand_crc=np.array([crc(c) for c in cands], np.uint32),          # candidates sorted by (x, y, score)
        pyr_crc=np.array([crc(p) for p in tr["pyramid"]], np.uint32),
        blur_crc=np.array([crc(b) for b in tr["blurred"]], np.uint32),
        level_w=np.array([p.shape[1] for p in tr["pyramid"]], np.int32),
        level_h=np.array([p.shape[0] for p in tr["pyramid"]], np.int32),
        **meta)
    print(name, len(tab), "keypoints")
    return tab, desc


def main():
    os.makedirs(OUT, exist_ok=True)
    cv2.setNumThreads(1)
    # ---- extraction: synthetic frames (seed, frame, size) and the reference's own test image ----
    _, d0 = extract_case("extract_synth_320x240_s1_f0", co.synth_gray(1, 0, 320, 240), seed=1, frame=0)
    _, d1 = extract_case("extract_synth_320x240_s1_f1", co.synth_gray(1, 1, 320, 240), seed=1, frame=1)
    extract_case("extract_synth_417x301_s4_f0", co.synth_gray(4, 0, 417, 301), seed=4, frame=0)
    extract_case("extract_circles_640x480", circles_image(), image=circles_image())   # compresses to a few KB
    # ---- the benchmark stream (bench.py SEED, 1280x720): per-frame checksums of the compiled reference's output ----
    seed, frames = 20261018, list(range(32)) + [127, 128, 255, 1023, 4095]
    ref = ro.RefExtractor()
    res = [ref.extract(co.synth_gray(seed, f, 1280, 720)) for f in frames]
    np.savez_compressed(os.path.join(OUT, "ref_stream_1280x720.npz"), seed=seed, frames=np.array(frames, np.int32),
                        count=np.array([len(r["kps"]) for r in res], np.int32),
                        kps_crc=np.array([crc(r["kps"]) for r in res], np.uint32), desc_crc=np.array([crc(r["desc"]) for r in res], np.uint32))
    # ---- profile C: cv::ORB itself (cv2.ORB_create) on the reference's three-circle test image (100 features, as the gtest) and on
    #      synthetic frames (1000 features, BASELINE configs[0]); rows sorted by (octave, response descending, y, x) ----
    import cvorb_oracle as cvc
    for name, img, nf in (("cvorb_circles_640x480_n100", circles_image(), 100), ("cvorb_synth_640x480_f0", co.synth_gray(20261018, 0, 640, 480), 1000),
                          ("cvorb_synth_640x480_f1", co.synth_gray(20261018, 1, 640, 480), 1000)):
        kp, d = cv2.ORB_create(nf).detectAndCompute(img, None)
        t = np.zeros(len(kp), co.KP_DTYPE)
        for i, k in enumerate(kp):
            t[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
        order = np.lexsort((t["x"], t["y"], -t["response"].astype(np.float64), t["octave"]))
        ok, od = cvc.CvOrb(nfeatures=nf).extract(img)
        assert np.array_equal(t[order].view(np.uint8), ok.view(np.uint8)) and np.array_equal(d[order], od), name      # the restatement IS cv2's output
        np.savez_compressed(os.path.join(OUT, name + ".npz"), image_crc=crc(img), nfeatures=nf, kps=t[order], desc=d[order],
                            **({"image": img} if "circles" in name else {"seed": 20261018, "frame": int(name[-1]), "width": 640, "height": 480}))
        print(name, len(kp), "keypoints")
    # ---- geometric validation: cv::findFundamentalMat(FM_RANSAC, 2.0, 0.99) on the matches of a synthetic frame pair (frontend.cpp:1134-1154):
    #      the returned model and ITS inlier mask; the scoring restatement / kernel must reproduce that mask for that model ----
    orc = co.COracle()
    ra, rb = orc.extract(co.synth_gray(20261018, 0, 640, 480)), orc.extract(co.synth_gray(20261018, 1, 640, 480))
    mm = co.match(rb["desc"], ra["desc"])
    mm = mm[mm["distance"] < 50.0]
    p_prev = np.stack([ra["kps"]["x"][mm["trainIdx"]], ra["kps"]["y"][mm["trainIdx"]]], 1).astype(np.float32)
    p_curr = np.stack([rb["kps"]["x"][mm["queryIdx"]], rb["kps"]["y"][mm["queryIdx"]]], 1).astype(np.float32)
    rngf = np.random.default_rng(99)
    bad = rngf.choice(len(p_curr), len(p_curr) // 5, replace=False)                  # 20 % gross outliers
    p_curr[bad] += rngf.uniform(-40, 40, (len(bad), 2)).astype(np.float32)
    Fm, msk = cv2.findFundamentalMat(p_prev, p_curr, cv2.FM_RANSAC, 2.0, 0.99)
    n_in, m2 = co.fmat_inliers(p_prev, p_curr, Fm[:3], 2.0)
    assert np.array_equal(m2, msk.ravel()), "scoring restatement differs from cv2's mask for cv2's own model"
    np.savez_compressed(os.path.join(OUT, "fmat_ransac.npz"), pts_prev=p_prev, pts_curr=p_curr, F=Fm[:3].astype(np.float64), mask=msk.ravel().astype(np.uint8))
    print("fmat_ransac", len(p_prev), "pairs,", int(msk.sum()), "inliers")
    # ---- primitives ----
    rng = np.random.default_rng(12345)
    noise = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    smooth = cv2.GaussianBlur(rng.integers(0, 256, (120, 160), dtype=np.uint8), (9, 9), 3)
    prim = {}
    prim["noise"] = noise
    prim["smooth"] = smooth
    prim["resize_noise_109x81"] = cv2.resize(noise, (109, 81), interpolation=cv2.INTER_LINEAR)
    prim["resize_smooth_133x100"] = cv2.resize(smooth, (133, 100), interpolation=cv2.INTER_LINEAR)
    prim["blur_noise"] = cv2.GaussianBlur(noise.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    prim["blur_smooth"] = cv2.GaussianBlur(smooth.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    for th in (20, 7):
        for nm, im in (("noise", noise), ("smooth", smooth)):
            kp = cv2.FastFeatureDetector_create(th, True).detect(im)
            prim["fast%d_%s" % (th, nm)] = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kp], np.int32).reshape(-1, 3)
    ys = rng.integers(-3000000, 3000000, 4096).astype(np.float32)
    xs = rng.integers(-3000000, 3000000, 4096).astype(np.float32)
    ys[:8] = [0, 0, 1, -1, 5, -5, 0, 7]
    xs[:8] = [0, 1, 0, 0, 5, 5, -3, -7]
    prim["atan2_y"], prim["atan2_x"] = ys, xs
    prim["atan2"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    np.savez_compressed(os.path.join(OUT, "primitives.npz"), **prim)
    # ---- matching: BFMatcher(NORM_HAMMING).match / knnMatch(k=2), incl. lowest-index tie-breaks ----
    q = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (96, 32), dtype=np.uint8)
    t[10] = t[3]; t[50] = t[3]; q[5] = t[3]                         # exact duplicates: ties at distance 0
    q[6] = t[20]; q[6, 0] ^= 1; t[70] = t[20]; t[70, 1] ^= 2         # two rows at distance 1 and 2... and a tie below
    t[80] = t[20]; t[80, 5] ^= 4                                     # distance-2 tie between rows 70 and 80
    m = cvo.bf_match(q, t)
    k2 = cvo.bf_knn2(q, t)
    mf = cvo.bf_match(d1, d0)
    np.savez_compressed(
        os.path.join(OUT, "match.npz"), q=q, t=t,
        match=np.array(m, np.float32), knn2=np.array(k2, np.float32).reshape(len(q), 2, 3),
        frame_match=np.array(mf, np.float32))                        # extract_synth_320x240 f1 vs f0
    print("golden vectors written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
