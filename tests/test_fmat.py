"""Geometric validation, hypothesis scoring (reference frontend.cpp:1134-1154, :625-645:
cv::findFundamentalMat(prev_pts, curr_pts, mask, FM_RANSAC, 2.0, 0.99)).  OpenCV's RANSAC draws its samples from its own RNG, so the
model it returns is not reproducible elsewhere; what is pinned is the SCORING: for a given model the inlier mask equals OpenCV's
(tests/golden/fmat_ransac.npz = cv2's returned model and mask on the matches of a synthetic frame pair with 20 % gross outliers)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fmat_ransac.npz")


def _hypotheses(F, k, seed=3):
    rng = np.random.default_rng(seed)
    H = np.repeat(F.reshape(1, 9), k, 0) * (1.0 + rng.normal(0, 0.02, (k, 9)))
    H[k // 3] = F.reshape(9)                                        # the true model somewhere in the middle
    H[k // 3 + 5] = F.reshape(9)                                    # and again later: ties resolve to the lowest index
    return H


def test_oracle_scoring_equals_cv2_golden(oracle):
    g = np.load(GOLD)
    n, mask = oracle.fmat_inliers(g["pts_prev"], g["pts_curr"], g["F"], 2.0)
    assert n == int(g["mask"].sum()) and np.array_equal(mask, g["mask"])


def test_oracle_scoring_vs_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for trial in range(4):
        X = rng.uniform(-2, 2, (400, 3)) + [0, 0, 6]
        K = np.array([[600.0, 0, 320], [0, 600, 240], [0, 0, 1]])
        R, _ = cv2.Rodrigues(rng.normal(0, 0.05, 3))
        t = rng.normal(0, 0.3, 3)
        a = (K @ X.T).T
        b = (K @ (R @ X.T + t[:, None])).T
        p1 = (a[:, :2] / a[:, 2:]).astype(np.float32) + rng.normal(0, 0.3, (400, 2)).astype(np.float32)
        p2 = (b[:, :2] / b[:, 2:]).astype(np.float32) + rng.normal(0, 0.3, (400, 2)).astype(np.float32)
        p2[:60] += rng.uniform(-50, 50, (60, 2)).astype(np.float32)
        F, m = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 2.0, 0.99)
        n, mask = oracle.fmat_inliers(p1, p2, F[:3], 2.0)
        assert np.array_equal(mask, m.ravel()), trial


@pytest.mark.gpu
def test_gpu_scoring_bit_exact(built, oracle):
    import orbx
    g = np.load(GOLD)
    p1, p2, F = g["pts_prev"], g["pts_curr"], g["F"]
    H = _hypotheses(F, 1000)
    ex = orbx.ORBextractor(max_width=640, max_height=480)
    try:
        counts, best, mask = ex.fmat_score(p1, p2, H, 2.0)
        want = [oracle.fmat_inliers(p1, p2, h, 2.0) for h in H]
        assert counts.tolist() == [w[0] for w in want]
        top = max(w[0] for w in want)
        assert best == [w[0] for w in want].index(top)
        assert np.array_equal(mask, want[best][1])
        assert counts[1000 // 3] == int(g["mask"].sum())            # cv2's own model scores cv2's own inlier count
        c1, b1, m1 = ex.fmat_score(p1, p2, F.reshape(1, 9), 2.0)
        assert b1 == 0 and np.array_equal(m1, g["mask"])            # ... and its mask, bit for bit
        c0, b0, m0 = ex.fmat_score(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), F.reshape(1, 9), 2.0)
        assert c0.tolist() == [0] and len(m0) == 0
    finally:
        ex.close()
