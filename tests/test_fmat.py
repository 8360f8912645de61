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


@pytest.mark.gpu
def test_gpu_ransac_model_and_its_inlier_set(built, oracle):
    """orbx_fmat_ransac: 8-point hypotheses generated and scored on the device.  The sample sequence is its own (OpenCV's RNG cannot be
    reproduced), so the checks are (1) the contract of the result — the returned mask is exactly the inlier set OpenCV's error function
    gives for the returned model — and (2) quality: on a synthetic two-view scene with 15 % gross outliers it finds (nearly) all true inliers,
    and as many as cv2.findFundamentalMat where cv2 is importable."""
    import orbx
    rng = np.random.default_rng(5)
    ex = orbx.ORBextractor(max_width=640, max_height=480)
    try:
        for trial in range(3):
            X = rng.uniform(-2, 2, (400, 3)) + [0, 0, 6]
            K = np.array([[600.0, 0, 320], [0, 600, 240], [0, 0, 1]])
            a = rng.normal(0, 0.05, 3)
            th = np.linalg.norm(a); k = a / th
            Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
            t = rng.normal(0, 0.3, 3)
            pa = (K @ X.T).T
            pb = (K @ (R @ X.T + t[:, None])).T
            p1 = (pa[:, :2] / pa[:, 2:]).astype(np.float32) + rng.normal(0, 0.3, (400, 2)).astype(np.float32)
            p2 = (pb[:, :2] / pb[:, 2:]).astype(np.float32) + rng.normal(0, 0.3, (400, 2)).astype(np.float32)
            p2[:60] += rng.uniform(-50, 50, (60, 2)).astype(np.float32)
            F, mask, n_in = ex.fmat_ransac(p1, p2, iters=1000, threshold=2.0, seed=7 + trial)
            n_or, m_or = oracle.fmat_inliers(p1, p2, F, 2.0)
            assert n_in == n_or and np.array_equal(mask, m_or), "the mask must be OpenCV's inlier set of the returned model"
            assert abs(F[2, 2] - 1.0) < 1e-12 and abs(np.linalg.det(F)) < 1e-9 * max(1.0, np.abs(F).max() ** 3)        # scaled like OpenCV's, rank 2
            assert mask[60:].sum() >= 0.9 * 340 and mask[:60].sum() <= 12, (trial, int(mask[60:].sum()), int(mask[:60].sum()))
            try:
                import cv2
                _, mcv = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 2.0, 0.99)
                assert n_in >= 0.95 * int(mcv.sum()), (trial, n_in, int(mcv.sum()))
            except ImportError:
                pass
        F2, mask2, n2 = ex.fmat_ransac(p1, p2, iters=1000, threshold=2.0, seed=9)
        assert np.array_equal(F, F2) and np.array_equal(mask, mask2)                                                    # deterministic in the seed
        with pytest.raises(orbx.OrbxError):
            ex.fmat_ransac(p1[:7], p2[:7])
    finally:
        ex.close()
