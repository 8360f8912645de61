"""Pose-hypothesis scoring of cv::solvePnPRansac (reference frontend.cpp:906-923) and the correspondence loop in front of it (:858-892).
CPU: the oracle's restatement of PnPRansacCallback::computeError pinned against cv2.projectPoints (the arithmetic OpenCV runs per hypothesis).
GPU: orbx_pnp_points / orbx_pnp_score against the oracle, bit for bit."""
import numpy as np
import pytest


def _scene(rng, n, out_frac=0.2):
    K = (615.3, 615.9, 640.2, 360.4)
    X = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.0, 1.0, n), rng.uniform(0.4, 3.0, n)], 1).astype(np.float32)
    a = rng.normal(0, 0.04, 3); th = np.linalg.norm(a); k = a / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    t = rng.normal(0, 0.05, 3)
    pc = X.astype(np.float64) @ R.T + t
    uv = np.stack([K[0] * pc[:, 0] / pc[:, 2] + K[2], K[1] * pc[:, 1] / pc[:, 2] + K[3]], 1)
    uv += rng.normal(0, 1.5, uv.shape)
    bad = rng.random(n) < out_frac
    uv[bad] += rng.uniform(-60, 60, (int(bad.sum()), 2))
    return X, uv.astype(np.float32), a, R, t, K


def test_oracle_error_equals_opencv_projectpoints(oracle):
    """the inlier mask of a pose = (squared float distance to cv2.projectPoints' float output) <= 16, for poses near and far from the truth,
    incl. points behind the camera and exactly on the threshold's neighbourhood"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for trial in range(20):
        X, uv, a, R, t, K = _scene(rng, 700)
        rvec = (a + rng.normal(0, 0.02 * (trial % 4), 3)).reshape(3, 1)
        tvec = (t + rng.normal(0, 0.03 * (trial % 3), 3)).reshape(3, 1)
        if trial % 5 == 4:
            X[:20, 2] *= -1                                                      # behind the camera: projectPoints still divides
        Rm = cv2.Rodrigues(rvec)[0]
        Kmat = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float64)
        proj = cv2.projectPoints(X.reshape(-1, 1, 3), rvec, tvec, Kmat, None)[0].reshape(-1, 2).astype(np.float32)
        d = uv - proj                                                            # float32
        err = (d[:, 0] * d[:, 0]).astype(np.float32)
        err = (err + (d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32)
        want = err <= np.float32(16.0)
        n, mask = oracle.pnp_inliers(X, uv, Rm, tvec, *K, thresh=4.0)
        assert np.array_equal(mask.astype(bool), want), (trial, int((mask.astype(bool) != want).sum()))
        assert n == int(want.sum())


def test_oracle_correspondences_follow_the_reference_loop(oracle):
    """match order kept; rounded pixel lookup; depth gate (0.3, 3.0]; out-of-image keypoints skipped"""
    rng = np.random.default_rng(5)
    w, h = 640, 480
    depth = rng.integers(0, 4000, (h, w), dtype=np.uint16)
    kp = np.zeros(300, oracle.KP_DTYPE); kc = np.zeros(280, oracle.KP_DTYPE)
    kp["x"] = rng.uniform(-2, w + 2, 300).astype(np.float32); kp["y"] = rng.uniform(-2, h + 2, 300).astype(np.float32)
    kp["x"][:5] = [0.5, 1.5, 2.5, 639.5, 639.4999]                            # std::round: half away from zero; 639.5 -> 640 is outside
    kc["x"] = rng.uniform(0, w, 280).astype(np.float32); kc["y"] = rng.uniform(0, h, 280).astype(np.float32)
    m = np.zeros(250, oracle.DM_DTYPE)
    m["queryIdx"] = rng.integers(0, 280, 250); m["trainIdx"] = rng.integers(0, 300, 250); m["trainIdx"][:5] = np.arange(5)
    fx, fy, cx, cy = np.float32(615.3), np.float32(615.9), np.float32(320.2), np.float32(240.4)
    p3, p2 = oracle.pnp_points(kp, kc, m, depth, fx, fy, cx, cy)
    exp3, exp2 = [], []
    for q, tr in zip(m["queryIdx"], m["trainIdx"]):
        px, py = kp["x"][tr], kp["y"][tr]
        xp = int(np.floor(abs(px) + np.float32(0.5)) * np.sign(px)); yp = int(np.floor(abs(py) + np.float32(0.5)) * np.sign(py))
        if xp < 0 or yp < 0 or xp >= w or yp >= h:
            continue
        d = np.float32(depth[yp, xp]) * np.float32(0.001)
        if d <= np.float32(0.3) or d > np.float32(3.0):
            continue
        exp3.append([np.float32(np.float32((px - cx) * d) / fx), np.float32(np.float32((py - cy) * d) / fy), d]); exp2.append([kc["x"][q], kc["y"][q]])
    assert len(p3) == len(exp3) and len(p3) > 50
    assert np.array_equal(p3.view(np.uint32), np.array(exp3, np.float32).view(np.uint32)) and np.array_equal(p2, np.array(exp2, np.float32))


@pytest.mark.gpu
def test_gpu_pose_scoring_and_correspondences(built, oracle):
    """orbx_pnp_score: counts and masks of 300 pose hypotheses equal to the oracle's (= OpenCV's) per hypothesis, the winner = most inliers at the lowest
    index; orbx_pnp_points equal to the oracle's loop incl. rounding, depth gate and order"""
    import orbx
    rng = np.random.default_rng(17)
    ex = orbx.ORBextractor(max_width=640, max_height=480)
    try:
        X, uv, a, R, t, K = _scene(rng, 900)
        X[:10, 2] *= -1
        nh = 300
        Rs, ts = [], []
        for i in range(nh):
            da = a + rng.normal(0, 0.01 * (i % 7), 3); th = np.linalg.norm(da); k = da / th
            Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            Rs.append(np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx); ts.append(t + rng.normal(0, 0.01 * (i % 5), 3))
        Rs[7] = Rs[3].copy(); ts[7] = ts[3].copy()                                 # a tie: the lower index must win if it is the best
        counts, best, mask = ex.pnp_score(X, uv, np.array(Rs), np.array(ts), *K, threshold=4.0)
        want = [oracle.pnp_inliers(X, uv, Rs[i], ts[i], *K, thresh=4.0) for i in range(nh)]
        assert np.array_equal(counts, [w[0] for w in want])
        wbest = int(np.argmax(counts))                                            # argmax returns the first maximum
        assert best == wbest and np.array_equal(mask, want[wbest][1])
        assert counts.max() > 500 and counts.min() < counts.max()
        # correspondences
        w, h = 640, 480
        depth = rng.integers(0, 4000, (h, w), dtype=np.uint16)
        kp = np.zeros(1500, orbx.KP_DTYPE); kc = np.zeros(1400, orbx.KP_DTYPE)
        kp["x"] = rng.uniform(-2, w + 2, 1500).astype(np.float32); kp["y"] = rng.uniform(-2, h + 2, 1500).astype(np.float32)
        kp["x"][:5] = [0.5, 1.5, 2.5, 639.5, 639.4999]
        kc["x"] = rng.uniform(0, w, 1400).astype(np.float32); kc["y"] = rng.uniform(0, h, 1400).astype(np.float32)
        m = np.zeros(2300, orbx.DM_DTYPE)                                          # more than one 1024-match chunk
        m["queryIdx"] = rng.integers(0, 1400, 2300); m["trainIdx"] = rng.integers(0, 1500, 2300); m["trainIdx"][:5] = np.arange(5)
        fx, fy, cx, cy = 615.3, 615.9, 320.2, 240.4
        p3, p2 = ex.pnp_points(kp, kc, m, depth, fx, fy, cx, cy)
        w3, w2 = oracle.pnp_points(kp.view(oracle.KP_DTYPE), kc.view(oracle.KP_DTYPE), m.view(oracle.DM_DTYPE), depth, fx, fy, cx, cy)
        assert len(p3) == len(w3) and len(p3) > 500
        assert np.array_equal(p3.view(np.uint32), w3.view(np.uint32)) and np.array_equal(p2, w2)
        m["trainIdx"][100] = 1500                                                  # outside the previous keypoint array
        with pytest.raises(orbx.OrbxError):
            ex.pnp_points(kp, kc, m, depth, fx, fy, cx, cy)
    finally:
        ex.close()
