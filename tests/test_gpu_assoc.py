"""GPU parity of the landmark database paths against the oracle (Backend::associateObservation, reference backend.cpp:1064-1120):
descriptor stage (radius 50 candidates, top-2), the reprojection-gated association as one call, and the shard merges."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(oracle, nrows, nq, seed):
    rng = np.random.default_rng(seed)
    rows = oracle.synth_descriptors(1234, 0, nrows)
    A = rng.standard_normal((3, 3))
    R, _ = np.linalg.qr(A)
    if np.linalg.det(R) < 0:
        R[:, 0] = -R[:, 0]
    t = rng.standard_normal(3) * 0.3
    fx, fy, cx, cy = 615.3, 615.9, 640.2, 360.4
    # landmark positions: in front of the camera (mostly), a few behind it
    pc = np.stack([rng.uniform(-2, 2, nrows), rng.uniform(-1.2, 1.2, nrows), rng.uniform(0.4, 6.0, nrows)], 1)
    pc[rng.random(nrows) < 0.02, 2] *= -1
    pos = (pc @ R.T + t).astype(np.float32)                       # X = R * pc + t  =>  R.t() * (X - t) = pc
    src = rng.integers(0, nrows, nq)
    q = rows[src].copy()
    flips = rng.integers(0, 70, nq)                               # some beyond the 50-bit radius
    for i in range(nq):
        for b in rng.choice(256, flips[i], replace=False):
            q[i, b >> 3] ^= 1 << (b & 7)
    uv = np.stack([oracle.reproject(pos[j], R, t, fx, fy, cx, cy) for j in src])
    qpx = (uv + rng.normal(0, 2.5, (nq, 2))).astype(np.float32)    # some beyond the 5-px gate
    # near-duplicate landmarks: a second landmark with the same descriptor a few pixels away competes on reprojection error
    dup = rng.choice(nrows, 64, replace=False)
    rows[dup] = rows[src[:64]]
    pos[dup] = pos[src[:64]] + rng.normal(0, 0.004, (64, 3)).astype(np.float32)
    return rows, pos, q, qpx, R, t, (fx, fy, cx, cy)


def _assert_assoc(got, want):
    idx, err, dist = want
    assert np.array_equal(got["landmark"], idx), "associated landmark rows"
    assert (idx >= 0).any() and (idx < 0).any()
    sel = idx >= 0
    assert np.array_equal(got["reproj_error"][sel].view(np.uint64), err[sel].view(np.uint64)), "reprojection errors are bit-identical"
    assert np.array_equal(got["distance"][sel], dist[sel])


@pytest.mark.parametrize("engine", [0, 2, 3])
def test_associate_matches_reference_semantics(built, oracle, engine):
    """engine: ORBX_OPT_MATCH_MMA 0 = k_assoc_partial (POPC), 2 = k_assoc_mma (mma.sync int8 GEMM), 3 = k_assoc_umma (tcgen05), each with the same reprojection epilogue"""
    import orbx
    rows, pos, q, qpx, R, t, K = _scene(oracle, 20000 + 37 * engine, 700 + engine, 5)      # row / query counts off the 64 / 16 tile sizes
    ex = orbx.ORBextractor(max_width=320, max_height=240)
    ex.set_match_mma(engine)
    db = orbx.LandmarkDB(ex, 32768)
    try:
        db.append(rows)
        db.set_positions(pos)
        got = db.associate(q, qpx, orbx.LandmarkDB.pose(R, t, *K))
        _assert_assoc(got, oracle.associate(q, qpx, rows, pos, R, t, *K))
        # thresholds are parameters (reference backend.cpp:225-226)
        got = db.associate(q, qpx, orbx.LandmarkDB.pose(R, t, *K), max_desc_dist=30.0, max_reproj_err=2.0)
        _assert_assoc(got, oracle.associate(q, qpx, rows, pos, R, t, *K, max_desc=30.0, max_reproj=2.0))
    finally:
        db.close(); ex.close()


@pytest.mark.parametrize("engine", [0, 2, 3])
def test_associate_sharded_merge(built, oracle, engine):
    """Two shards (global row ranges) merged by orbx_merge_assoc_device == one unsharded database."""
    import torch
    import orbx
    rows, pos, q, qpx, R, t, K = _scene(oracle, 12001, 300, 9)
    ex = orbx.ORBextractor(max_width=320, max_height=240)
    ex.set_match_mma(engine)
    cut = 7000
    dbs = [orbx.LandmarkDB(ex, 8192, first_index=0), orbx.LandmarkDB(ex, 8192, first_index=cut)]
    try:
        dbs[0].append(rows[:cut]); dbs[0].set_positions(pos[:cut])
        dbs[1].append(rows[cut:]); dbs[1].set_positions(pos[cut:])
        pose = orbx.LandmarkDB.pose(R, t, *K)
        parts = np.stack([d.associate(q, qpx, pose) for d in dbs])
        dev = torch.device("cuda", 0)
        d_parts = torch.from_numpy(parts.view(np.uint8).reshape(2, len(q), 16)).to(dev)
        d_out = torch.zeros((len(q), 16), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ex._check(ex.L.orbx_merge_assoc_device(ex.handle, d_parts.data_ptr(), 2, len(q), d_out.data_ptr()))
        ex.sync()
        got = d_out.cpu().numpy().view(orbx.ASSOC_DTYPE).reshape(-1)
        _assert_assoc(got, oracle.associate(q, qpx, rows, pos, R, t, *K))
    finally:
        for d in dbs:
            d.close()
        ex.close()


def test_descriptor_stage_radius_and_top2(built, oracle):
    import orbx
    rows, _, q, _, _, _, _ = _scene(oracle, 9000, 200, 3)
    ex = orbx.ORBextractor(max_width=320, max_height=240)
    db = orbx.LandmarkDB(ex, 16384, first_index=100)
    try:
        db.append(rows[:5000]); db.append(rows[5000:])
        assert db.rows == 9000
        # every landmark with distance < 50 (backend.cpp:1074-1076), sorted by (query, landmark)
        cand = db.query_radius(q, 50.0)
        bits = np.unpackbits(q[:, None, :] ^ rows[None, :, :], axis=2).sum(2)
        want = [(i, 100 + j, float(bits[i, j])) for i in range(len(q)) for j in np.nonzero(bits[i] < 50)[0]]
        assert [(int(m["queryIdx"]), int(m["trainIdx"]), float(m["distance"])) for m in cand] == want
        top = db.query_top2(q)
        k2 = oracle.knn2(q, rows)
        assert np.array_equal(top["dist0"], k2[:, 0]["distance"].astype(np.uint32)) and np.array_equal(top["idx0"], (k2[:, 0]["trainIdx"] + 100).astype(np.uint32))
        assert np.array_equal(top["dist1"], k2[:, 1]["distance"].astype(np.uint32)) and np.array_equal(top["idx1"], (k2[:, 1]["trainIdx"] + 100).astype(np.uint32))
    finally:
        db.close(); ex.close()
