"""CPU, world_size 2 over gloo: the host-side sharding of the landmark-association path (SURVEY §8(e), configs[3]).
The per-shard query and the merge are stood in by the oracle (the product runs them as CUDA kernels); what is under
test is the partitioning, the global-index bookkeeping and the all-gather layout."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def merge_top2_reference(parts):
    """Lexicographic (distance, index) top-2 over shards — numpy statement of orbx_merge_top2_device (test stand-in
    for the CUDA merge kernel, so that the gather layout can be checked without a GPU)."""
    p = np.asarray(parts).view(np.uint32).reshape(parts.shape[0], -1, 4)
    world, nq, _ = p.shape
    keys = np.concatenate([(p[:, :, 0].astype(np.uint64) << 32) | p[:, :, 1], (p[:, :, 2].astype(np.uint64) << 32) | p[:, :, 3]], 0)
    keys = np.sort(keys, axis=0)[:2]
    out = np.zeros((nq, 4), np.uint32)
    out[:, 0], out[:, 1] = keys[0] >> 32, keys[0] & 0xFFFFFFFF
    out[:, 2], out[:, 3] = keys[1] >> 32, keys[1] & 0xFFFFFFFF
    return out


def _top2_oracle(co, q, rows, first_index):
    """per-shard top-2 as (dist0, idx0, dist1, idx1) u32 with GLOBAL indices; 0xFFFFFFFF marks a missing entry"""
    out = np.full((len(q), 4), 0xFFFFFFFF, np.uint32)
    if len(rows):
        k2 = co.knn2(q, rows)
        for j in range(2):
            ok = k2[:, j]["trainIdx"] >= 0
            out[ok, 2 * j] = k2[ok, j]["distance"].astype(np.uint32)
            out[ok, 2 * j + 1] = (k2[ok, j]["trainIdx"] + first_index).astype(np.uint32)
    return out


def _worker(rank, world, port, total_rows, nq, result_dir):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "dynamic-visual-slam_b200", "python"))
    import c_oracle as co
    from orbx import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db_all = co.synth_descriptors(1234, 0, total_rows)
        q = db_all[::max(1, total_rows // nq)][:nq].copy()
        q[1::2, 3] ^= 0x5A                                       # half the queries are a few bits away from their row
        db_all[7] = db_all[total_rows - 1]                       # a duplicate row on the other shard: tie -> lowest global index
        q[0] = db_all[7]

        def local_query(query, n):
            first, cnt = sharding.block_range(total_rows, world, rank)
            return torch.from_numpy(_top2_oracle(co, query[:n], db_all[first:first + cnt], first).view(np.int32))

        def merge(gathered):
            return torch.from_numpy(merge_top2_reference(gathered.numpy()).view(np.int32))

        sdb = sharding.ShardedLandmarkDB(total_rows, dist=dist, local_query=local_query, merge=merge)
        assert (sdb.first_index, sdb.rows) == sharding.block_range(total_rows, world, rank)
        merged = sdb.query_top2(q, len(q)).numpy().view(np.uint32)
        want = _top2_oracle(co, q, db_all, 0)
        assert np.array_equal(merged, want), "sharded top-2 differs from the unsharded answer on rank %d" % rank
        assert merged[0, 0] == 0 and merged[0, 1] == 7 and merged[0, 3] == total_rows - 1
        np.save(os.path.join(result_dir, "rank%d.npy" % rank), merged)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total_rows,nq", [(1001, 64), (4096, 33)])
def test_sharded_association_world2(built, tmp_path, total_rows, nq):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, total_rows, nq, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b), "every rank must hold the same merged result"


def test_block_range_partitions():
    sys.path.insert(0, os.path.join(ROOT, "dynamic-visual-slam_b200", "python"))
    from orbx import sharding
    for total in (0, 1, 7, 8, 4096, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.block_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert sharding.block_range(1 << 20, 8, 3) == (3 * 131072, 131072)        # configs[3]: 131 072 rows per GPU
    assert sharding.block_range(4096, 8, 7) == (3584, 512)                    # configs[2]: 512 frames per GPU
    assert sharding.boundary_pairs(4096, 8) == [(512 * r, 512 * r - 1) for r in range(1, 8)]
    with pytest.raises(ValueError):
        sharding.block_range(10, 2, 2)
    # stream blocks: every rank but the first re-extracts its predecessor frame as a preamble
    assert sharding.stream_block(4096, 8, 0) == (0, 512, None)
    assert sharding.stream_block(4096, 8, 3) == (1536, 512, 1535)
    assert sharding.stream_block(3, 8, 5) == (3, 0, None)


def test_merge_reference_handles_missing_entries():
    sys.path.insert(0, os.path.join(ROOT, "dynamic-visual-slam_b200", "python"))
    from orbx import sharding
    F = 0xFFFFFFFF
    parts = np.array([[[5, 10, F, F]], [[5, 3, 9, 4]], [[F, F, F, F]]], np.uint32)
    assert merge_top2_reference(parts).tolist() == [[5, 3, 5, 10]]
