"""sharding.py — host-side partitioning of the two sub-paths that shard.

* frame-parallel extraction (BASELINE configs[2], [4]): contiguous blocks of frames per rank, no data-path
  collective (the reference's extractor has no cross-frame state, ORBextractor.hpp:84 is scratch).
* landmark association against a large database (configs[3]; reference backend.cpp:1064-1083 loops over every
  landmark of the category): rows are sharded by contiguous GLOBAL index ranges, the queries are replicated,
  every rank computes its shard's top-2 (distance, global index) per query, ONE all-gather exchanges
  nq x 16 B per rank, and every rank merges the `world` candidates per query with the lexicographic
  (distance, index) minimum — which reproduces cv::BFMatcher's lowest-trainIdx tie-break across shards.

One process per GPU; torch.distributed is the plumbing (NCCL on the GPU box, gloo in the CPU tests).
The compute (per-shard query, merge) is the CUDA library's; the CPU tests inject stand-ins for the two kernels
to exercise the partitioning and the gather layout only.
"""

def block_range(total, world, rank):
    """Contiguous block partition: (first, count) of `total` units for `rank` of `world`; the first
    total % world ranks hold one extra unit.  Used for frames and for database rows."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad partition arguments")
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def boundary_pairs(total_frames, world):
    """Frame pairs (f, f-1) whose two frames live on different ranks under block_range:
    the previous rank's last descriptors must be carried over (or that frame re-extracted)."""
    out = []
    for r in range(1, world):
        first, cnt = block_range(total_frames, world, r)
        if cnt > 0 and first > 0:
            out.append((first, first - 1))
    return out


def stream_block(total_frames, world, rank):
    """Frame-to-frame matching over a stream that is block-partitioned across ranks (BASELINE configs[1]/[2]):
    rank r owns frames [first, first + count); to match its first frame against frame first-1 — which lives on rank r-1 — it
    re-extracts that one frame as a PREAMBLE (cheaper than a peer copy that would serialise the ranks) and discards its outputs.
    Returns (first, count, preamble) with preamble = first - 1 or None for the rank that holds frame 0."""
    first, count = block_range(total_frames, world, rank)
    return first, count, (first - 1 if first > 0 and count > 0 else None)


class ShardedLandmarkDB:
    """Row-sharded landmark database with an all-gather merge (BASELINE configs[3]).

    local_query(d_query, nq) -> per-shard top-2 tensor [nq, 4] (dist0, idx0, dist1, idx1; int32 bit patterns of u32)
    merge(gathered [world, nq, 4]) -> merged tensor [nq, 4]
    On the GPU both are liborbx kernels (orbx_db_query_top2_device / orbx_merge_top2_device).
    """

    def __init__(self, total_rows, dist=None, local_query=None, merge=None):
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.total_rows = total_rows
        self.first_index, self.rows = block_range(total_rows, self.world, self.rank)
        self.local_query = local_query
        self.merge = merge

    def query_top2(self, query, nq, gathered=None):
        """Per-shard top-2 -> all_gather_into_tensor -> merge.  Returns the merged [nq, 4] tensor (every rank holds it)."""
        import torch
        part = self.local_query(query, nq)
        if self.dist is None or self.world == 1:
            return part
        if gathered is None:
            gathered = torch.empty((self.world,) + tuple(part.shape), dtype=part.dtype, device=part.device)
        # concatenated layout [world * nq, 4] (the form both gloo and NCCL accept) == [shard][nq] of orbx_merge_top2_device
        self.dist.all_gather_into_tensor(gathered.view((-1,) + tuple(part.shape[1:])), part)
        return self.merge(gathered)
