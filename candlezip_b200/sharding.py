"""Multi-GPU sharding of the coding path: one process per GPU, each rank takes a contiguous range of the
independently coded segments, holds its own weight replica and runs cz_encode / cz_decode on its range.  There is NO
collective on the data path: chunks and segments are independent given the token ids.  torch.distributed is used only to
hand the per-segment bitstreams (encode) or token ids (decode) back to rank 0, which concatenates them and prefix-sums the
offsets into the container's SEG1 table.  The result is byte-identical for every world size.

The encode/decode callables are injected so that the host logic is testable without a GPU (tests/test_multi_rank.py runs
it with world_size 2 over gloo).
"""
import numpy as np

from .api import split_segments


def shard_range(n_segments, world, rank):
    """contiguous segment range [g0, g1) of `rank`; ranges differ in size by at most one segment"""
    base, rem = divmod(n_segments, world)
    g0 = rank * base + min(rank, rem)
    return g0, g0 + base + (1 if rank < rem else 0)


def _gather(obj, rank, world, dist):
    if world == 1 or dist is None:
        return [obj]
    out = [None] * world if rank == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out


def encode_sharded(encode_fn, ids, n_segments, rank=0, world=1, dist=None):
    """encode_fn(ids_slice, local_seg_start) -> list of payload bytes, one per local segment (e.g. a closure over
    Model.encode(..., seg_start=...)).  Returns (payloads, seg_start) on rank 0 and (None, seg_start) elsewhere."""
    ids = np.ascontiguousarray(ids, np.uint32)
    seg_start = split_segments(len(ids), n_segments)
    S = len(seg_start) - 1
    g0, g1 = shard_range(S, world, rank)
    a, b = int(seg_start[g0]), int(seg_start[g1])
    local = encode_fn(ids[a:b], (seg_start[g0 : g1 + 1] - seg_start[g0]).astype(np.uint64)) if g1 > g0 else []
    assert len(local) == g1 - g0
    parts = _gather(local, rank, world, dist)
    if rank != 0:
        return None, seg_start
    return [p for part in parts for p in part], seg_start


def decode_sharded(decode_fn, payloads, seg_start, rank=0, world=1, dist=None):
    """decode_fn(local_payloads, local_seg_start) -> uint32 ids of the local segments.  Every rank holds the container
    (payloads, seg_start); rank 0 gets the concatenated ids."""
    seg_start = np.asarray(seg_start, np.uint64)
    S = len(seg_start) - 1
    g0, g1 = shard_range(S, world, rank)
    local = decode_fn(payloads[g0:g1], (seg_start[g0 : g1 + 1] - seg_start[g0]).astype(np.uint64)) if g1 > g0 else np.zeros(0, np.uint32)
    parts = _gather(np.asarray(local, np.uint32), rank, world, dist)
    if rank != 0:
        return None
    return np.concatenate(parts) if parts else np.zeros(0, np.uint32)
