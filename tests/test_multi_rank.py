"""world_size-2 gloo test of the N>1 path's host logic: segment sharding across ranks, gather of the
per-segment bitstreams, container assembly.  The per-segment coder is the oracle (table-driven model) standing in for a
GPU's cz_encode / cz_decode -- the sharding code is engine-agnostic.  Invariant: identical container bytes for world 1 / 2 / 3."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

V = 300


def _table():
    rng = np.random.default_rng(1)
    return rng.normal(0, 2.0, (64, V)).astype(np.float32)


def _encode_fn(ids, seg_start):
    import oracle

    out = []
    for g in range(len(seg_start) - 1):
        s = oracle.Session.table(_table())
        seq = np.concatenate([[0], ids[int(seg_start[g]) : int(seg_start[g + 1])]]).astype(np.uint32)
        out.append(s.encode_tokens(seq)[0])
    return out


def _decode_fn(payloads, seg_start):
    import oracle

    out = []
    for g, p in enumerate(payloads):
        n = int(seg_start[g + 1] - seg_start[g])
        s = oracle.Session.table(_table())
        out.append(s.decode_tokens(p, 0, n)[0][1:])
    return np.concatenate(out) if out else np.zeros(0, np.uint32)


def _job(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import torch.distributed as dist

    from candlezip_b200 import container, sharding

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    ids = rng.integers(0, V, 2100).astype(np.uint32)
    pays, seg_start = sharding.encode_sharded(_encode_fn, ids, 7, rank, world, dist)
    blob = None
    if rank == 0:
        f = dict(token_count=len(ids), orig_len_bytes=len(ids), vocab_size=V)
        blob = container.write_container(f, b"model", pays, seg_tokens=np.diff(seg_start))
    # every rank reads the container (broadcast of the file), decodes its own segments
    lst = [blob]
    dist.broadcast_object_list(lst, src=0)
    _, _, _, _, st, payloads = container.read_container(lst[0])
    out = sharding.decode_sharded(_decode_fn, payloads, np.concatenate([[0], np.cumsum(st)]), rank, world, dist)
    if rank == 0:
        q.put((lst[0], out is not None and bool(np.array_equal(out, ids))))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_job, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_shard_ranges_cover_and_balance():
    from candlezip_b200 import sharding

    for n in (1, 7, 8, 512):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(400)
def test_world2_gloo_bytes_identical_to_world1():
    blob1, ok1 = _run(1)
    blob2, ok2 = _run(2)
    assert ok1 and ok2
    assert blob1 == blob2
