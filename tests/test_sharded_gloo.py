"""N > 1 host logic on CPU: world_size-2 gloo.  Exercises the product's shard plan
and candidate exchange (local_rag_system_b200.sharded) with candidates produced
by the oracle and keys packed by the C-ABI host helpers; the merged result must
equal the oracle's answer on the whole corpus, ties included (keys carry global
rows, so a plain 64-bit sort is the cross-shard order)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import unit_rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, dim, B, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from local_rag_system_b200 import _native
        from local_rag_system_b200.sharded import exchange_candidates, shard_plan
        from oracle.exact_search import exact_search
        lib = _native.load()
        x = unit_rows(n, dim, 3)
        x[n // 2 + 5] = x[7]                     # an exact duplicate living in the other shard
        q = unit_rows(B, dim, 4)
        q[0] = x[7]
        stride, counts = shard_plan(n, world)
        lo = rank * stride
        rows, d = exact_search("cosine", q, x[lo:lo + counts[rank]], k)
        keys = np.full((B, k), _native.EMPTY_KEY, dtype=np.uint64)
        for b in range(B):
            for j, (r, dd) in enumerate(zip(rows[b], d[b])):
                keys[b, j] = lib.rag_key_pack(float(np.float32(dd)), int(r) + lo)       # global row in the key
        local = torch.from_numpy(keys.view(np.int64).reshape(-1).copy())
        gathered = exchange_candidates(local, world)                                    # [world, B*k], product code
        allk = gathered.numpy().view(np.uint64).reshape(world, B, k)
        merged = np.sort(allk.transpose(1, 0, 2).reshape(B, world * k), axis=1)[:, :k]
        got_rows = (merged & np.uint64(0xFFFFFFFF)).astype(np.int64)
        got_d = np.array([[lib.rag_key_dist(int(v)) for v in row] for row in merged])
        want_rows, want_d = exact_search("cosine", q, x, k)
        ok = all(np.array_equal(got_rows[b], want_rows[b]) for b in range(B)) and \
            np.allclose(got_d, np.stack(want_d), atol=1e-6)
        ret[rank] = (bool(ok), got_rows[0].tolist(), want_rows[0].tolist())
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_and_merge_matches_single_store():
    world, n, dim, B, k = 2, 1001, 32, 5, 10
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, dim, B, k, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        ok, got, want = ret[rank]
        assert ok, (rank, got, want)
    assert ret[0][1] == ret[1][1]                 # every rank holds the same merged answer
    assert ret[0][1][:2] == [7, n // 2 + 5]       # duplicate pair: lower global row first


def test_shard_plan():
    from local_rag_system_b200.sharded import shard_plan
    stride, counts = shard_plan(10_000_000, 8)
    assert stride == 1_250_000 and counts == [1_250_000] * 8
    stride, counts = shard_plan(10, 4)
    assert stride == 3 and counts == [3, 3, 3, 1] and sum(counts) == 10
    stride, counts = shard_plan(2, 4)
    assert counts == [1, 1, 0, 0]
    with pytest.raises(ValueError):
        shard_plan(2 ** 33, 2)
