"""K6 on ONE device: G logical shards, per-shard asynchronous search with a global row base
(rag_store_query_dev), the G x B x k candidate keys laid out as an all-gather would leave
them, then the cross-shard merge kernel (rag_merge_keys_dev).  The result must be
BIT-IDENTICAL to one store over the union of the shards -- rows, distances, counts, ties
included -- because keys carry global rows and a cross-shard merge is a plain 64-bit compare
(SURVEY.md 8e "test without 8 GPUs").  GPU only; the real NCCL / peer-memory
exchange over 2-8 devices is exercised by tests/multi_gpu_check.py and by bench.py's global
verification under torchrun.
"""
import numpy as np
import pytest

from local_rag_system_b200 import DeviceStore, merge_keys_device
from oracle.exact_search import round_to_bf16
from tests.conftest import unit_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _scan_planes_only(monkeypatch):
    """This module checks the scan kernels against the oracle on the values the scan reads (bf16 rows:
    inputs rounded to bf16 on both sides).  bf16 stores are therefore created WITHOUT the fp32 re-ranking
    plane; tests/test_gpu_rerank.py covers the store as the product configures it (plane on)."""
    monkeypatch.setenv("RAG_B200_RERANK", "0")

torch = pytest.importorskip("torch")


def _bounds(n, G, seed, empty_shard):
    """Contiguous, deliberately ragged shard boundaries; optionally one shard with no rows."""
    rng = np.random.default_rng(seed)
    cuts = np.sort(rng.choice(np.arange(1, n), G - 1, replace=False))
    if empty_shard and G > 2:
        cuts[G // 2] = cuts[G // 2 - 1]
    return [0, *cuts.tolist(), n]


def _sharded_search(shards, bounds, q_dev, B, k, regime, slot):
    """Per-shard search -> [G][B][k] keys -> merge kernel; returns numpy rows / dists / counts."""
    G = len(shards)
    dev = q_dev.device
    keys = torch.empty((G, B, k), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    for g, st in enumerate(shards):
        st.query_device(q_dev.data_ptr(), B, k, keys[g].data_ptr(), stream=stream, mask_slot=slot,
                        row_base=bounds[g], regime=regime)
    rows = torch.empty((B, k), dtype=torch.int64, device=dev)
    dists = torch.empty((B, k), dtype=torch.float32, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    merge_keys_device(0, G, B, k, keys.data_ptr(), 0, rows.data_ptr(), dists.data_ptr(), counts.data_ptr(),
                      stream=stream)
    torch.cuda.synchronize(dev)
    return rows.cpu().numpy(), dists.cpu().numpy(), counts.cpu().numpy()


@pytest.mark.parametrize("G", [2, 8])
@pytest.mark.parametrize("dtype,regime,B", [("bf16", "stream", 1), ("bf16", "stream", 3), ("bf16", "tensor", 64),
                                            ("bf16", "tensor", 200), ("f32", "stream", 5)])
@pytest.mark.parametrize("k", [10, 100])
def test_emulated_shards_equal_one_store(G, dtype, regime, B, k):
    n, dim = 24_011, 384
    x = unit_rows(n, dim, 31)
    if dtype == "bf16":
        x = round_to_bf16(x)
    bounds = _bounds(n, G, seed=G + k, empty_shard=True)
    # an exact duplicate pair living in two different shards: the tie must go to the lower GLOBAL row
    a, b = bounds[1] - 1, bounds[-2] + 1
    x[b] = x[a]
    q = unit_rows(B, dim, 32 + B)
    q[0] = x[a]
    if dtype == "bf16":
        q = round_to_bf16(q)
    rng = np.random.default_rng(7)
    dead = rng.choice(n, 1500, replace=False)
    dead = dead[(dead != a) & (dead != b)]
    passing = rng.random(n) < 0.4
    passing[[a, b]] = True

    full = DeviceStore(dim, dtype, "cosine")
    shards = [DeviceStore(dim, dtype, "cosine") for _ in range(G)]
    try:
        full.upsert(x)
        for g, st in enumerate(shards):
            if bounds[g + 1] > bounds[g]:
                st.upsert(x[bounds[g]:bounds[g + 1]])
        q_dev = torch.from_numpy(q).cuda()
        for phase in ("dense", "tombstones", "filter"):
            slot = -1
            if phase == "tombstones":
                full.delete(dead)
                for g, st in enumerate(shards):
                    mine = dead[(dead >= bounds[g]) & (dead < bounds[g + 1])] - bounds[g]
                    if mine.size:
                        st.delete(mine)
            if phase == "filter":
                slot = 0
                full.set_mask(0, passing)
                for g, st in enumerate(shards):
                    st.set_mask(0, passing[bounds[g]:bounds[g + 1]])
            want_r, want_d, want_c = full.query(q, k, mask_slot=slot, regime=regime)
            got_r, got_d, got_c = _sharded_search(shards, bounds, q_dev, B, k, regime, slot)
            assert np.array_equal(got_c, want_c), phase
            assert np.array_equal(got_r, want_r), (phase, np.nonzero((got_r != want_r).any(axis=1))[0][:4])
            assert np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32)), phase
            assert got_r[0, 0] == a and got_r[0, 1] == b          # cross-shard duplicate: lower global row first
            assert full.last_query_info()["regime"] == regime
    finally:
        full.close()
        for st in shards:
            st.close()


def test_fp32_tensor_regime_shards_equal_one_store():
    """Split-precision regime per shard (approximate ranking + exact re-ranking): same rows and counts as
    one store; distances are exact fp32 in both, compared to 1e-6."""
    n, dim, B, k, G = 30_007, 192, 40, 10, 4
    x = unit_rows(n, dim, 41)
    q = unit_rows(B, dim, 42)
    bounds = _bounds(n, G, seed=3, empty_shard=False)
    full = DeviceStore(dim, "f32", "l2")
    shards = [DeviceStore(dim, "f32", "l2") for _ in range(G)]
    try:
        full.upsert(x)
        for g, st in enumerate(shards):
            st.upsert(x[bounds[g]:bounds[g + 1]])
        want_r, want_d, want_c = full.query(q, k, regime="tensor")
        got_r, got_d, got_c = _sharded_search(shards, bounds, torch.from_numpy(q).cuda(), B, k, "tensor", -1)
        assert np.array_equal(got_c, want_c) and np.array_equal(got_r, want_r)
        assert np.allclose(got_d, want_d, rtol=1e-5, atol=1e-6)
    finally:
        full.close()
        for st in shards:
            st.close()


def test_merge_kernel_fan_in_beyond_one_warp():
    """G = 40 shards (> 32: the general merge kernel, not the one-warp-per-query fast path), k = 70."""
    n, dim, B, k, G = 8_003, 96, 6, 70, 40
    x = unit_rows(n, dim, 51)
    q = unit_rows(B, dim, 52)
    bounds = _bounds(n, G, seed=9, empty_shard=True)
    full = DeviceStore(dim, "f32", "ip")
    shards = [DeviceStore(dim, "f32", "ip") for _ in range(G)]
    try:
        full.upsert(x)
        for g, st in enumerate(shards):
            if bounds[g + 1] > bounds[g]:
                st.upsert(x[bounds[g]:bounds[g + 1]])
        want = full.query(q, k, regime="stream")
        got = _sharded_search(shards, bounds, torch.from_numpy(q).cuda(), B, k, "stream", -1)
        for g_, w_ in zip(got, want):
            assert np.array_equal(g_, w_)
    finally:
        full.close()
        for st in shards:
            st.close()
