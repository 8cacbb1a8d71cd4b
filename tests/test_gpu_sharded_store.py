"""The single-process multi-device store (rag_sharded_*, Collection metadata "b200:devices") on ONE
device: logical shards that share device 0 take the per-shard search + key gather + merge-kernel path
(the fused one-launch path needs distinct devices; tests/multi_gpu_check.py and the 2-8 GPU runs cover
it).  Everything must equal a single store over the same rows bit for bit: global rows are dense and
keys carry them.  GPU only."""
import numpy as np
import pytest

import local_rag_system_b200 as rag
from local_rag_system_b200 import DeviceStore, ShardedDeviceStore
from oracle.exact_search import OracleCollection, round_to_bf16
from tests.conftest import unit_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("G", [2, 3, 8])
@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_sharded_store_equals_single_store(G, dtype):
    n, dim = 9_500, 192                      # 9.3 chunks of 1024 rows: ragged over every G, some shards short
    x = unit_rows(n, dim, 61)
    if dtype == "bf16":
        x = round_to_bf16(x)
    x[7000] = x[100]                         # duplicate in another chunk (another shard): tie by global row
    one = DeviceStore(dim, dtype, "cosine", rerank=False)
    many = ShardedDeviceStore(dim, dtype, "cosine", devices=[0] * G, rerank=False)
    try:
        assert many.shards() == G and not many.fused
        # mixed write pattern: a bulk load, then small appends (parked writes), then in-place upserts
        for st in (one, many):
            r = st.upsert(x[:6000])
            assert r.tolist() == list(range(6000))
            for s in range(6000, n, 50):
                st.upsert(x[s:s + 50])
            assert st.count() == n and st.rows() == n
        assert sum(many.shard_counts()) == n and max(many.shard_counts()) - min(many.shard_counts()) <= 1024
        assert np.array_equal(many.fetch(np.arange(n)), one.fetch(np.arange(n)))
        rng = np.random.default_rng(3)
        dead = rng.choice(n, 700, replace=False)
        dead = dead[(dead != 100) & (dead != 7000)]
        passing = rng.random(n) < 0.35
        passing[[100, 7000]] = True
        for B, regime, k in ((1, "stream", 10), (5, "stream", 33), (40, "tensor", 10), (150, "tensor", 100)):
            q = unit_rows(B, dim, 62 + B)
            q[0] = x[100]
            if dtype == "bf16":
                q = round_to_bf16(q)
            for phase in ("dense", "tombstones", "filter"):
                slot = -1
                if phase == "tombstones":
                    one.delete(dead)
                    many.delete(dead)
                    assert many.count() == one.count()
                if phase == "filter":
                    slot = 1
                    one.set_mask(1, passing)
                    many.set_mask(1, passing)
                want = one.query(q, k, mask_slot=slot, regime=regime)
                got = many.query(q, k, mask_slot=slot, regime=regime)
                assert np.array_equal(got[2], want[2]), (B, regime, phase)
                assert np.array_equal(got[0], want[0]), (B, regime, phase)
                if dtype == "bf16" or regime == "stream":
                    assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)), (B, regime, phase)
                else:
                    assert np.allclose(got[1], want[1], rtol=1e-5, atol=1e-6)
                assert got[0][0, 0] == 100 and got[0][0, 1] == 7000
                assert many.last_query_info()["path"] == "gather"
            # row reuse: both stores hand out the same freed rows for the next appends
            y = unit_rows(20, dim, 99)
            if dtype == "bf16":
                y = round_to_bf16(y)
            assert np.array_equal(many.upsert(y), one.upsert(y))
            # incremental mask maintenance on both
            rows = np.array([5, 1500, 2047, 2048, 9000])
            bits = np.array([1, 0, 1, 1, 0], dtype=np.uint8)
            one.patch_mask(1, rows, bits) if phase == "filter" else None
            many.patch_mask(1, rows, bits) if phase == "filter" else None
            passing[rows] = bits.astype(bool)
            want = one.query(q, k, mask_slot=1, regime=regime)
            got = many.query(q, k, mask_slot=1, regime=regime)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2])
    finally:
        one.close()
        many.close()


def test_collection_over_logical_shards_matches_the_oracle_model():
    """Collection with b200:devices: same random op sequence as the single-store test, checked against
    OracleCollection (ids, distances, documents, counts)."""
    import random
    client = rag.EphemeralClient()
    try:
        col = client.get_or_create_collection("sharded", metadata={"hnsw:space": "l2", "b200:devices": "0,0,0"})
        oc = OracleCollection("l2")
        rng, nrng, dim = random.Random(11), np.random.default_rng(11), 32
        pool = [f"id{j}" for j in range(3000)]
        # enough rows to spread over several chunks and shards
        ids = pool[:2500]
        emb = nrng.standard_normal((2500, dim)).astype(np.float32)
        metas = [{"source_key": f"k{j % 16}", "n": j % 10} for j in range(2500)]
        col.add(ids=ids, embeddings=emb, metadatas=metas, documents=[f"d{j}" for j in range(2500)])
        oc.add(ids, emb, metas, [f"d{j}" for j in range(2500)])
        assert isinstance(col.device_store, ShardedDeviceStore) and col.device_store.shards() == 3
        for step in range(60):
            op = rng.choice(["add", "upsert", "delete_ids", "delete_where", "query", "query", "query"])
            if op in ("add", "upsert"):
                ids = rng.sample(pool, rng.randint(1, 12))
                emb = nrng.standard_normal((len(ids), dim)).astype(np.float32)
                metas = [{"source_key": f"k{rng.randint(0, 15)}", "n": rng.randint(0, 9)} for _ in ids]
                docs = [f"doc {i} {step}" for i in ids]
                getattr(col, op)(ids=ids, embeddings=emb.tolist(), metadatas=metas, documents=docs)
                getattr(oc, op)(ids, emb, metas, docs)
            elif op == "delete_ids":
                ids = rng.sample(pool, rng.randint(1, 30))
                col.delete(ids=ids)
                oc.delete(ids=ids)
            elif op == "delete_where":
                w = {"$and": [{"source_key": f"k{rng.randint(0, 15)}"}, {"n": {"$lt": 3}}]}
                assert sorted(col.delete(where=w)) == sorted(oc.delete(where=w))
            else:
                B, k = rng.randint(1, 4), rng.choice([1, 5, 10, 50])
                q = nrng.standard_normal((B, dim)).astype(np.float32)
                w = rng.choice([None, {"n": {"$lt": 5}}, {"source_key": f"k{rng.randint(0, 15)}"}])
                res = col.query(query_embeddings=q.tolist(), n_results=k, where=w)
                want = oc.query(q, k, where=w)
                for b in range(B):
                    assert res["ids"][b] == want["ids"][b], (step, w)
                    assert np.allclose(res["distances"][b], want["distances"][b], rtol=1e-5, atol=1e-5)
                    assert res["documents"][b] == want["documents"][b]
            assert col.count() == oc.count()
    finally:
        client.reset()
