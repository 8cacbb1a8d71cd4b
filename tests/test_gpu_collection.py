"""Collection semantics on the real engine vs OracleCollection, and the golden
fixture end to end (reference call shapes of api/app.py).  GPU only."""
import random

import numpy as np
import pytest

import local_rag_system_b200 as rag
from oracle.exact_search import OracleCollection

pytestmark = pytest.mark.gpu


@pytest.fixture
def client():
    c = rag.EphemeralClient()
    yield c
    c.reset()


@pytest.mark.parametrize("space,dtype", [("l2", "f32"), ("cosine", "f32"), ("cosine", "bf16")])
def test_golden_known_answers_on_device(client, golden, space, dtype):
    col = client.get_or_create_collection("gamefantasy", metadata={"hnsw:space": space, "b200:dtype": dtype})
    oc = OracleCollection(space, dtype)
    for rec, v in zip(golden["wal"]["records"], golden["vectors"]):
        col.upsert(ids=[rec["id"]], embeddings=[v.tolist()], metadatas=[rec["metadata"]], documents=[rec["document"]])
        oc.upsert([rec["id"]], [v], [rec["metadata"]], [rec["document"]])
    assert col.count() == 25
    ids = [r["id"] for r in golden["wal"]["records"]]
    for case in ("q1", "q2", "q3"):
        c = golden["known"][case]
        q = golden["vectors"][ids.index(c["query_id"])]
        res = col.query(query_embeddings=[q.tolist()], n_results=max(1, min(c["k"], 20)), where=c["where"],
                        include=["documents", "metadatas", "distances"])
        want = oc.query([q], c["k"], where=c["where"])
        if dtype == "f32":
            assert res["ids"][0] == want["ids"][0]
            assert np.allclose(res["distances"][0], want["distances"][0], rtol=1e-5, atol=1e-6)
            if space == "l2":       # the reference's space: the SURVEY 8c numbers themselves
                assert res["ids"][0] == c["ids"]
                assert np.allclose(res["distances"][0], c["distances"], rtol=1e-5, atol=1e-6)
        else:
            assert res["ids"][0][0] == want["ids"][0][0]
            assert np.allclose(res["distances"][0], want["distances"][0], rtol=0, atol=2e-2)
        assert res["documents"][0][0] == want["documents"][0][0]
        assert res["metadatas"][0] == [oc.metas[oc.row_of[i]] for i in res["ids"][0]]
    seg = golden["segment"]["ids"]
    got = col.get()
    assert sorted(got["ids"]) == sorted(seg)
    for i, m, d in zip(got["ids"], got["metadatas"], got["documents"]):
        assert m == seg[i]["metadata"] and d == seg[i]["document"]


def test_random_op_sequences_match_the_oracle_model(client):
    rng = random.Random(5)
    nrng = np.random.default_rng(5)
    dim = 48
    col = client.get_or_create_collection("ops", metadata={"hnsw:space": "l2"})
    oc = OracleCollection("l2")
    pool = [f"id{j}" for j in range(400)]
    for step in range(120):
        op = rng.choice(["add", "add", "upsert", "upsert", "delete_ids", "delete_where", "query", "query", "query"])
        if op in ("add", "upsert"):
            ids = rng.sample(pool, rng.randint(1, 12))
            emb = nrng.standard_normal((len(ids), dim)).astype(np.float32)
            metas = [{"source_key": f"k{rng.randint(0, 15)}", "n": rng.randint(0, 9)} for _ in ids]
            docs = [f"doc {i} {step}" for i in ids]
            getattr(col, op)(ids=ids, embeddings=emb.tolist(), metadatas=metas, documents=docs)
            getattr(oc, op)(ids, emb, metas, docs)
        elif op == "delete_ids":
            ids = rng.sample(pool, rng.randint(1, 6))
            col.delete(ids=ids)
            oc.delete(ids=ids)
        elif op == "delete_where":
            w = {"source_key": f"k{rng.randint(0, 15)}"}
            assert sorted(col.delete(where=w)) == sorted(oc.delete(where=w))
        else:
            B, k = rng.randint(1, 4), rng.choice([1, 5, 10, 50])
            q = nrng.standard_normal((B, dim)).astype(np.float32)
            w = rng.choice([None, {"n": {"$lt": 5}}, {"source_key": f"k{rng.randint(0, 15)}"},
                            {"$or": [{"n": 1}, {"n": 2}]}])
            res = col.query(query_embeddings=q.tolist(), n_results=k, where=w)
            want = oc.query(q, k, where=w)
            for b in range(B):
                assert res["ids"][b] == want["ids"][b], (step, w)
                assert np.allclose(res["distances"][b], want["distances"][b], rtol=1e-5, atol=1e-5)
                assert res["documents"][b] == want["documents"][b]
        assert col.count() == oc.count()


def test_concurrent_readers_and_writers(client):
    """FastAPI worker threads query while BackgroundTasks add/delete (api/routes/kb.py)."""
    import threading
    dim = 64
    col = client.get_or_create_collection("threads")
    base = np.random.default_rng(0).standard_normal((2000, dim)).astype(np.float32)
    col.add(ids=[f"b{j}" for j in range(2000)], embeddings=base.tolist(),
            metadatas=[{"namespace": "a" if j % 2 else "b"} for j in range(2000)])
    errors = []

    def reader(seed):
        rng = np.random.default_rng(seed)
        try:
            for _ in range(30):
                q = rng.standard_normal((1, dim)).astype(np.float32)
                res = col.query(query_embeddings=q.tolist(), n_results=5,
                                where={"namespace": "a"} if seed % 2 else None)
                assert len(res["ids"][0]) == 5
                d = res["distances"][0]
                assert all(d[i] <= d[i + 1] for i in range(4))
        except Exception as e:      # noqa: BLE001
            errors.append(e)

    def writer():
        rng = np.random.default_rng(77)
        try:
            for j in range(40):
                col.add(ids=[f"w{j}"], embeddings=rng.standard_normal((1, dim)).astype(np.float32).tolist(),
                        metadatas=[{"namespace": "a", "source_key": f"s{j}"}])
                if j % 3 == 0:
                    col.delete(where={"source_key": f"s{j}"})
        except Exception as e:      # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=reader, args=(s,)) for s in range(6)] + [threading.Thread(target=writer)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    assert col.count() == 2000 + 40 - 14


def test_persistence_roundtrip_and_chroma_import(tmp_path, golden):
    """Journal replay: a second process-lifetime (registry reset) sees the same
    collection; deletes and upserts survive."""
    from local_rag_system_b200.collection import _reset_registry_for_tests
    path = str(tmp_path / "store")
    col = rag.PersistentClient(path=path).get_or_create_collection("gamefantasy")
    for rec, v in zip(golden["wal"]["records"], golden["vectors"]):
        col.upsert(ids=[rec["id"]], embeddings=[v.tolist()], metadatas=[rec["metadata"]], documents=[rec["document"]])
    col.delete(ids=["fyp_core::summary"])
    q = golden["vectors"][1]
    before = col.query(query_embeddings=[q.tolist()], n_results=5)
    _reset_registry_for_tests()
    col2 = rag.PersistentClient(path=path).get_or_create_collection("gamefantasy")
    assert col2.count() == 24
    after = col2.query(query_embeddings=[q.tolist()], n_results=5)
    assert after["ids"] == before["ids"] and after["documents"] == before["documents"]
    assert np.allclose(after["distances"][0], before["distances"][0], atol=1e-7)
    _reset_registry_for_tests()


def test_shim_import_on_device(tmp_path, golden):
    """Drop-in at the reference's own boundary ON THE ENGINE: a subprocess with shim/ first on PYTHONPATH does
    `import chromadb` and replays the reference's call shapes (tests/shim_on_device_script.py) -- no test double."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "shim"), root]), PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "shim_on_device_script.py"), str(tmp_path / "vs")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    assert out["count"] == 25 and out["engine"] == "DeviceStore" and out["kernel_launches"] > 0 and out["regime"] == "stream"
    for name, a in out["known"].items():
        assert a["ok"], (name, a, golden["known"][name])
    assert out["uris_none"]
    assert out["count_after_add"] == 26 and out["count_after_delete_where"] == 25 and out["count_after_delete_ids"] == 24
    assert os.path.exists(tmp_path / "vs" / "rag_b200.sqlite3")
