"""Host-layer semantics of the Chroma-compatible Collection (ids, metadata, `where`
compilation, result assembly, include handling, validation).  The device store
is replaced by tests/fake_store.py (an oracle-backed TEST DOUBLE) so this runs
without a GPU; tests/test_gpu_collection.py runs the same checks on the real
engine."""
import numpy as np
import pytest

import local_rag_system_b200 as rag
from local_rag_system_b200 import collection as colmod
from tests.fake_store import FakeDeviceStore


@pytest.fixture
def client(monkeypatch):
    monkeypatch.setattr(colmod, "DeviceStore", FakeDeviceStore)
    c = rag.EphemeralClient()
    yield c
    c.reset()


def load_golden(col, golden):
    for rec, v in zip(golden["wal"]["records"], golden["vectors"]):
        col.upsert(ids=[rec["id"]], embeddings=[v.tolist()], metadatas=[rec["metadata"]], documents=[rec["document"]])


def test_reference_search_call_shape(client, golden):
    """The exact call api/app.py:544-549 makes, on the reference's shipped index."""
    col = client.get_or_create_collection(name="gamefantasy", embedding_function=None)
    load_golden(col, golden)
    assert col.count() == 25
    known = golden["known"]
    for case in ("q1", "q2", "q3"):
        c = known[case]
        q = golden["vectors"][[r["id"] for r in golden["wal"]["records"]].index(c["query_id"])]
        res = col.query(query_embeddings=[q.tolist()], n_results=max(1, min(c["k"], 20)), where=c["where"],
                        include=["documents", "metadatas", "distances"])
        assert res["ids"][0] == c["ids"]
        assert np.allclose(res["distances"][0], c["distances"], atol=1e-6)
        assert len(res["documents"][0]) == len(res["metadatas"][0]) == 5
        assert res["metadatas"][0][0]["namespace"]
        # the reshaping loop of api/app.py:553-566 works on the result
        docs = res.get("documents", [[]])[0]
        assert [float(d) for d in res.get("distances", [[]])[0]] and len(docs) == 5
    seg = golden["segment"]["ids"]
    got = col.get(include=["metadatas", "documents"])
    assert sorted(got["ids"]) == sorted(seg)
    for i, m, d in zip(got["ids"], got["metadatas"], got["documents"]):
        assert m == seg[i]["metadata"] and d == seg[i]["document"]


def test_two_key_where_is_implicit_and(client, golden):
    col = client.get_or_create_collection("gamefantasy")
    load_golden(col, golden)
    q = golden["vectors"][1].tolist()
    res = col.query(query_embeddings=[q], n_results=20, where={"namespace": "history", "canonicality": "non"})
    assert res["ids"][0]
    for m in res["metadatas"][0]:
        assert m["namespace"] == "history" and m["canonicality"] == "non"
    none = col.query(query_embeddings=[q], n_results=5, where={"namespace": "nope"})
    assert none["ids"] == [[]] and none["distances"] == [[]] and none["documents"] == [[]]


def test_include_uris_does_not_raise(client, golden):
    """scripts/query_local.py:33 asks for "uris" and reads res["ids"][0] unconditionally."""
    col = client.get_or_create_collection("gamefantasy")
    load_golden(col, golden)
    res = col.query(query_embeddings=[golden["vectors"][0].tolist()], n_results=3,
                    include=["documents", "metadatas", "distances", "uris"])
    assert res["uris"] is None and len(res["ids"][0]) == 3
    res = col.query(query_embeddings=[golden["vectors"][0].tolist()], n_results=3, include=["distances"])
    assert res["documents"] is None and res["metadatas"] is None and len(res["distances"][0]) == 3
    res = col.query(query_embeddings=[golden["vectors"][0].tolist()], n_results=2, include=["embeddings"])
    assert np.allclose(res["embeddings"][0][0], golden["vectors"][0], atol=1e-7)
    with pytest.raises(ValueError):
        col.query(query_embeddings=[[0.0] * 384], include=["nonsense"])


def test_add_upsert_delete_semantics(client):
    col = client.get_or_create_collection("c")
    assert col.count() == 0
    assert col.query(query_embeddings=[[1.0, 0.0]], n_results=3)["ids"] == [[]]
    col.add(ids=["a", "b"], embeddings=[[1, 0], [0, 1]], metadatas=[{"source_key": "s1"}, {"source_key": "s2"}],
            documents=["da", "db"])
    col.add(ids=["a"], embeddings=[[9, 9]], metadatas=[{"source_key": "zzz"}], documents=["changed"])   # skipped
    assert col.count() == 2
    assert col.get(ids=["a"])["documents"] == ["da"]
    col.upsert(ids=["a", "c"], embeddings=[[2, 0], [3, 3]], metadatas=[{"source_key": "s1", "v": 2}, None],
               documents=["da2", None])
    assert col.count() == 3
    g = col.get(ids=["a"])
    assert g["documents"] == ["da2"] and g["metadatas"] == [{"source_key": "s1", "v": 2}]
    res = col.query(query_embeddings=[[2, 0]], n_results=10)           # clamps to count with a warning
    assert res["ids"][0] == ["a", "b", "c"] and res["distances"][0] == [0.0, 5.0, 10.0]
    # api/app.py:269 / 311: delete by where; :306 delete by ids; missing ids only warn
    assert col.delete(where={"source_key": "s2"}) == ["b"]
    col.delete(ids=["nope"])
    col.delete(ids=["a"])
    assert col.count() == 1 and col.query(query_embeddings=[[2, 0]], n_results=5)["ids"] == [["c"]]
    with pytest.raises(ValueError):
        col.delete()
    # freed rows are reused and stale metadata never leaks into filters
    col.add(ids=["d"], embeddings=[[0, 5]], metadatas=[{"k": "new"}])
    assert col.query(query_embeddings=[[0, 5]], n_results=5, where={"source_key": "s2"})["ids"] == [[]]
    assert col.query(query_embeddings=[[0, 5]], n_results=5, where={"k": "new"})["ids"] == [["d"]]


def test_validation(client):
    col = client.get_or_create_collection("v")
    with pytest.raises(ValueError):
        col.add(ids=["x", "x"], embeddings=[[1, 2], [3, 4]])
    with pytest.raises(ValueError):
        col.add(ids=["x"], embeddings=[[1, 2]], metadatas=[{"bad": [1, 2]}])     # api/models.py:58 may carry JSON
    with pytest.raises(ValueError):
        col.add(ids=["x"])                                                     # neither embeddings nor documents
    with pytest.raises(ValueError):
        col.add(ids=["x"], documents=["text"])                                 # no embedding function
    col.add(ids=["x"], embeddings=[[1, 2]])
    with pytest.raises(ValueError):
        col.add(ids=["y"], embeddings=[[1, 2, 3]])                             # dimension is fixed by first insert
    with pytest.raises(ValueError):
        col.query(query_embeddings=[[1, 2, 3]])
    with pytest.raises(ValueError):
        col.query(query_embeddings=[[1, 2]], n_results=0)
    with pytest.raises(ValueError):
        col.query()
    with pytest.raises(ValueError):
        col.query(query_embeddings=[[1, 2]], where={"a": {"$bogus": 1}})


def test_embedding_function_hook_and_shared_state(client):
    calls = []

    def ef(texts):
        calls.append(list(texts))
        return [[float(len(t)), 1.0] for t in texts]

    a = client.get_or_create_collection("shared", embedding_function=ef)
    b = client.get_or_create_collection("shared")          # api/app.py:268: same collection, no EF
    a.add(ids=["1", "2"], documents=["ab", "abcd"], metadatas=[{"source_key": "k"}, {"source_key": "j"}])
    assert b.count() == 2
    res = a.query(query_texts=["abc"], n_results=1)
    assert res["ids"] == [["1"]] or res["ids"] == [["2"]]
    assert calls == [["ab", "abcd"], ["abc"]]
    b.delete(where={"source_key": "k"})
    assert a.count() == 1


def test_get_paging_and_filters(client):
    col = client.get_or_create_collection("g")
    col.add(ids=[f"i{j}" for j in range(10)], embeddings=[[j, 0] for j in range(10)],
            metadatas=[{"n": j, "par": "even" if j % 2 == 0 else "odd"} for j in range(10)],
            documents=[f"doc {j}" for j in range(10)])
    assert col.get(where={"par": "even"})["ids"] == ["i0", "i2", "i4", "i6", "i8"]
    assert col.get(where={"n": {"$gte": 7}})["ids"] == ["i7", "i8", "i9"]
    assert col.get(limit=3, offset=2)["ids"] == ["i2", "i3", "i4"]
    assert col.get(where_document={"$contains": "doc 4"})["ids"] == ["i4"]
    assert col.get(ids=["i3", "zz"])["ids"] == ["i3"]
    res = col.query(query_embeddings=[[0, 0]], n_results=2, where={"$or": [{"n": 9}, {"n": {"$in": [5, 6]}}]})
    assert res["ids"] == [["i5", "i6"]]
    res = col.query(query_embeddings=[[9, 0]], n_results=2, where_document={"$not_contains": "9"})
    assert res["ids"] == [["i8", "i7"]]


def test_mask_cache_is_invalidated_by_writes(client):
    col = client.get_or_create_collection("m")
    col.add(ids=["a"], embeddings=[[0.0, 1.0]], metadatas=[{"t": "x"}])
    w = {"t": "x"}
    assert col.query(query_embeddings=[[0, 1]], n_results=5, where=w)["ids"] == [["a"]]
    col.add(ids=["b"], embeddings=[[0.0, 2.0]], metadatas=[{"t": "x"}])
    assert col.query(query_embeddings=[[0, 1]], n_results=5, where=w)["ids"] == [["a", "b"]]
    col.upsert(ids=["a"], embeddings=[[0.0, 1.0]], metadatas=[{"t": "y"}])
    assert col.query(query_embeddings=[[0, 1]], n_results=5, where=w)["ids"] == [["b"]]
    for j in range(40):   # more distinct filters than device mask slots
        assert col.query(query_embeddings=[[0, 1]], n_results=5, where={"t": f"v{j}"})["ids"] == [[]]
    assert col.query(query_embeddings=[[0, 1]], n_results=5, where=w)["ids"] == [["b"]]


def test_cached_masks_are_patched_not_rebuilt(client):
    """A write changes only the written rows' bits of every cached `where` bitmap (Collection._patch_masks);
    the predicate is evaluated over all rows once, however many writes and searches follow."""
    col = client.get_or_create_collection("p")
    col.add(ids=[f"a{j}" for j in range(50)], embeddings=[[float(j), 1.0] for j in range(50)],
            metadatas=[{"t": "x" if j % 2 else "y", "n": j} for j in range(50)])
    w1, w2 = {"t": "x"}, {"n": {"$lt": 10}}
    st = col._s
    assert len(col.query(query_embeddings=[[0, 1]], n_results=50, where=w1)["ids"][0]) == 25
    assert len(col.query(query_embeddings=[[0, 1]], n_results=50, where=w2)["ids"][0]) == 10
    assert st.mask_uploads == 2
    for j in range(50, 80):
        col.add(ids=[f"a{j}"], embeddings=[[float(j), 1.0]], metadatas=[{"t": "x", "n": j % 20}])
        assert f"a{j}" in col.query(query_embeddings=[[float(j), 1.0]], n_results=3, where=w1)["ids"][0]
    assert len(col.query(query_embeddings=[[0, 1]], n_results=100, where=w1)["ids"][0]) == 55
    assert len(col.query(query_embeddings=[[0, 1]], n_results=100, where=w2)["ids"][0]) == 10 + 10
    col.update(ids=["a1"], metadatas=[{"t": "y"}])
    col.delete(ids=["a3"])
    col.add(ids=["b"], embeddings=[[3.0, 1.0]], metadatas=[{"t": "y", "n": 99}])        # takes a3's row: bit must clear
    got = col.query(query_embeddings=[[0, 1]], n_results=100, where=w1)["ids"][0]
    assert "a1" not in got and "a3" not in got and "b" not in got and len(got) == 53
    assert st.mask_uploads == 2 and st.mask_patches > 0
    # a bulk write beyond the patch limit invalidates instead (one re-evaluation at the next query)
    big = 5000
    col.add(ids=[f"c{j}" for j in range(big)], embeddings=np.ones((big, 2), np.float32), metadatas=[{"t": "x"}] * big)
    assert len(col.query(query_embeddings=[[0, 1]], n_results=10, where=w1)["ids"][0]) == 10
    assert st.mask_uploads == 3


def test_upsert_merges_metadata_and_keeps_the_document(client):
    """Chroma's metadata segment applies an UPSERT of an existing id as an update: keys given are replaced,
    keys not given stay, the document stays when none is passed [dep: chromadb 0.5.3 _update_metadata]."""
    col = client.get_or_create_collection("u")
    col.upsert(ids=["a"], embeddings=[[1.0, 0.0]], metadatas=[{"k1": "v1", "k2": 2}], documents=["doc a"])
    col.upsert(ids=["a"], embeddings=[[0.0, 1.0]], metadatas=[{"k2": 3, "k3": True}])
    got = col.get(ids=["a"], include=["metadatas", "documents", "embeddings"])
    assert got["metadatas"] == [{"k1": "v1", "k2": 3, "k3": True}] and got["documents"] == ["doc a"]
    assert np.allclose(got["embeddings"][0], [0.0, 1.0])
    col.upsert(ids=["a"], embeddings=[[0.0, 1.0]], documents=["doc b"])
    got = col.get(ids=["a"])
    assert got["metadatas"] == [{"k1": "v1", "k2": 3, "k3": True}] and got["documents"] == ["doc b"]
    assert col.query(query_embeddings=[[0, 1]], n_results=1, where={"k1": "v1"})["ids"] == [["a"]]


def test_device_list_metadata():
    from local_rag_system_b200.collection import _parse_devices
    assert _parse_devices("0-7") == list(range(8))
    assert _parse_devices("0,2,4") == [0, 2, 4]
    assert _parse_devices("0-1, 4-5") == [0, 1, 4, 5]
    assert _parse_devices([1, 3]) == [1, 3] and _parse_devices(2) == [2]
    assert _parse_devices(None) is None and _parse_devices("") is None


def test_persistence_semantics(monkeypatch, tmp_path):
    """Journal: a dropped collection is not resurrected from a Chroma file in the same directory; the stored
    space / dtype win over a later caller's; vector-less update records replay as metadata updates."""
    from local_rag_system_b200 import persist
    from local_rag_system_b200.collection import _reset_registry_for_tests
    monkeypatch.setattr(colmod, "DeviceStore", FakeDeviceStore)
    path = str(tmp_path / "store")
    col = rag.PersistentClient(path=path).get_or_create_collection("c", metadata={"hnsw:space": "cosine"})
    col.add(ids=["a", "b"], embeddings=[[1.0, 0.0], [0.0, 1.0]], metadatas=[{"t": 1}, {"t": 2}], documents=["da", "db"])
    col.update(ids=["a"], metadatas=[{"u": 5}])
    _reset_registry_for_tests()
    col2 = rag.PersistentClient(path=path).get_or_create_collection("c", metadata={"hnsw:space": "l2"})
    assert col2._s.space == "cosine"                     # the journal was written under cosine
    assert col2.get(ids=["a"])["metadatas"] == [{"t": 1, "u": 5}] and col2.count() == 2
    # a vector-less Chroma "update" record
    state = col2._s
    persist._replay(state, [("update", "b", None, {"t": 9}, None)])
    assert col2.get(ids=["b"])["metadatas"] == [{"t": 9}] and col2.get(ids=["b"])["documents"] == ["db"]
    client = rag.PersistentClient(path=path)
    client.delete_collection("c")
    _reset_registry_for_tests()
    j = persist.Journal(str(tmp_path / "store" / persist.JOURNAL_FILE), "c")
    assert j.was_dropped() and j.stored_metadata() == (None, False)
    j.close()
    col3 = rag.PersistentClient(path=path).get_or_create_collection("c")
    assert col3.count() == 0
    _reset_registry_for_tests()
