"""Pins the oracle (oracle/exact_search.py) against everything the reference ships
for the path: the WAL + Chroma's own materialised segment in vector_store/, and
the known answers of SURVEY.md 8c.  CPU only."""
import numpy as np
import pytest

from oracle.exact_search import (OracleCollection, distance_matrix, exact_search, fast_topk_f32,
                                 normalise_rows, prepare_corpus, round_to_bf16, topk_stable)


def build(golden, space="l2"):
    oc = OracleCollection(space=space)
    for rec, v in zip(golden["wal"]["records"], golden["vectors"]):
        oc.upsert([rec["id"]], [v], [rec["metadata"]], [rec["document"]])
    return oc


def test_wal_replay_matches_chromas_own_segment(golden):
    """37 UPSERT records -> the 25 ids, metadata and documents Chroma 0.5.3 itself
    materialised (a reference OUTPUT of the upsert path): last write wins."""
    oc = build(golden)
    seg = golden["segment"]
    assert len(golden["wal"]["records"]) == 37 and seg["count"] == 25 and seg["max_seq_id"] == 37
    assert oc.count() == 25
    assert set(oc.row_of) == set(seg["ids"])
    for id_, want in seg["ids"].items():
        r = oc.row_of[id_]
        assert oc.metas[r] == want["metadata"], id_
        assert oc.docs[r] == want["document"], id_


def test_shipped_vectors_are_unit_norm_and_reupserts_identical(golden):
    v = golden["vectors"]
    assert v.shape == (37, 384) and v.dtype == np.float32
    assert np.allclose(np.linalg.norm(v.astype(np.float64), axis=1), 1.0, atol=2e-7)
    seen = {}
    for rec, vec in zip(golden["wal"]["records"], v):
        if rec["id"] in seen:
            assert np.array_equal(seen[rec["id"]], vec)
        seen[rec["id"]] = vec


@pytest.mark.parametrize("case", ["q1", "q2", "q3"])
def test_known_answers(golden, case):
    oc = build(golden)
    c = golden["known"][case]
    res = oc.query([oc.vector_of(c["query_id"])], n_results=c["k"], where=c["where"])
    assert res["ids"][0] == c["ids"]
    assert np.allclose(res["distances"][0], c["distances"], rtol=0, atol=1e-9)


def test_known_answer_values_quoted_in_survey(golden):
    # SURVEY.md 8c, first known answer, l2 space
    d = golden["known"]["q1"]["distances"]
    assert np.allclose(d, [0.0, 0.2736222, 0.4299542, 0.4920169, 0.7895305], atol=5e-7)
    assert golden["known"]["q1"]["ids"][0] == "fyp_core::summary"
    assert np.allclose(golden["known"]["q2"]["distances"],
                       [0.9466366, 1.0657333, 1.1426963, 1.1821779, 1.1928142], atol=5e-7)


def test_cosine_is_half_l2_on_unit_vectors(golden):
    l2, cos = build(golden, "l2"), build(golden, "cosine")
    q = l2.vector_of("fyp_core::summary")
    a, b = l2.query([q], 25), cos.query([q], 25)
    assert a["ids"][0][:10] == b["ids"][0][:10]
    assert np.allclose(np.array(a["distances"][0]) / 2, b["distances"][0], atol=3e-7)


def test_distance_definitions():
    rng = np.random.default_rng(0)
    q, x = rng.standard_normal((3, 16)).astype(np.float32), rng.standard_normal((7, 16)).astype(np.float32)
    l2 = distance_matrix("l2", q, x)
    ip = distance_matrix("ip", q, x)
    for i in range(3):
        for j in range(7):
            assert np.isclose(l2[i, j], np.sum((q[i].astype(np.float64) - x[j]) ** 2))
            assert np.isclose(ip[i, j], 1 - np.dot(q[i].astype(np.float64), x[j]))
    qn, xn = normalise_rows(q), normalise_rows(x)
    cos = distance_matrix("cosine", qn, xn)
    ref = 1 - (q @ x.T) / (np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(x, axis=1)[None, :])
    assert np.allclose(cos, ref, atol=1e-6)


def test_topk_is_stable_on_ties():
    d = np.array([[1.0, 0.5, 0.5, 2.0, 0.5]])
    (rows, dd), = topk_stable(d, 3)
    assert rows.tolist() == [1, 2, 4] and dd.tolist() == [0.5, 0.5, 0.5]
    (rows, _), = topk_stable(d, 3, np.array([1, 0, 1, 1, 1], bool))
    assert rows.tolist() == [2, 4, 0]
    (rows, _), = topk_stable(d, 10)
    assert rows.tolist() == [1, 2, 4, 0, 3]


def test_bf16_rounding_is_rne_and_idempotent():
    x = np.array([1.0, 1.00390625, 1.005859375, -3.140625, 1e-20, 65504.0], np.float32)
    r = round_to_bf16(x)
    assert np.array_equal(round_to_bf16(r), r)
    assert r[0] == 1.0 and r[1] == 1.0            # tie -> even mantissa
    assert r[2] == np.float32(1.0078125)
    import torch
    t = torch.from_numpy(np.random.default_rng(1).standard_normal(4096).astype(np.float32))
    assert np.array_equal(round_to_bf16(t.numpy()), t.to(torch.bfloat16).to(torch.float32).numpy())


def test_fast_blas_search_agrees_with_exact():
    rng = np.random.default_rng(5)
    for space in ("l2", "ip", "cosine"):
        x = prepare_corpus(space, rng.standard_normal((5000, 64)).astype(np.float32))
        q = prepare_corpus(space, rng.standard_normal((9, 64)).astype(np.float32))
        rows, d = fast_topk_f32(space, q, x, 10, chunk=1024)
        er, ed = exact_search(space, q, x, 10)
        assert np.array_equal(rows, np.stack(er))
        assert np.allclose(d, np.stack(ed), rtol=1e-5, atol=2e-6)


def test_collection_semantics():
    oc = OracleCollection("l2")
    oc.add(["a", "b"], [[1, 0], [0, 1]], [{"t": "x"}, {"t": "y"}], ["da", "db"])
    oc.add(["a"], [[5, 5]], [{"t": "z"}], ["new"])              # existing id: skipped
    assert oc.count() == 2 and oc.docs[oc.row_of["a"]] == "da"
    oc.upsert(["a"], [[2, 0]], [{"t": "z"}], ["up"])            # replace in place
    assert oc.count() == 2 and oc.row_of["a"] == 0 and oc.docs[0] == "up"
    with pytest.raises(ValueError):
        oc.add(["c", "c"], [[0, 0], [1, 1]])
    with pytest.raises(ValueError):
        oc.add(["c"], [[0, 0, 0]])
    assert oc.delete(where={"t": "y"}) == ["b"] and oc.count() == 1
    assert oc.delete(ids=["nope"]) == []
    res = oc.query([[2, 0]], n_results=10)
    assert res["ids"] == [["a"]] and res["distances"] == [[0.0]]
    assert oc.query([[0, 0]], 5, where={"t": "none"})["ids"] == [[]]
