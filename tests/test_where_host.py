"""The vectorised `where` compiler (product, host side) against the oracle's
record-at-a-time evaluator, on seeded random metadata.  CPU only."""
import random

import numpy as np
import pytest

from local_rag_system_b200.where import (MetadataColumns, evaluate_where_document, validate_where,
                                         validate_where_document)
from oracle.exact_search import where_document_matches, where_matches

KEYS = ["namespace", "canonicality", "n", "score", "flag", "source_key"]


def random_meta(rng):
    m = {}
    if rng.random() < 0.9:
        m["namespace"] = rng.choice(["history", "docs", "game_systems", "media_coverage"])
    if rng.random() < 0.8:
        m["canonicality"] = rng.choice(["canon", "semi", "non"])
    if rng.random() < 0.7:
        m["n"] = rng.choice([rng.randint(-5, 5), float(rng.randint(-5, 5)), str(rng.randint(-5, 5))])
    if rng.random() < 0.5:
        m["score"] = rng.choice([0.25, 0.5, 1.5, 2.0, 7])
    if rng.random() < 0.5:
        m["flag"] = rng.choice([True, False, 1, 0])
    if rng.random() < 0.3:
        m["source_key"] = "k%d" % rng.randint(0, 9)
    return m or None


def random_where(rng, depth=0):
    r = rng.random()
    if depth < 2 and r < 0.25:
        return {rng.choice(["$and", "$or"]): [random_where(rng, depth + 1) for _ in range(rng.randint(1, 3))]}
    key = rng.choice(KEYS)
    vals = {"namespace": ["history", "docs", "zzz"], "canonicality": ["canon", "non"], "n": [0, 1, 2.0, -3, "1"],
            "score": [0.5, 1.5, 2.0, 7, 7.0], "flag": [True, False, 1, 0], "source_key": ["k1", "k5"]}[key]
    v = rng.choice(vals)
    kind = rng.random()
    if kind < 0.35:
        return {key: v}
    if kind < 0.5:
        return {key: {rng.choice(["$eq", "$ne"]): v}}
    if kind < 0.7 and isinstance(v, (int, float)) and not isinstance(v, bool):
        return {key: {rng.choice(["$gt", "$gte", "$lt", "$lte"]): v}}
    same = [w for w in vals if type(w) is type(v)]
    if kind < 0.85:
        return {key: {rng.choice(["$in", "$nin"]): rng.sample(same, rng.randint(1, len(same)))}}
    # multi-key dict = implicit $and (what api/app.py:540-542 builds)
    return {"namespace": rng.choice(["history", "docs"]), "canonicality": rng.choice(["canon", "non"])}


@pytest.mark.parametrize("seed", range(8))
def test_compiler_matches_oracle(seed):
    rng = random.Random(seed)
    n = 300
    metas = [random_meta(rng) for _ in range(n)]
    cols = MetadataColumns()
    for r, m in enumerate(metas):
        cols.set_row(r, None, m)
    # overwrite some rows (upsert) and clear some (delete) to exercise column maintenance
    for r in rng.sample(range(n), 60):
        new = random_meta(rng) if rng.random() < 0.7 else None
        cols.set_row(r, metas[r], new)
        metas[r] = new
    for _ in range(120):
        w = random_where(rng)
        validate_where(w)
        got = cols.evaluate(w, n)
        want = np.array([where_matches(w, m) for m in metas])
        assert np.array_equal(got, want), w


def test_rows_beyond_columns_and_unknown_keys():
    cols = MetadataColumns()
    cols.set_row(0, None, {"a": 1})
    assert cols.evaluate({"a": 1}, 5).tolist() == [True, False, False, False, False]
    assert cols.evaluate({"a": {"$ne": 1}}, 5).tolist() == [False, True, True, True, True]
    assert cols.evaluate({"zzz": "x"}, 3).tolist() == [False] * 3
    assert cols.evaluate({"zzz": {"$nin": ["x"]}}, 3).tolist() == [True] * 3
    assert cols.evaluate(None, 3).tolist() == [True] * 3


@pytest.mark.parametrize("bad", [
    {"a": {"$foo": 1}}, {"$and": []}, {"$and": {"a": 1}}, {"a": {"$in": []}}, {"a": {"$in": [1, "x"]}},
    {"a": {"$gt": "x"}}, {"a": {"$gt": True}}, {"a": [1, 2]}, {"a": None}, {"$not": {"a": 1}}, "a=1", {"a": {}},
])
def test_validation_rejects(bad):
    with pytest.raises(ValueError):
        validate_where(bad)


def test_where_document():
    docs = np.array(["alpha beta", "beta gamma", None, "gamma"], dtype=object)
    for wd in [{"$contains": "beta"}, {"$not_contains": "beta"},
               {"$and": [{"$contains": "beta"}, {"$not_contains": "alpha"}]},
               {"$or": [{"$contains": "alpha"}, {"$contains": "gamma"}]}]:
        validate_where_document(wd)
        got = evaluate_where_document(wd, docs, 4)
        assert got.tolist() == [where_document_matches(wd, d) for d in docs], wd
    for bad in [{"$contains": ""}, {"$like": "x"}, {"$contains": 3}, "x"]:
        with pytest.raises(ValueError):
            validate_where_document(bad)


@pytest.mark.parametrize("seed", range(6))
def test_single_record_matcher_agrees_with_the_column_compiler(seed):
    """match_record (used to patch cached bitmaps row by row after a write) must decide every record exactly
    as MetadataColumns.evaluate decides it in bulk."""
    import random
    from local_rag_system_b200.where import MetadataColumns, match_record
    rng = random.Random(100 + seed)
    keys = ["namespace", "n", "score", "flag"]
    metas = []
    for _ in range(300):
        m = {}
        if rng.random() < 0.8:
            m["namespace"] = rng.choice(["history", "docs", "x"])
        if rng.random() < 0.7:
            m["n"] = rng.randint(0, 9)
        if rng.random() < 0.5:
            m["score"] = rng.choice([0.5, 1.5, 2.0])
        if rng.random() < 0.4:
            m["flag"] = rng.random() < 0.5
        if rng.random() < 0.1:
            m["n"] = "seven"              # another type under the same key
        metas.append(m or None)
    cols = MetadataColumns()
    for r, m in enumerate(metas):
        cols.set_row(r, None, m)
    wheres = [{"namespace": "history"}, {"n": {"$lt": 5}}, {"n": {"$gte": 3, "$lte": 7}}, {"score": {"$gt": 1.0}},
              {"flag": True}, {"flag": {"$ne": True}}, {"namespace": {"$in": ["docs", "x"]}}, {"n": {"$nin": [1, 2, 3]}},
              {"namespace": "history", "n": 3}, {"$or": [{"n": 1}, {"namespace": "x"}]},
              {"$and": [{"namespace": {"$ne": "x"}}, {"$or": [{"score": 2.0}, {"n": {"$gt": 6}}]}]}, {"missing": 1},
              {"missing": {"$ne": 1}}, {"n": "seven"}, {"n": 7}, {"score": {"$lt": 2}}]
    for w in wheres:
        bulk = cols.evaluate(w, len(metas))
        one = np.array([match_record(w, m) for m in metas])
        assert np.array_equal(bulk, one), w


def test_single_document_matcher():
    from local_rag_system_b200.where import evaluate_where_document, match_document
    docs = np.array(["alpha beta", None, "beta", "gamma"], dtype=object)
    for wd in ({"$contains": "beta"}, {"$not_contains": "beta"}, {"$or": [{"$contains": "alpha"}, {"$contains": "gamma"}]},
               {"$and": [{"$contains": "beta"}, {"$not_contains": "alpha"}]}):
        assert np.array_equal(evaluate_where_document(wd, docs, 4), np.array([match_document(wd, d) for d in docs])), wd
