#!/usr/bin/env python
"""Extract the reference's shipped vector index into a small golden fixture.

Run in the build container only (reads /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_fixture.py

Source: /root/reference/vector_store/chroma.sqlite3
  * table `embeddings_queue`  - the write-ahead log Chroma 0.5.3 replays on
    start-up: 37 UPSERT records (seq_id, id, fp32[384] blob, JSON metadata
    with the document under "chroma:document").  These are the INPUTS of the
    upsert path (reference call site scripts/build_index.py:89-96).
  * tables `embeddings`, `embedding_metadata`, `embedding_fulltext_search_content`
    - the metadata segment Chroma itself materialised from that log.  These
    are reference OUTPUTS of the upsert path (25 live ids, last write wins)
    and pin our upsert-replace semantics.

Writes (committed):
  tests/golden/gamefantasy_wal.npz        vectors fp32 [37,384] + seq ids
  tests/golden/gamefantasy_wal.json       ids / metadata / documents per WAL record
  tests/golden/gamefantasy_segment.json   Chroma's own materialised state (25 ids)
  tests/golden/known_answers.json         top-5 answers listed in SURVEY.md 8c,
                                          recomputed here with oracle/ and
                                          cross-checked against the SURVEY values
"""
import json
import os
import sqlite3
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_DB = "/root/reference/vector_store/chroma.sqlite3"


def main():
    con = sqlite3.connect(f"file:{REF_DB}?mode=ro&immutable=1", uri=True)
    coll = con.execute("select id, name, dimension from collections where name='gamefantasy'").fetchone()
    coll_id, name, dim = coll
    rows = con.execute(
        "select seq_id, operation, id, vector, encoding, metadata from embeddings_queue "
        "where topic like ? order by seq_id", (f"%{coll_id}",)).fetchall()
    vecs, recs = [], []
    for seq_id, op, eid, blob, enc, meta in rows:
        assert enc == "FLOAT32" and op == 2, (enc, op)
        v = np.frombuffer(blob, dtype="<f4")
        assert v.shape == (dim,)
        vecs.append(v)
        m = json.loads(meta)
        doc = m.pop("chroma:document", None)
        recs.append({"seq_id": seq_id, "operation": "UPSERT", "id": eid, "metadata": m, "document": doc})
    vecs = np.stack(vecs).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "gamefantasy_wal.npz"), vectors=vecs,
                        seq_ids=np.array([r["seq_id"] for r in recs], dtype=np.int64))
    with open(os.path.join(HERE, "gamefantasy_wal.json"), "w", encoding="utf-8") as f:
        json.dump({"collection": name, "dimension": dim, "space": "l2", "records": recs}, f,
                  ensure_ascii=False, indent=1)

    # Chroma's own materialised metadata segment (reference output of the upsert path)
    seg = {}
    for pk, eid, seq in con.execute("select id, embedding_id, seq_id from embeddings order by id"):
        md = {}
        for key, s, i, fl, b in con.execute(
                "select key, string_value, int_value, float_value, bool_value from embedding_metadata where id=?", (pk,)):
            val = s if s is not None else i if i is not None else fl if fl is not None else bool(b)
            md[key] = val
        doc = md.pop("chroma:document", None)
        seg[eid] = {"seq_id": int.from_bytes(seq, "big"), "metadata": md, "document": doc}
    with open(os.path.join(HERE, "gamefantasy_segment.json"), "w", encoding="utf-8") as f:
        json.dump({"count": len(seg), "max_seq_id": int.from_bytes(
            con.execute("select seq_id from max_seq_id").fetchone()[0], "big"), "ids": seg},
            f, ensure_ascii=False, indent=1)

    # Known answers recomputed with the oracle.
    from oracle.exact_search import OracleCollection
    oc = OracleCollection(space="l2")
    for r, v in zip(recs, vecs):
        oc.upsert([r["id"]], [v], [r["metadata"]], [r["document"]])
    assert oc.count() == 25
    survey = {
        "q1": [("fyp_core::summary", 0.0), ("fyp_core::key_elements", 0.2736222),
               ("fyp_core::design_concept", 0.4299542), ("media_fyp_launched::summary", 0.4920169),
               ("media_fyp_launched::claims", 0.7895305)],
        "q2": [(None, 0.9466366), (None, 1.0657333), (None, 1.1426963), (None, 1.1821779), (None, 1.1928142)],
        "q3": [("media_fyp_launched::claims", 0.0), ("media_fyp_launched::summary", 0.4696369),
               (None, 0.7240932), ("fyp_core::summary", 0.7895305),
               ("fyp_core_background::design details", 0.8836023)],
    }
    cases = {
        "q1": {"query_id": "fyp_core::summary", "k": 5, "where": None},
        "q2": {"query_id": "fyp_core::summary", "k": 5, "where": {"namespace": "history"}},
        "q3": {"query_id": "media_fyp_launched::claims", "k": 5, "where": None},
    }
    out = {}
    for name_, c in cases.items():
        q = oc.vector_of(c["query_id"])
        res = oc.query([q], n_results=c["k"], where=c["where"])
        ids, d = res["ids"][0], res["distances"][0]
        for (sid, sd), gid, gd in zip(survey[name_], ids, d):
            assert abs(sd - gd) < 5e-7, (name_, sd, gd)
            assert sid is None or sid == gid, (name_, sid, gid)
        out[name_] = dict(c, ids=ids, distances=[float(x) for x in d])
    with open(os.path.join(HERE, "known_answers.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=1)
    print("fixture written:", vecs.shape, len(seg), "live ids")


if __name__ == "__main__":
    main()
