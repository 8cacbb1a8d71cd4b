"""Run by tests/test_gpu_collection.py::test_shim_import_on_device in a subprocess on the GPU box, with
`shim/` first on PYTHONPATH: `import chromadb` resolves to the shim, and the reference's call shapes --
api/app.py:87-91 (client / collection construction), scripts/build_index.py:92-96 (upsert, 1-5 chunks per call),
api/app.py:544-549 (query_texts / n_results / where / include), api/app.py:221 (single add), 269 / 306 / 311
(delete by where / ids), api/routes/system.py:33 (count), scripts/query_local.py:33 (include "uris") -- run on the
REAL engine (no test double, no /root/reference needed) against the golden fixture's known answers."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import chromadb  # noqa: E402  (the shim: PYTHONPATH puts shim/ first)
from chromadb.utils import embedding_functions  # noqa: E402

assert os.path.normpath(chromadb.__file__).startswith(os.path.join(ROOT, "shim")), chromadb.__file__
g = os.path.join(ROOT, "tests", "golden")
wal = json.load(open(os.path.join(g, "gamefantasy_wal.json"), encoding="utf-8"))
vecs = np.load(os.path.join(g, "gamefantasy_wal.npz"))["vectors"]
known = json.load(open(os.path.join(g, "known_answers.json"), encoding="utf-8"))
by_id = {r["id"]: v for r, v in zip(wal["records"], vecs)}
by_doc = {r["document"]: v for r, v in zip(wal["records"], vecs)}


class LookupEF:
    """stands in for SentenceTransformerEmbeddingFunction (no MiniLM weights offline): a text that is a stored
    id or a stored document embeds to that record's vector"""

    def __init__(self, model_name=None):
        self.model_name = model_name

    def __call__(self, texts):
        return [(by_id.get(t, by_doc.get(t, vecs[0] * 0 + 0.05))).tolist() for t in texts]


embedding_functions.SentenceTransformerEmbeddingFunction = LookupEF      # as tests/test_kb_crud.py:62-66 patches it
persist = sys.argv[1]
client = chromadb.PersistentClient(path=persist)                          # api/app.py:89
embedder = embedding_functions.SentenceTransformerEmbeddingFunction(model_name="all-MiniLM-L6-v2")
collection = client.get_or_create_collection(name="gamefantasy", embedding_function=embedder)      # api/app.py:91
recs = wal["records"]
for s in range(0, len(recs), 3):                                           # scripts/build_index.py:92-96
    chunk = recs[s:s + 3]
    if len({r["id"] for r in chunk}) != len(chunk):
        for r in chunk:
            collection.upsert(ids=[r["id"]], documents=[r["document"]], metadatas=[r["metadata"]])
    else:
        collection.upsert(ids=[r["id"] for r in chunk], documents=[r["document"] for r in chunk],
                          metadatas=[r["metadata"] for r in chunk])
out = {"count": collection.count()}                                        # api/routes/system.py:33
ans = {}
for name, case in known.items():
    res = collection.query(query_texts=[case["query_id"]], n_results=max(1, min(case["k"], 20)), where=case["where"],
                           include=["documents", "metadatas", "distances"])          # api/app.py:544-549
    ans[name] = {"ids": res["ids"][0], "distances": res["distances"][0],
                 "ok": res["ids"][0] == case["ids"] and bool(np.allclose(res["distances"][0], case["distances"], rtol=1e-5, atol=1e-6))}
out["known"] = ans
res = collection.query(query_texts=["fyp_core::summary"], n_results=3, include=["documents", "metadatas", "distances", "uris"])
out["uris_none"] = res["uris"] is None and len(res["ids"][0]) == 3        # scripts/query_local.py:33
# a second client on the same path, no embedding function (api/app.py:267-269): same state; delete by where
col2 = chromadb.PersistentClient(path=persist).get_or_create_collection(name="gamefantasy")
collection.add(ids=["new-doc"], documents=["fyp_core::summary"], metadatas=[{"source_key": "k-new", "namespace": "docs"}])  # api/app.py:221
out["count_after_add"] = col2.count()
col2.delete(where={"source_key": "k-new"})                                 # api/app.py:269 / 311
out["count_after_delete_where"] = collection.count()
col2.delete(ids=["fyp_core::summary"])                                     # api/app.py:306
out["count_after_delete_ids"] = collection.count()
store = collection.device_store
out["engine"] = type(store).__name__
out["kernel_launches"] = store.kernel_launches()
out["regime"] = store.last_query_info()["regime"]
print("RESULT " + json.dumps(out))
