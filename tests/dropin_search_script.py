"""Run by tests/test_dropin_reference.py in a subprocess (build container only).

Imports the UNMODIFIED reference app (/root/reference/api/app.py) with the
`chromadb` shim on PYTHONPATH, pointed at the reference's own shipped
vector_store/ (imported read-only by persist.py), and drives POST /search,
GET /health and the ingest/delete helpers.  The device store is the oracle-
backed test double (no GPU in the build container); the embedding model is
replaced by a lookup that returns a stored vector (no MiniLM weights offline).
Prints one JSON document with what it observed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "shim"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402

import chromadb  # noqa: E402  (the shim)
from chromadb.utils import embedding_functions  # noqa: E402
from local_rag_system_b200 import collection as colmod  # noqa: E402
from tests.fake_store import FakeDeviceStore  # noqa: E402

colmod.DeviceStore = FakeDeviceStore

wal = json.load(open(os.path.join(ROOT, "tests/golden/gamefantasy_wal.json"), encoding="utf-8"))
vecs = np.load(os.path.join(ROOT, "tests/golden/gamefantasy_wal.npz"))["vectors"]
by_id = {r["id"]: v for r, v in zip(wal["records"], vecs)}
calls = []


class LookupEF:
    """query text = a stored id -> that record's vector; anything else -> a fixed new vector"""

    def __init__(self, model_name=None):
        pass

    def __call__(self, texts):
        out = []
        for t in texts:
            v = by_id.get(t)
            out.append((v if v is not None else np.full(384, 1 / np.sqrt(384), np.float32)).tolist())
        return out


embedding_functions.SentenceTransformerEmbeddingFunction = LookupEF
orig_query = colmod.Collection.query


def spy(self, *a, **kw):
    calls.append({k: v for k, v in kw.items()})
    return orig_query(self, *a, **kw)


colmod.Collection.query = spy

from fastapi.testclient import TestClient  # noqa: E402
import api.app as app_module  # noqa: E402

client = TestClient(app_module.app)
H = {"x-api-key": os.environ.get("API_KEY", "")}
out = {}
out["health"] = client.get("/health", headers=H).json()
r = client.post("/search", json={"query": "fyp_core::summary", "k": 5}, headers=H)
out["search_status"] = r.status_code
out["search"] = r.json()
r = client.post("/search", json={"query": "fyp_core::summary", "k": 5, "namespace": "history", "canonicality": "non"}, headers=H)
out["search_filtered"] = r.json()
out["query_kwargs"] = calls
# ingest + delete helpers (api/app.py:209-225, 284-315) through the shim
ok = app_module._chroma_add("doc-new", "brand new text", {"source_key": "sk-new", "title": "T", "updated_ts": 1})
out["chroma_add_ok"] = ok
out["count_after_add"] = app_module.collection.count()
app_module._delete_doc_from_stores("doc-new", "sk-new")
out["count_after_delete"] = app_module.collection.count()
bad = app_module._chroma_add("doc-bad", "text", {"nested": {"a": 1}})      # non-scalar metadata -> swallowed -> False
out["chroma_add_bad"] = bad
print("RESULT " + json.dumps(out, ensure_ascii=False))
