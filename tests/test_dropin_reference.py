"""Drop-in proof at the reference's own boundary, build container
only: /root/reference does not exist on the GPU box, so these are CPU tests and
skip themselves when the reference checkout is absent."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "api")), reason="reference checkout not present")


def _env(tmp_path, persist):
    env = dict(os.environ)
    env.update({
        "PYTHONPATH": os.pathsep.join([os.path.join(ROOT, "shim"), ROOT]),
        "PYTHONDONTWRITEBYTECODE": "1",
        "PERSIST_DIR": persist, "DOCS_DIR": str(tmp_path / "docs"),
        "KB_DB_PATH": str(tmp_path / "kb.sqlite"), "CONV_DB_PATH": str(tmp_path / "conv" / "conv.sqlite"),
        "API_KEY": "testkey",
    })
    return env


def test_reference_test_suite_passes_unmodified_on_the_shim(tmp_path):
    """pytest /root/reference/tests with `chromadb` resolved to shim/chromadb."""
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(REF, "tests"), "-q", "-p", "no:cacheprovider",
                        "--rootdir", str(tmp_path)], cwd=str(tmp_path),
                       env=_env(tmp_path, str(tmp_path / "persist")), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "3 passed" in r.stdout, r.stdout[-2000:]


def test_search_route_serves_the_shipped_index(tmp_path, golden):
    """The unmodified FastAPI app, PERSIST_DIR = a copy of the reference's vector_store/:
    /health counts 25, /search returns the known answers with the reference's
    hit shape, the engine receives exactly the kwargs of api/app.py:544-549."""
    # a private copy of the shipped index: the reference tree is read-only for this repo, and the journal
    # (rag_b200.sqlite3) that PersistentClient creates next to chroma.sqlite3 must not outlive the test --
    # otherwise later runs would find the collection "known" and never exercise the Chroma import again
    import shutil
    store = tmp_path / "vector_store"
    shutil.copytree(os.path.join(REF, "vector_store"), store,
                    ignore=shutil.ignore_patterns("rag_b200.sqlite3*"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_search_script.py")], cwd=str(tmp_path),
                       env=_env(tmp_path, str(store)), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    assert out["health"]["docs_count"] == 25 and out["health"]["chroma_ok"] is True
    assert out["search_status"] == 200
    hits = [h for h in out["search"]["hits"] if h.get("id") in golden["known"]["q1"]["ids"]]
    assert [h["id"] for h in hits] == golden["known"]["q1"]["ids"]
    assert np.allclose([h["score"] for h in hits], golden["known"]["q1"]["distances"], atol=1e-6)
    assert [h["rank"] for h in hits] == [1, 2, 3, 4, 5]
    assert all(set(h) >= {"rank", "id", "score", "metadata", "text"} for h in hits)
    f = [h for h in out["search_filtered"]["hits"] if h.get("id") in golden["known"]["q2"]["ids"]]
    assert [h["id"] for h in f] == golden["known"]["q2"]["ids"]
    kw = out["query_kwargs"]
    assert kw[0] == {"query_texts": ["fyp_core::summary"], "n_results": 5, "where": None,
                     "include": ["documents", "metadatas", "distances"]}
    assert kw[1]["where"] == {"namespace": "history", "canonicality": "non"}
    assert out["chroma_add_ok"] is True and out["count_after_add"] == 26 and out["count_after_delete"] == 25
    assert out["chroma_add_bad"] is False
    assert os.path.exists(store / "rag_b200.sqlite3")       # the import was carried into our own journal
