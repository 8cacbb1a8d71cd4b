import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        from local_rag_system_b200 import _native
        return _native.load().rag_device_count() > 0
    except Exception:
        return False


HAS_GPU = None


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip: the product has no fallback
    return


@pytest.fixture(scope="session")
def golden():
    import json
    wal = json.load(open(os.path.join(GOLDEN, "gamefantasy_wal.json"), encoding="utf-8"))
    vec = np.load(os.path.join(GOLDEN, "gamefantasy_wal.npz"))["vectors"]
    seg = json.load(open(os.path.join(GOLDEN, "gamefantasy_segment.json"), encoding="utf-8"))
    known = json.load(open(os.path.join(GOLDEN, "known_answers.json"), encoding="utf-8"))
    return {"wal": wal, "vectors": vec, "segment": seg, "known": known}


def unit_rows(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)
