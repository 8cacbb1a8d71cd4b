"""The C-ABI library loads and exports every symbol include/rag_b200.h declares.
CPU only: no compute calls are made without a GPU."""
import ctypes
import os
import re

import pytest

from local_rag_system_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rag_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rag_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(_native.LIB_PATH), "run `python local-rag-system_b200/build.py`"
    assert os.path.dirname(_native.LIB_PATH).endswith("local-rag-system_b200")


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rag_b200.h but not exported"
    assert sorted(_native.SIGNATURES) == names, "ctypes signature table out of step with the header"


def test_only_the_abi_is_exported():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert exported and all(s.startswith("rag_") for s in exported), exported


def test_host_side_key_helpers():
    lib = _native.load()
    assert lib.rag_abi_version() == _native.ABI_VERSION == 3
    vals = [-2.5, -0.0, 0.0, 2.0 ** -100, 0.25, 1.0, 3.5, float("inf")]
    keys = [lib.rag_key_pack(v, 7) for v in vals]
    assert keys == sorted(keys)                      # ordered like the floats
    for v, k in zip(vals, keys):
        assert lib.rag_key_dist(k) == v and lib.rag_key_row(k) == 7
    assert lib.rag_key_pack(0.5, 3) < lib.rag_key_pack(0.5, 4) < _native.EMPTY_KEY   # ties -> lower row first


def test_no_gpu_means_a_loud_error_not_a_fallback():
    lib = _native.load()
    if lib.rag_device_count() > 0:
        pytest.skip("a GPU is present")
    from local_rag_system_b200 import DeviceStore
    with pytest.raises(_native.EngineError, match="no CPU fallback"):
        DeviceStore(8)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "local-rag-system_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in text.replace("# oracle", "").replace("the oracle", "").replace("oracle's", "") \
                    or "import oracle" not in text and "from oracle" not in text, f
                assert "from oracle" not in text and "import oracle" not in text, f
