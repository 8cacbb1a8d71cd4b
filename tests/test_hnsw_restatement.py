"""The CPU HNSW comparator (oracle/hnsw_restatement.c): pinned to the
parameters of the reference's shipped index header and sanity-checked for recall."""
import struct

import numpy as np

from oracle.exact_search import exact_search
from oracle.hnsw import CHROMA_DEFAULTS, HnswIndex
from tests.conftest import unit_rows

# values read from /root/reference/vector_store/70ef2421-.../header.bin during the survey
HEADER_BIN = {"max_elements": 1000, "size_data_per_element": 1676, "label_offset": 1668, "offsetData": 132,
              "maxM": 16, "maxM0": 32, "M": 16, "mult": 0.360674, "ef_construction": 100}


def test_parameters_match_the_shipped_index_header():
    assert CHROMA_DEFAULTS["M"] == HEADER_BIN["M"] == HEADER_BIN["maxM"]
    assert 2 * CHROMA_DEFAULTS["M"] == HEADER_BIN["maxM0"]
    assert CHROMA_DEFAULTS["ef_construction"] == HEADER_BIN["ef_construction"]
    idx = HnswIndex(unit_rows(200, 16, 0))
    assert abs(idx.mult - HEADER_BIN["mult"]) < 1e-6            # 1 / ln(16)
    # hnswlib level-0 element layout: (maxM0 + 1) 4-byte link slots + vector + 8-byte label
    dim = 384
    assert (HEADER_BIN["maxM0"] + 1) * 4 == HEADER_BIN["offsetData"]
    assert HEADER_BIN["offsetData"] + dim * 4 == HEADER_BIN["label_offset"]
    assert HEADER_BIN["label_offset"] + struct.calcsize("q") == HEADER_BIN["size_data_per_element"]
    idx.close()


def test_recall_and_ordering():
    x = unit_rows(20000, 48, 1)
    q = unit_rows(200, 48, 2)
    idx = HnswIndex(x)
    want, _ = exact_search("l2", q, x, 10)
    ids_hi, d_hi = idx.query(q, k=10, ef=200)
    ids_lo, _ = idx.query(q, k=10, ef=10)            # Chroma's default search_ef
    rec = lambda ids: np.mean([len(set(ids[i]) & set(want[i])) / 10 for i in range(len(q))])
    assert rec(ids_hi) >= 0.97, rec(ids_hi)
    assert 0.3 <= rec(ids_lo) <= rec(ids_hi)
    assert np.all(np.diff(d_hi, axis=1) >= 0)
    # distances are the squared-l2 of the returned ids
    assert np.allclose(d_hi[0], np.sum((x[ids_hi[0]] - q[0]) ** 2, axis=1), atol=1e-5)
    # a stored vector finds itself
    ids, d = idx.query(x[123], k=1, ef=50)
    assert ids[0, 0] == 123 and d[0, 0] < 1e-6
    idx.close()
