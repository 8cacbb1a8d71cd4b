"""The north_star's bar for the bf16 path: recall@k >= 0.999 against EXACT fp32 brute force on the
UN-ROUNDED inputs (>= 1000 unit-norm queries, >= 1M rows) -- i.e. what the reference's fp32 index would
return -- not merely exactness on inputs that were rounded to bf16 first (tests/test_gpu_parity.py does
that for the scan kernels).  Ranking on bf16 rows alone cannot meet it (~0.993 on 1M unit-norm rows: the
10th and 11th neighbours are ~8e-4 apart, bf16 rounding moves a dot product by ~1e-4); the store therefore
keeps the un-rounded fp32 rows next to the bf16 ones, lets the HBM-bound scan rank bf16, and re-ranks the
k + slack best exactly.  Distances returned are then fp32-exact as well.  GPU only."""
import numpy as np
import pytest

from local_rag_system_b200 import DeviceStore
from oracle.exact_search import exact_search, fast_topk_f32, normalise_rows
from tests.conftest import unit_rows

pytestmark = pytest.mark.gpu


def _queries(x, nq, seed):
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((nq, x.shape[1]), dtype=np.float32)
    planted = rng.choice(x.shape[0], nq // 10, replace=False)       # 10 % planted next to a corpus row
    q[: nq // 10] = x[planted] + 0.05 * rng.standard_normal((nq // 10, x.shape[1]), dtype=np.float32) / np.sqrt(x.shape[1])
    return normalise_rows(q), planted


def _recall(got_rows, want_rows, k):
    return float(np.mean([len(set(got_rows[b, :k].tolist()) & set(want_rows[b, :k].tolist())) / k
                          for b in range(got_rows.shape[0])]))


def test_bf16_store_recall_vs_fp32_oracle_on_unrounded_inputs():
    n, dim, nq, k = 1_000_000, 384, 1024, 10
    x = unit_rows(n, dim, 2024)                       # fp32, NOT representable in bf16
    q, planted = _queries(x, nq, 7)
    want_r, want_d = fast_topk_f32("cosine", q, x, k)              # exact fp32 search on the un-rounded inputs
    st = DeviceStore(dim, "bf16", "cosine", capacity_hint=n, rerank=True)
    plain = DeviceStore(dim, "bf16", "cosine", capacity_hint=n, rerank=False)
    try:
        assert st.rerank and not plain.rerank
        st.upsert(x)
        plain.upsert(x)
        # tensor regime (one batch of 1024) and stream regime (batches of 2)
        rows_t, d_t, c_t = st.query(q, k)
        assert st.last_query_info()["regime"] == "tensor" and np.all(c_t == k)
        rec_t = _recall(rows_t, want_r, k)
        rows_s = np.concatenate([st.query(q[i:i + 2], k, regime="stream")[0] for i in range(0, 128, 2)])
        rec_s = _recall(rows_s, want_r[:128], k)
        rows_p, _, _ = plain.query(q, k)
        rec_p = _recall(rows_p, want_r, k)
        print(f"recall@{k} vs exact fp32 on un-rounded inputs: bf16+rerank tensor {rec_t:.5f} stream {rec_s:.5f}; "
              f"bf16 alone {rec_p:.5f}")
        assert rec_t >= 0.999 and rec_s >= 0.999
        assert 0.97 <= rec_p < rec_t                    # what the fp32 plane buys
        # distances are exact fp32 (not bf16-rounded): same tolerance as an fp32 store
        hit = rows_t == want_r
        assert np.allclose(d_t[hit], want_d[hit], rtol=1e-5, atol=2e-6)
        assert np.array_equal(rows_t[: nq // 10, 0], planted)
        # embeddings come back un-rounded
        back = st.fetch(np.arange(5), exact=True)
        assert np.allclose(back, x[:5], atol=2e-7) and not np.allclose(st.fetch(np.arange(5)), x[:5], atol=2e-7)
    finally:
        st.close()
        plain.close()


@pytest.mark.parametrize("space,k,B", [("l2", 100, 64), ("ip", 40, 3), ("cosine", 17, 1), ("l2", 5, 8)])
def test_rerank_all_spaces_and_large_k(space, k, B):
    """Both regimes, every space, k beyond the register lists, filter + tombstones: rows must be the
    exact-fp32 top-k (ties within 1e-6 aside) and distances fp32-exact."""
    n, dim = 60_000, 256
    rng = np.random.default_rng(k)
    x = unit_rows(n, dim, 5 + k) * (1.0 if space == "cosine" else rng.uniform(0.8, 1.2, (n, 1)).astype(np.float32))
    q = unit_rows(B, dim, 6 + k)
    q[0] = x[123] * 1.0001
    dead = rng.choice(n, 3000, replace=False)
    passing = rng.random(n) < 0.5
    st = DeviceStore(dim, "bf16", space, rerank=True)
    try:
        st.upsert(x)
        st.delete(dead)
        st.set_mask(0, passing)
        valid = passing.copy()
        valid[dead] = False
        want_r, want_d = exact_search(space, q, x, k, valid, "f32")
        for regime in ("stream", "tensor"):
            rows, d, c = st.query(q, k, mask_slot=0, regime=regime)
            assert np.all(c == k) and st.last_query_info()["regime"] == regime
            assert np.all(valid[rows])
            rec = _recall(rows, np.stack(want_r), k)
            assert rec >= 0.999, (regime, rec)
            for b in range(B):
                assert np.all(np.diff(d[b]) >= 0)
                assert np.allclose(np.sort(d[b]), np.sort(want_d[b]), rtol=1e-5, atol=2e-6), (regime, b)
    finally:
        st.close()


def test_rerank_store_growth_and_in_place_upsert():
    """The fp32 plane follows growth (realloc + copy) and in-place upserts like the bf16 rows do."""
    dim, k = 128, 10
    x = unit_rows(5000, dim, 77)
    st = DeviceStore(dim, "bf16", "cosine", capacity_hint=1024, rerank=True)
    try:
        for s in range(0, 5000, 700):                 # grows several times
            st.upsert(x[s:s + 700])
        y = unit_rows(50, dim, 78)
        st.upsert(y, rows=np.arange(100, 150))        # replace rows 100..149 in place
        x2 = x.copy()
        x2[100:150] = y
        assert np.allclose(st.fetch(np.arange(5000), exact=True), normalise_rows(x2), atol=2e-7)
        q = unit_rows(4, dim, 79)
        q[0] = y[3]
        want_r, want_d = exact_search("cosine", q, x2, k, None, "f32")
        rows, d, c = st.query(q, k)
        assert rows[0, 0] == 103
        assert _recall(rows, np.stack(want_r), k) >= 0.999
        assert np.allclose(d, np.stack(want_d), rtol=1e-5, atol=2e-6)
    finally:
        st.close()
