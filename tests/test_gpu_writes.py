"""Write path of the device store: parked small writes (one copy + one launch per burst instead of per
call), stream-ordered writes without device-wide synchronisation, the vectorised upsert kernel on every
row layout, incremental `where` bitmaps.  GPU only."""
import numpy as np
import pytest

import local_rag_system_b200 as rag
from local_rag_system_b200 import DeviceStore
from oracle.exact_search import exact_search, prepare_corpus
from tests.conftest import unit_rows

pytestmark = pytest.mark.gpu


def test_small_writes_are_coalesced_and_visible_to_the_next_read():
    """The reference's pattern: one col.add per document (api/app.py:209-225), 1-5 rows per col.upsert
    (scripts/build_index.py:89-96), then a search."""
    dim = 384
    x = unit_rows(600, dim, 1)
    st = DeviceStore(dim, "f32", "l2")
    try:
        base = st.kernel_launches()
        for i in range(200):                            # 200 single-row adds
            assert st.upsert(x[i:i + 1]).tolist() == [i]
        assert st.count() == 200 and st.rows() == 200   # bookkeeping is immediate
        assert st.kernel_launches() == base             # ... the device has not been touched yet
        rows, d, c = st.query(x[150:151], 3)            # the read flushes: ONE upsert launch + the search
        assert rows[0, 0] == 150 and d[0, 0] < 1e-10
        assert st.kernel_launches() - base <= 2
        # more than the 256-row block between two reads: flushed in blocks, nothing lost
        for i in range(200, 600, 4):
            st.upsert(x[i:i + 4])
        # overwrite a parked row before it ever reached the device, and delete another parked row
        st.upsert(x[0:1] * 3.0, rows=np.array([599]))
        st.delete([598])
        assert st.count() == 599
        rows, d, c = st.query(np.concatenate([x[0:1] * 3.0, x[598:599]]), 1)
        assert rows[0, 0] == 599 and d[0, 0] < 1e-9 and rows[1, 0] != 598
        want_r, want_d = exact_search("l2", x[300:305], np.concatenate([x[:598], x[598:599] * 0 + 9, x[0:1] * 3.0]), 5,
                                      np.arange(600) != 598)
        rows, d, c = st.query(x[300:305], 5)
        assert np.array_equal(rows, np.stack(want_r)) and np.allclose(d, np.stack(want_d), rtol=1e-5, atol=1e-6)
        st.upsert(x[:3], rows=np.array([10, 11, 12]))
        st.flush()
        assert np.allclose(st.fetch([10, 11, 12]), x[:3], atol=0)
    finally:
        st.close()


@pytest.mark.parametrize("dim", [3, 17, 96, 100, 384, 768, 1000, 1024, 1536, 2048, 2052, 4100])
@pytest.mark.parametrize("dtype,space", [("f32", "cosine"), ("bf16", "cosine"), ("bf16", "l2")])
def test_upsert_kernel_all_row_layouts(dim, dtype, space):
    """K1 against the oracle's prepare_corpus on every code path: vector path (dim % 4 == 0, 1-8 groups of
    8 elements per lane), the generic path (odd dims, dim > 2048), scattered destination rows, both planes."""
    n = 70
    rng = np.random.default_rng(dim)
    x = (rng.standard_normal((n, dim)) * rng.uniform(0.1, 5.0, (n, 1))).astype(np.float32)
    x[3] = 0.0                                        # a zero row stays zero (cosine: scale 0, not NaN)
    st = DeviceStore(dim, dtype, space, rerank=True)
    try:
        st.upsert(x[:40])
        st.upsert(x[40:], rows=None)
        st.upsert(x[:5], rows=np.array([60, 2, 33, 7, 41]))       # scattered, in place
        want = x.copy()
        want[[60, 2, 33, 7, 41]] = x[:5]
        stored = st.fetch(np.arange(n))
        ref = prepare_corpus(space, want, dtype)
        if dtype == "f32":
            assert np.allclose(stored, ref, rtol=0, atol=3e-7 * max(1.0, float(np.abs(ref).max())))
        else:
            # normalisation scale may differ by an ulp from numpy's: a bf16 rounding boundary can flip
            bad = stored != ref
            assert bad.mean() < 2e-3 and np.allclose(stored, ref, rtol=2 ** -7, atol=1e-30)
            exact = st.fetch(np.arange(n), exact=True)
            assert np.allclose(exact, prepare_corpus(space, want, "f32"), rtol=0, atol=3e-7 * max(1.0, float(np.abs(want).max())))
        assert not np.isnan(stored).any() and np.all(stored[3] == 0)
        rows, d, c = st.query(want[10:12], 3)
        assert rows[0, 0] == 10 and rows[1, 0] == 11
    finally:
        st.close()


def test_masks_stay_valid_across_writes_without_reupload():
    """Collection: a cached `where` bitmap is patched by the writes that follow, not re-evaluated and
    re-uploaded (api/app.py interleaves col.add with filtered /search)."""
    client = rag.EphemeralClient()
    try:
        col = client.get_or_create_collection("m", metadata={"hnsw:space": "l2"})
        dim = 16
        rng = np.random.default_rng(0)
        x = rng.standard_normal((3000, dim)).astype(np.float32)
        col.add(ids=[f"a{j}" for j in range(1200)], embeddings=x[:1200],
                metadatas=[{"namespace": "history" if j % 3 else "docs", "n": j} for j in range(1200)])
        w = {"namespace": "history"}
        st = col._s
        res = col.query(query_embeddings=x[5:6], n_results=4, where=w)
        assert st.mask_uploads == 1
        for j in range(2000, 2100):                     # interleaved single adds and filtered searches (ids a2000..)
            col.add(ids=[f"a{j}"], embeddings=x[j:j + 1], metadatas=[{"namespace": "history" if j % 2 else "docs", "n": j}])
            res = col.query(query_embeddings=x[j:j + 1], n_results=1, where=w)
            assert (res["ids"][0] == [f"a{j}"]) == bool(j % 2)
        assert st.mask_uploads == 1 and st.mask_patches >= 100
        col.upsert(ids=["a2001"], embeddings=x[2001:2002], metadatas=[{"namespace": "docs"}])      # leaves the filter
        assert col.query(query_embeddings=x[2001:2002], n_results=1, where=w)["ids"][0] != ["a2001"]
        col.update(ids=["a2001"], metadatas=[{"namespace": "history"}])                             # and comes back
        assert col.query(query_embeddings=x[2001:2002], n_results=1, where=w)["ids"][0] == ["a2001"]
        col.delete(ids=["a2001"])
        col.add(ids=["zz"], embeddings=x[2001:2002], metadatas=[{"namespace": "docs"}])             # reuses the freed row
        assert col.query(query_embeddings=x[2001:2002], n_results=1, where=w)["ids"][0] != ["zz"]
        assert st.mask_uploads == 1
        # the whole filtered result still equals a fresh evaluation
        fresh = col.get(where=w)["ids"]
        res = col.query(query_embeddings=x[:1], n_results=len(fresh) + 50, where=w)
        assert sorted(res["ids"][0]) == sorted(fresh)
        # upsert merges metadata keys of an existing id and keeps its document when none is passed
        col.upsert(ids=["a7"], embeddings=x[7:8], metadatas=[{"extra": 1}])
        got = col.get(ids=["a7"])
        assert got["metadatas"][0] == {"namespace": "history", "n": 7, "extra": 1}
    finally:
        client.reset()


def test_async_reader_is_ordered_against_writes():
    """rag_store_query_dev on a caller stream, writes in between: every search sees exactly the writes that
    returned before it was launched (no device-wide synchronisation on the write path)."""
    torch = pytest.importorskip("torch")
    dim, k = 64, 1
    x = unit_rows(4000, dim, 5)
    st = DeviceStore(dim, "f32", "l2")
    try:
        st.upsert(x[:2000])
        stream = torch.cuda.Stream()
        q = torch.from_numpy(x[2000:2064].copy()).cuda()
        outs = []
        with torch.cuda.stream(stream):
            for i in range(64):
                rows = torch.empty((1, k), dtype=torch.int64, device="cuda")
                st.query_device(q[i:i + 1].data_ptr(), 1, k, 0, stream=stream.cuda_stream, out_rows_ptr=rows.data_ptr())
                outs.append(rows)                                     # launched BEFORE row 2000 + i exists
                st.upsert(x[2000 + i:2001 + i])                       # parked
                rows2 = torch.empty((1, k), dtype=torch.int64, device="cuda")
                st.query_device(q[i:i + 1].data_ptr(), 1, k, 0, stream=stream.cuda_stream, out_rows_ptr=rows2.data_ptr())
                outs.append(rows2)                                    # launched AFTER: must find it
        stream.synchronize()
        for i in range(64):
            assert int(outs[2 * i][0, 0]) != 2000 + i
            assert int(outs[2 * i + 1][0, 0]) == 2000 + i
    finally:
        st.close()


def test_queries_in_flight_submit_collect():
    """rag_store_query_submit / _wait: up to 4 host-buffer queries in the air, results identical to the
    blocking call, in both regimes, with a filter; a fifth submit is refused; writes in between are seen."""
    dim, k = 192, 7
    x = unit_rows(30_000, dim, 8)
    st = DeviceStore(dim, "f32", "cosine")
    try:
        st.upsert(x)
        st.set_mask(2, (np.arange(30_000) % 3) == 0)
        qs = [unit_rows(B, dim, 50 + i) for i, B in enumerate((1, 2, 1, 40, 3, 1, 8, 1))]
        for slot in (-1, 2):
            want = [st.query(q, k, mask_slot=slot) for q in qs]
            tickets, got = [], []
            for q in qs:
                tickets.append(st.submit(q, k, mask_slot=slot))
                if len(tickets) == 4:
                    with pytest.raises(ValueError):
                        st.submit(q, k)                       # a fifth query in flight is refused
                    got.append(st.collect(tickets.pop(0)))
            while tickets:
                got.append(st.collect(tickets.pop(0)))
            for g, w in zip(got, want):
                assert all(np.array_equal(a, b) for a, b in zip(g, w))
        t = st.submit(x[123], 1)
        st.upsert(x[123:124] * 0.5, rows=np.array([123]))      # parked write AFTER the submit: not seen by it
        assert st.collect(t)[0][0, 0] == 123
        y = unit_rows(1, dim, 999)
        r = st.upsert(y)
        t = st.submit(y[0], 1)                                  # submit flushes the parked write first
        assert st.collect(t)[0][0, 0] == r[0]
        with pytest.raises(ValueError):
            st.collect((0, 1, 1))                               # nothing in flight on that ticket
    finally:
        st.close()
