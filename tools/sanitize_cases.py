"""A few SMALL cases through every kernel of the engine, for compute-sanitizer
(`compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_cases.py`): upsert (vector and
generic path, both planes), scan (dense, tombstones, filtered walk, exact re-ranking, several query groups),
tensor regime (single CTA and cta_group::2 pairs, k <= 16 and the heap lists, l2 refine, fp32 split + guard),
merge kernels, mask patching, the sharded store's gather path.  Every result is checked against the oracle, so
a sanitizer run is also a parity run.  One device; the cross-GPU flag protocol is covered by
tests/multi_gpu_check.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import local_rag_system_b200 as rag  # noqa: E402
from oracle.exact_search import exact_search  # noqa: E402
from tests.conftest import unit_rows  # noqa: E402


def check(tag, got, want_r, k):
    rows = got[0]
    rec = np.mean([len(set(rows[b, :k].tolist()) & set(want_r[b].tolist())) / k for b in range(rows.shape[0])])
    assert rec >= 0.999, (tag, rec)
    print(f"ok {tag} recall {rec:.4f}", flush=True)


def main():
    n = int(os.environ.get("SANITIZE_ROWS", "6000"))
    for dim, dtype, space in ((384, "bf16", "cosine"), (100, "f32", "l2"), (768, "bf16", "l2"), (192, "f32", "cosine")):
        x = unit_rows(n, dim, dim)
        st = rag.DeviceStore(dim, dtype, space, rerank=True)
        st.upsert(x[: n // 2])
        for s in range(n // 2, n, 37):
            st.upsert(x[s:s + 37])                                   # parked writes
        dead = np.arange(5, n, 17)
        st.delete(dead)
        passing = (np.arange(n) % 9) < 2
        st.set_mask(0, passing)
        st.patch_mask(0, np.array([3, 4, n - 1]), np.array([1, 0, 1], np.uint8))
        passing[[3, 4, n - 1]] = [True, False, True]
        valid = np.ones(n, bool)
        valid[dead] = False
        for B, k in ((1, 10), (11, 5), (40, 10), (150, 20)):
            q = unit_rows(B, dim, B)
            for regime in ("stream", "tensor"):
                if regime == "tensor" and (dim % 8 or (dtype == "f32" and dim % 16)):
                    continue
                for slot, v in ((-1, valid), (0, valid & passing)):
                    want_r, _ = exact_search(space, q, x, k, v, "f32")
                    check(f"{dim}/{dtype}/{space} B={B} k={k} {regime} slot={slot}", st.query(q, k, mask_slot=slot, regime=regime), want_r, k)
        st.close()
    x = unit_rows(5000, 64, 1)
    sh = rag.ShardedDeviceStore(64, "f32", "ip", devices=[0, 0, 0])
    sh.upsert(x)
    q = unit_rows(7, 64, 2)
    want_r, _ = exact_search("ip", q, x, 10, None, "f32")
    check("sharded gather", sh.query(q, 10), want_r, 10)
    sh.close()
    print("SANITIZE_CASES OK", flush=True)


if __name__ == "__main__":
    main()
