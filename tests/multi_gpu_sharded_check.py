"""Run as ONE process on a box with N >= 2 GPUs (gpurun --gpus N -- python tests/multi_gpu_sharded_check.py):
the single-process multi-device store (rag_sharded_*, Collection metadata "b200:devices") on REAL devices.
Small batches must take the fused path (one scan + exchange + merge launch per device, issued by the
resident worker threads), everything else the gather + merge path; both must equal a single store over the
same rows bit for bit.  Prints SHARDED_CHECK OK and the B = 1 latency of the two paths."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import local_rag_system_b200 as rag  # noqa: E402
from local_rag_system_b200 import _native  # noqa: E402
from oracle.exact_search import round_to_bf16  # noqa: E402
from tests.conftest import unit_rows  # noqa: E402


def main():
    ndev = _native.load().rag_device_count()
    G = int(os.environ.get("SHARDS", ndev))
    assert ndev >= G >= 2, f"needs >= 2 GPUs, sees {ndev}"
    ok = True
    n, dim, k = 300_011, 768, 10
    x = round_to_bf16(unit_rows(n, dim, 11))
    x[250_000] = x[42]                                   # duplicate in another shard: tie broken by global row
    one = rag.DeviceStore(dim, "bf16", "cosine", device=0, rerank=False)
    many = rag.ShardedDeviceStore(dim, "bf16", "cosine", devices=list(range(G)), rerank=False)
    print(f"shards={many.shards()} fused_available={many.fused}", flush=True)
    ok = ok and many.fused
    one.upsert(x)
    many.upsert(x)
    rng = np.random.default_rng(5)
    dead = rng.choice(n, 5000, replace=False)
    dead = dead[(dead != 42) & (dead != 250_000)]
    passing = rng.random(n) < 0.3
    passing[[42, 250_000]] = True
    for B, regime, want_path in ((1, "stream", "fused"), (2, "stream", "fused"), (64, "tensor", "gather"), (200, "tensor", "gather")):
        q = round_to_bf16(unit_rows(B, dim, 100 + B))
        q[0] = x[42]
        for phase in ("dense", "tombstones", "filter"):
            slot = -1
            if phase == "tombstones" and B == 1:
                one.delete(dead)
                many.delete(dead)
            if phase == "filter":
                one.set_mask(0, passing)
                many.set_mask(0, passing)
                slot = 0
            for rep in range(3):                         # epochs advance: both exchange buffer halves are reused
                want = one.query(q, k, mask_slot=slot, regime=regime)
                got = many.query(q, k, mask_slot=slot, regime=regime)
                good = all(np.array_equal(g, w) for g, w in zip(got, want)) and many.last_query_info()["path"] == want_path
                good = good and got[0][0, 0] == 42 and got[0][0, 1] == 250_000
                if not good:
                    print(f"MISMATCH B={B} {regime} {phase} rep={rep} path={many.last_query_info()['path']} "
                          f"got {got[0][0].tolist()} want {want[0][0].tolist()}", flush=True)
                ok = ok and good
    # interleaved small writes and searches: parked writes reach every shard before the fused launch
    y = round_to_bf16(unit_rows(64, dim, 77))
    for i in range(64):
        r1, r2 = one.upsert(y[i:i + 1]), many.upsert(y[i:i + 1])
        got, want = many.query(y[i], k), one.query(y[i], k)
        good = np.array_equal(r1, r2) and got[0][0, 0] == r2[0] and all(np.array_equal(g, w) for g, w in zip(got, want))
        if not good:
            print(f"MISMATCH after write {i}: rows {r1} {r2} got {got[0][0].tolist()} want {want[0][0].tolist()}", flush=True)
        ok = ok and good
    # the Collection front end over real devices, with the rerank plane (product default)
    client = rag.EphemeralClient()
    col = client.get_or_create_collection("c", metadata={"hnsw:space": "cosine", "b200:dtype": "bf16",
                                                         "b200:devices": f"0-{G - 1}"})
    xs = unit_rows(50_000, 384, 3)
    col.add(ids=[f"i{j}" for j in range(50_000)], embeddings=xs, metadatas=[{"ns": "a" if j % 2 else "b"} for j in range(50_000)])
    res = col.query(query_embeddings=xs[777:778], n_results=5, where={"ns": "a"})
    good = res["ids"][0][0] == "i777" and abs(res["distances"][0][0]) < 1e-6 and col.device_store.last_query_info()["path"] == "fused"
    if not good:
        print("MISMATCH collection over devices", res["ids"], res["distances"], col.device_store.last_query_info(), flush=True)
    ok = ok and good
    client.reset()
    # latency of the two paths, B = 1 (host buffers in, host buffers out)
    q = round_to_bf16(unit_rows(256, dim, 9))
    for label, env in (("fused", None), ("gather", "0")):
        if env is not None:
            os.environ["RAG_B200_FUSED_EXCHANGE"] = env
            st = rag.ShardedDeviceStore(dim, "bf16", "cosine", devices=list(range(G)), rerank=False)
            st.upsert(x)
        else:
            st = many
        for i in range(20):
            st.query(q[i], k)
        t0 = time.perf_counter()
        for i in range(256):
            st.query(q[i], k)
        dt = (time.perf_counter() - t0) / 256
        print(f"single-process {G}-device store, B=1, path={st.last_query_info()['path']}: {dt * 1e6:.1f} us per query "
              f"({n // G} rows x {dim} per shard)", flush=True)
        if env is not None:
            st.close()
            os.environ.pop("RAG_B200_FUSED_EXCHANGE")
    one.close()
    many.close()
    print("SHARDED_CHECK", "OK" if ok else "FAILED", f"devices={G}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
