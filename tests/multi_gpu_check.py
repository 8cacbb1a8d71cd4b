"""Run under torchrun on N GPUs (gpurun --gpus N): the row-sharded search over N
ranks must return exactly what a single store over the whole corpus returns
(rows and distances, ties included), in both kernel regimes, with a filter and
tombstones.  Rank 0 prints MULTI_GPU_CHECK OK."""
import os
import sys

os.environ.setdefault("RAG_B200_RERANK", "0")     # bit-identity of sharded vs single search is a property of the scan planes

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import local_rag_system_b200 as rag  # noqa: E402
from local_rag_system_b200.sharded import ShardedSearcher, shard_plan  # noqa: E402
from oracle.exact_search import round_to_bf16  # noqa: E402
from tests.conftest import unit_rows  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, dim, k = 200_003, 768, 10
    x = round_to_bf16(unit_rows(n, dim, 11))
    x[150_000] = x[42]                                   # duplicate across shards: tie broken by global row
    stride, counts = shard_plan(n, world)
    lo = rank * stride
    shard = rag.DeviceStore(dim, "bf16", "cosine", device=local)
    shard.upsert(x[lo:lo + counts[rank]])
    full = rag.DeviceStore(dim, "bf16", "cosine", device=local)      # the single-store reference, on every rank
    full.upsert(x)
    rng = np.random.default_rng(5)
    dead = rng.choice(n, 5000, replace=False)
    passing = rng.random(n) < 0.3
    searchers = {"fused": ShardedSearcher(shard, rank, world, row_base=lo, fused_exchange=True),
                 "nccl": ShardedSearcher(shard, rank, world, row_base=lo, fused_exchange=False)}
    ok = True
    for B, regime in ((1, "stream"), (3, "stream"), (64, "tensor"), (200, "tensor")):
        q = round_to_bf16(unit_rows(B, dim, 100 + B))
        q[0] = x[42]
        for phase in ("dense", "tombstones", "filter"):
            if phase == "tombstones":
                full.delete(dead)
                mine = dead[(dead >= lo) & (dead < lo + counts[rank])] - lo
                shard.delete(mine)
            slot = -1
            if phase == "filter":
                full.set_mask(0, passing)
                shard.set_mask(0, passing[lo:lo + counts[rank]])
                slot = 0
            want_r, want_d, want_c = full.query(q, k, mask_slot=slot, regime=regime)
            for name, searcher in searchers.items():
                got_r, got_d, got_c = searcher.search(q, k, mask_slot=slot, regime=regime)
                good = np.array_equal(got_r, want_r) and np.array_equal(got_d, want_d) and np.array_equal(got_c, want_c)
                if name == "fused" and regime == "stream":      # the one-launch path must be the one that ran
                    good = good and searcher.last_path == "fused"
                if not good:
                    bad = np.nonzero((got_r != want_r).any(axis=1))[0][:3]
                    print(f"[rank {rank}] MISMATCH {name} B={B} {regime} {phase}: path {searcher.last_path} queries "
                          f"{bad.tolist()} got {got_r[bad[0]].tolist() if len(bad) else None} "
                          f"want {want_r[bad[0]].tolist() if len(bad) else None}", flush=True)
                ok = ok and good
        if phase == "filter":      # restore for the next batch size
            pass
    # queries in flight through the fused exchange (submit / collect, 3 in the air): same answers, same order
    qs = [round_to_bf16(unit_rows(1, dim, 300 + i)) for i in range(12)]
    want = [full.query(q, k, regime="stream") for q in qs]
    pend, got = [], []
    for q in qs:
        pend.append(searchers["fused"].submit(q, k, regime="stream"))
        if len(pend) == 3:
            got.append(searchers["fused"].collect(pend.pop(0)))
    while pend:
        got.append(searchers["fused"].collect(pend.pop(0)))
    good = all(np.array_equal(g[j], w[j]) for g, w in zip(got, want) for j in range(3)) and searchers["fused"].last_path == "fused"
    if not good:
        print(f"[rank {rank}] MISMATCH queries in flight through the fused exchange", flush=True)
    ok = ok and good
    ok = fp32_and_empty_shard_cases(rank, world, local) and ok
    ok = ok and not searchers["fused"].exchange.timed_out()
    latency_report(searchers, rank, dim, k)
    for s_ in searchers.values():
        s_.close()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if int(t.item()) == 1 else "FAILED", f"world={world}", flush=True)
    shard.close()
    full.close()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


def fp32_and_empty_shard_cases(rank, world, local):
    """fp32 store, l2, several query groups per launch (B = 20 -> 3 groups), k = 37, and a
    corpus so small that the last ranks hold nothing: they must still take part in the exchange."""
    ok = True
    for n, dim, B, k in ((5003, 96, 20, 37), (world - 1 if world > 2 else 1, 32, 3, 5)):
        x = unit_rows(n, dim, 21)
        stride, counts = shard_plan(n, world)
        lo = rank * stride
        shard = rag.DeviceStore(dim, "f32", "l2", device=local)
        if counts[rank]:
            shard.upsert(x[lo:lo + counts[rank]])
        full = rag.DeviceStore(dim, "f32", "l2", device=local)
        full.upsert(x)
        fused = ShardedSearcher(shard, rank, world, row_base=lo, fused_exchange=True)
        q = unit_rows(B, dim, 22)
        for rep in range(3):                              # epochs 1..3: both buffer halves get reused
            want = full.query(q, k, regime="stream")
            got = fused.search(q, k, regime="stream")
            good = all(np.array_equal(g, w) for g, w in zip(got, want)) and fused.last_path == "fused"
            if not good:
                print(f"[rank {rank}] MISMATCH fp32/empty-shard case n={n} rep={rep} path={fused.last_path}", flush=True)
            ok = ok and good
        # automatic regime (B = 20 on an fp32 store -> split-precision tensor regime + NCCL exchange): same answer
        want = full.query(q, k)
        got = fused.search(q, k)
        good = np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2]) and \
            np.allclose(got[1], want[1], rtol=1e-5, atol=2e-6)
        if not good:
            print(f"[rank {rank}] MISMATCH fp32 auto-regime case n={n} path={fused.last_path}", flush=True)
        ok = ok and good
        ok = ok and not fused.exchange.timed_out()
        fused.close()
        shard.close()
        full.close()
    return ok


def latency_report(searchers, rank, dim, k):
    """B = 1 step time of the two exchange implementations on this (small) shard."""
    q = torch.from_numpy(unit_rows(1, dim, 77)).cuda()
    for name, s_ in searchers.items():
        for _ in range(20):
            s_.search_device(q, k)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(300):
            s_.search_device(q, k)
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"exchange={name}: {e0.elapsed_time(e1) / 300 * 1e3:.1f} us per B=1 step "
                  f"(shard {s_.store.rows()} rows x {dim})", flush=True)


if __name__ == "__main__":
    main()
