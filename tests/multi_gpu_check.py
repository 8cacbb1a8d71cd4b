"""Run under torchrun on N GPUs (gpurun --gpus N): the row-sharded search over N
ranks must return exactly what a single store over the whole corpus returns
(rows and distances, ties included), in both kernel regimes, with a filter and
tombstones.  Rank 0 prints MULTI_GPU_CHECK OK."""
import os
import sys

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
    searcher = ShardedSearcher(shard, rank, world, row_base=lo)
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
            got_r, got_d, got_c = searcher.search(q, k, mask_slot=slot, regime=regime)
            good = np.array_equal(got_r, want_r) and np.array_equal(got_d, want_d) and np.array_equal(got_c, want_c)
            if not good:
                bad = np.nonzero((got_r != want_r).any(axis=1))[0][:3]
                print(f"[rank {rank}] MISMATCH B={B} {regime} {phase}: queries {bad.tolist()} "
                      f"got {got_r[bad[0]].tolist()} want {want_r[bad[0]].tolist()}", flush=True)
            ok = ok and good
        if phase == "filter":      # restore for the next batch size
            pass
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "OK" if int(t.item()) == 1 else "FAILED", f"world={world}", flush=True)
    shard.close()
    full.close()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
