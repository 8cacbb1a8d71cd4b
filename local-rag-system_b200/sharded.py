"""Row-sharded exact search over the GPUs of one box (K6).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  The corpus
is split row-wise into contiguous shards; rank g owns global rows
[g * stride, g * stride + rows_g).  A query batch is replicated to every rank,
each rank runs the shard-local scan + fused top-k (no data-path collective),
and the only exchange is one all-gather of B x k 64-bit candidate keys per
rank (80 B per rank for B=1, k=10) followed by a merge.  Because keys carry
global rows, the merged result is bit-identical to what a single store holding
the whole corpus returns -- ties included.

Two implementations of the exchange:
  * fused (small batches, the latency-bound case): the scan kernel's last CTA
    stores the shard's keys into every rank's peer-mapped buffer over NVLink,
    raises a flag, waits for the others and merges -- ONE launch per batch per
    GPU and no collective call (engine.Exchange, csrc/scan_stream.cu);
  * NCCL all_gather_into_tensor + the merge kernel (large batches / tensor
    regime, where the exchange is bandwidth- not latency-bound).

torch is used for what it is good at here: device buffers, streams and the
process group.  The scan, select and merge are the engine's own kernels.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .engine import DeviceStore, Exchange, merge_keys_device


def shard_plan(total_rows: int, world: int) -> Tuple[int, list]:
    """Contiguous row split: returns (stride, [rows per rank]).  stride =
    ceil(total / world) is also each rank's global row base multiplier."""
    stride = (total_rows + world - 1) // world
    counts = [max(0, min(stride, total_rows - g * stride)) for g in range(world)]
    if stride * world >= 2 ** 32:
        raise ValueError("global rows must stay below 2^32 (keys carry 32-bit rows)")
    return stride, counts


def exchange_candidates(local_keys: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """All-gather of per-shard candidate keys.  local_keys: int64 [B*k] (the
    uint64 key bits) on the rank's device (or CPU under gloo, for tests).
    Returns int64 [world, B*k], identical on every rank."""
    out = torch.empty((world,) + tuple(local_keys.shape), dtype=local_keys.dtype, device=local_keys.device)
    if world == 1:
        out[0].copy_(local_keys)
        return out
    dist.all_gather_into_tensor(out.view(-1), local_keys.contiguous().view(-1), group=group)
    return out


class ShardedSearcher:
    """Search front-end of one rank.  Every rank must call search*() with the
    same queries (the query batch is replicated)."""

    def __init__(self, store: DeviceStore, rank: int = 0, world: int = 1, row_base: int = 0, group=None,
                 fused_exchange: bool = True):
        self.store, self.rank, self.world, self.row_base, self.group = store, rank, world, int(row_base), group
        self.device = torch.device("cuda", store.device)
        self._buf = {}
        self.exchange: Optional[Exchange] = None
        self.last_path = None
        if world > 1 and fused_exchange and os.environ.get("RAG_B200_FUSED_EXCHANGE", "1") != "0":
            self.exchange = self._connect_exchange()

    def _connect_exchange(self) -> Optional[Exchange]:
        """Map every rank's exchange buffer (CUDA IPC).  Collective: if mapping fails on ANY rank (no peer
        access, IPC not permitted in the container) every rank drops to the NCCL exchange -- a mixed setup
        would deadlock."""
        x, err = None, None
        try:
            x = Exchange(self.store.device, self.rank, self.world, self._all_gather_bytes)
        except Exception as e:       # noqa: BLE001 - reported below, the NCCL path still works
            err = e
        ok = torch.tensor([0 if x is None else 1], device=self.device, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 1:
            return x
        if x is not None:
            x.close()
        if self.rank == 0:
            import warnings
            warnings.warn(f"fused cross-GPU exchange unavailable ({err or 'a peer could not map the buffers'}); "
                          "using the NCCL all-gather + merge kernel instead")
        return None

    def _all_gather_bytes(self, mine: bytes) -> bytes:
        t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(self.device)
        out = torch.empty(self.world * t.numel(), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out.cpu().numpy().tobytes()

    def close(self):
        if self.exchange is not None:
            if self.world > 1:
                torch.cuda.synchronize(self.device)
                dist.barrier(group=self.group)     # nobody unmaps while a peer may still store into it
            self.exchange.close()
            self.exchange = None

    def _buffers(self, B: int, k: int):
        key = (B, k)
        b = self._buf.get(key)
        if b is None:
            dev = self.device
            # rows | dists | counts live in ONE device block and one pinned host block: a single D2H per batch
            nr, nd, nc = B * k * 8, B * k * 4, B * 4
            out = torch.empty(nr + nd + nc, dtype=torch.uint8, device=dev)
            h_out = torch.empty(nr + nd + nc, dtype=torch.uint8).pin_memory()
            b = {
                "local": torch.empty(B * k, dtype=torch.int64, device=dev),
                "out": out, "h_out": h_out,
                "rows": out[:nr].view(torch.int64).view(B, k),
                "dists": out[nr:nr + nd].view(torch.float32).view(B, k),
                "counts": out[nr + nd:].view(torch.int32),
                "q": torch.empty((B, self.store.dim), dtype=torch.float32, device=dev),
                "h_rows": h_out[:nr].view(torch.int64).view(B, k),
                "h_dists": h_out[nr:nr + nd].view(torch.float32).view(B, k),
                "h_counts": h_out[nr + nd:].view(torch.int32),
            }
            self._buf[key] = b
        return b

    def search_device(self, q_dev: torch.Tensor, k: int, mask_slot: int = -1, regime: str = "auto"):
        """q_dev: fp32 [B, dim] on this rank's device.  Asynchronous on the current
        stream.  Returns device tensors (global rows int64 [B,k], dists fp32 [B,k],
        counts int32 [B])."""
        B = q_dev.shape[0]
        b = self._buffers(B, k)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if self.world == 1:      # one shard: the store's own kernels emit the final result, no exchange
            self.store.query_device(q_dev.data_ptr(), B, k, 0, stream=stream, mask_slot=mask_slot,
                                    row_base=self.row_base, regime=regime, out_rows_ptr=b["rows"].data_ptr(),
                                    out_dists_ptr=b["dists"].data_ptr(), out_counts_ptr=b["counts"].data_ptr())
            return b["rows"], b["dists"], b["counts"]
        if self.exchange is not None and self.store.fused_ok(self.exchange, B, k, regime):
            self.store.query_fused(self.exchange, q_dev.data_ptr(), B, k, b["rows"].data_ptr(), b["dists"].data_ptr(),
                                   b["counts"].data_ptr(), stream=stream, mask_slot=mask_slot,
                                   row_base=self.row_base, regime=regime)
            self.last_path = "fused"
            return b["rows"], b["dists"], b["counts"]
        self.last_path = "nccl"
        self.store.query_device(q_dev.data_ptr(), B, k, b["local"].data_ptr(), stream=stream,
                                mask_slot=mask_slot, row_base=self.row_base, regime=regime)
        gathered = exchange_candidates(b["local"], self.world, self.group)
        merge_keys_device(self.store.device, self.world, B, k, gathered.data_ptr(), 0, b["rows"].data_ptr(),
                          b["dists"].data_ptr(), b["counts"].data_ptr(), stream=stream)
        return b["rows"], b["dists"], b["counts"]

    def submit(self, queries: np.ndarray, k: int, mask_slot: int = -1, regime: str = "auto"):
        """Start a host-to-host search and return a handle for collect(); up to 4 may be in flight (a server
        with concurrent requests).  Every rank must submit and collect in the same order."""
        qn = np.ascontiguousarray(queries, dtype=np.float32)
        if qn.ndim == 1:
            qn = qn[None, :]
        B = qn.shape[0]
        if self.world == 1:
            return ("ticket", self.store.submit(qn, k, mask_slot=mask_slot, regime=regime, row_base=self.row_base))
        if self.exchange is not None and self.store.fused_ok(self.exchange, B, k, regime):
            self.last_path = "fused"
            return ("ticket", self.store.submit(qn, k, mask_slot=mask_slot, regime=regime, exchange=self.exchange,
                                                row_base=self.row_base))
        return ("done", self.search(qn, k, mask_slot=mask_slot, regime=regime))     # collective path: answered at once

    def collect(self, handle):
        kind, payload = handle
        return self.store.collect(payload) if kind == "ticket" else payload

    def search(self, queries: np.ndarray, k: int, mask_slot: int = -1, regime: str = "auto"):
        """Host-to-host call: pinned H2D of the queries, shard search, exchange,
        merge, D2H of the final B x k result.  Returns numpy (rows, dists, counts)."""
        if self.world == 1:
            # one shard: the engine's own host-buffer call (pinned staging, H2D, search, one D2H, sync -- all in C)
            rows, dists, counts = self.store.query(queries, k, mask_slot=mask_slot, regime=regime)
            if self.row_base:
                rows = np.where(rows >= 0, rows + self.row_base, rows)
            return rows, dists, counts
        qn = np.ascontiguousarray(queries, dtype=np.float32)
        if qn.ndim == 1:
            qn = qn[None, :]
        B = qn.shape[0]
        if self.exchange is not None and self.store.fused_ok(self.exchange, B, k, regime):
            # ONE C call per rank: pinned staging, H2D, the fused scan + exchange + merge launch, D2H, wait
            # (rag_store_query_fused).  No torch op, no Python between the copies and the launch.
            self.last_path = "fused"
            return self.store.query_fused_host(self.exchange, qn, k, mask_slot=mask_slot, row_base=self.row_base,
                                               regime=regime)
        q = torch.from_numpy(qn)
        b = self._buffers(B, k)
        b["q"].copy_(q if q.is_pinned() else q.pin_memory(), non_blocking=True)
        rows, dists, counts = self.search_device(b["q"], k, mask_slot, regime)
        b["h_out"].copy_(b["out"], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return b["h_rows"].numpy().copy(), b["h_dists"].numpy().copy(), b["h_counts"].numpy().copy()
