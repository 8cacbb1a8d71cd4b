"""DeviceStore: numpy-facing wrapper of one device-resident corpus shard.

Thin by design -- argument marshalling only.  All arithmetic happens in the
CUDA kernels behind the C ABI (include/rag_b200.h).  ctypes releases the GIL
around every call, so FastAPI worker threads (the reference's concurrency
model, api/routes/kb.py) overlap freely; the store's own reader/writer lock
orders queries against upserts and deletes.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as N


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _default_rerank() -> bool:
    return os.environ.get("RAG_B200_RERANK", "1") != "0"


class DeviceStore:
    """One device-resident corpus shard.  `rerank` (bf16 stores only; default on,
    RAG_B200_RERANK=0 turns the default off): also keep the un-rounded fp32 rows and
    re-rank the k + slack best bf16 hits exactly against them."""

    def __init__(self, dim: int, dtype: str = "f32", space: str = "l2", device: int = 0,
                 capacity_hint: int = 0, rerank=None):
        if space not in N.SPACES:
            raise ValueError(f"unknown space {space!r}; expected one of {sorted(N.SPACES)}")
        if dtype not in N.DTYPES:
            raise ValueError(f"unknown dtype {dtype!r}; expected f32 or bf16")
        self._lib = N.load()
        h = C.c_void_p()
        rerank = _default_rerank() if rerank is None else bool(rerank)
        N.check(self._lib.rag_store_create_ex(int(dim), N.DTYPES[dtype], N.SPACES[space], int(device),
                                              int(capacity_hint), 0 if rerank else N.STORE_NO_RERANK, C.byref(h)))
        self._h = h
        self.dim, self.space, self.device = int(dim), space, int(device)
        self.dtype = "bf16" if N.DTYPES[dtype] == N.DTYPE_BF16 else "f32"
        self.rerank = bool(self._lib.rag_store_has_rerank(self._h))

    # -- lifetime --------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.rag_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # -- state -------------------------------------------------------------------
    def count(self) -> int:
        return int(self._lib.rag_store_count(self._h))

    def rows(self) -> int:
        return int(self._lib.rag_store_rows(self._h))

    def capacity(self) -> int:
        return int(self._lib.rag_store_capacity(self._h))

    def kernel_launches(self) -> int:
        return int(self._lib.rag_store_kernel_launches(self._h))

    def is_live(self, row: int) -> bool:
        return bool(self._lib.rag_store_is_live(self._h, int(row)))

    def reserve(self, rows: int):
        N.check(self._lib.rag_store_reserve(self._h, int(rows)))

    # -- writes --------------------------------------------------------------------
    def upsert(self, vectors, rows=None) -> np.ndarray:
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        if v.ndim != 2 or v.shape[1] != self.dim:
            raise ValueError(f"vectors must be [n, {self.dim}], got {v.shape}")
        n = v.shape[0]
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        if r is not None and r.shape != (n,):
            raise ValueError("rows must have one entry per vector")
        out = np.empty(n, dtype=np.int64)
        N.check(self._lib.rag_store_upsert(self._h, n, _ptr(v), _ptr(r), _ptr(out)))
        return out

    def upsert_device(self, data_ptr: int, n: int, rows=None) -> np.ndarray:
        """Vectors already on this store's device as fp32 [n, dim] (e.g. a torch
        tensor's data_ptr()); used for bulk loads and synthetic corpora."""
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        out = np.empty(n, dtype=np.int64)
        N.check(self._lib.rag_store_upsert_dev(self._h, int(n), C.c_void_p(int(data_ptr)), _ptr(r), _ptr(out)))
        return out

    def last_upsert_ms(self) -> float:
        """Device time of the upsert kernel of the last upsert_device call (CUDA events)."""
        return float(self._lib.rag_store_last_upsert_ms(self._h))

    def delete(self, rows):
        r = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1)
        N.check(self._lib.rag_store_delete(self._h, r.shape[0], _ptr(r)))

    def flush(self):
        """Send parked small writes to the device and wait for them."""
        N.check(self._lib.rag_store_flush(self._h))

    def fetch(self, rows, exact: bool = False) -> np.ndarray:
        """Stored rows as fp32: the values the scan reads (normalised, bf16-rounded), or with
        exact=True the un-rounded fp32 plane where the store keeps one."""
        r = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1)
        out = np.empty((r.shape[0], self.dim), dtype=np.float32)
        fn = self._lib.rag_store_fetch_exact if exact else self._lib.rag_store_fetch
        N.check(fn(self._h, r.shape[0], _ptr(r), _ptr(out)))
        return out

    # -- filters ----------------------------------------------------------------------
    def set_mask(self, slot: int, passing: np.ndarray):
        """`passing`: boolean array, one entry per row (shorter is fine: missing
        rows do not pass)."""
        b = np.ascontiguousarray(passing, dtype=bool).reshape(-1)
        nbits = b.shape[0]
        packed = np.packbits(b, bitorder="little")
        pad = (-packed.shape[0]) % 8
        if pad:
            packed = np.concatenate([packed, np.zeros(pad, dtype=np.uint8)])
        words = packed.view(np.uint64) if packed.shape[0] else np.zeros(0, dtype=np.uint64)
        N.check(self._lib.rag_store_set_mask(self._h, int(slot), _ptr(words) if nbits else None, nbits))

    def patch_mask(self, slot: int, rows, passing):
        """Bit rows[i] of a parked mask := passing[i] (keeps cached masks valid across writes)."""
        r = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1)
        p = np.ascontiguousarray(passing, dtype=np.uint8).reshape(-1)
        if r.shape != p.shape:
            raise ValueError("rows and passing must have the same length")
        N.check(self._lib.rag_store_patch_mask(self._h, int(slot), r.shape[0], _ptr(r), _ptr(p)))

    def clear_mask(self, slot: int):
        N.check(self._lib.rag_store_clear_mask(self._h, int(slot)))

    # -- search -------------------------------------------------------------------------
    def query(self, queries, k: int, mask_slot: int = -1, regime: str = "auto"):
        """Returns (rows int64 [B,k], dists fp32 [B,k], counts int32 [B]); hits
        ascending by (distance, row); unused tail has row -1 / +inf."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [B, {self.dim}], got {q.shape}")
        B = q.shape[0]
        flags = {"auto": N.QUERY_AUTO, "stream": N.QUERY_FORCE_STREAM, "tensor": N.QUERY_FORCE_TENSOR}[regime]
        rows = np.empty((B, k), dtype=np.int64)
        dists = np.empty((B, k), dtype=np.float32)
        counts = np.empty(B, dtype=np.int32)
        N.check(self._lib.rag_store_query(self._h, B, _ptr(q), int(k), int(mask_slot), flags,
                                          _ptr(rows), _ptr(dists), _ptr(counts)))
        return rows, dists, counts

    def query_device(self, queries_ptr: int, B: int, k: int, out_keys_ptr: int = 0, stream: int = 0,
                     mask_slot: int = -1, row_base: int = 0, regime: str = "auto",
                     out_rows_ptr: int = 0, out_dists_ptr: int = 0, out_counts_ptr: int = 0):
        """Asynchronous shard-local search on device buffers (multi-GPU path).  Emits
        candidate keys (for the cross-shard exchange) and/or decoded rows / dists / counts."""
        flags = {"auto": N.QUERY_AUTO, "stream": N.QUERY_FORCE_STREAM, "tensor": N.QUERY_FORCE_TENSOR}[regime]
        vp = lambda p: C.c_void_p(int(p)) if p else None
        N.check(self._lib.rag_store_query_dev(self._h, int(B), C.c_void_p(int(queries_ptr)), int(k), int(mask_slot),
                                              flags, int(row_base), vp(out_keys_ptr), vp(out_rows_ptr),
                                              vp(out_dists_ptr), vp(out_counts_ptr), vp(stream)))

    def query_fused(self, exchange: "Exchange", queries_ptr: int, B: int, k: int, out_rows_ptr: int,
                    out_dists_ptr: int = 0, out_counts_ptr: int = 0, stream: int = 0, mask_slot: int = -1,
                    row_base: int = 0, regime: str = "auto"):
        """Multi-GPU search in ONE launch: shard scan, all-gather of the B x k keys over
        NVLink peer memory and the cross-shard merge are fused (see Exchange).  Every rank
        must make the same call; the GLOBAL result lands on every rank."""
        flags = {"auto": N.QUERY_AUTO, "stream": N.QUERY_FORCE_STREAM, "tensor": N.QUERY_FORCE_TENSOR}[regime]
        vp = lambda p: C.c_void_p(int(p)) if p else None
        N.check(self._lib.rag_store_query_fused_dev(self._h, exchange.handle, int(B), C.c_void_p(int(queries_ptr)),
                                                    int(k), int(mask_slot), flags, int(row_base), vp(out_rows_ptr),
                                                    vp(out_dists_ptr), vp(out_counts_ptr), vp(stream)))

    def query_fused_host(self, exchange: "Exchange", queries, k: int, mask_slot: int = -1, row_base: int = 0,
                         regime: str = "auto"):
        """query_fused with HOST buffers, synchronous: staging, H2D, the one fused launch, D2H and the
        wait all happen inside one C call (rag_store_query_fused).  Returns numpy (rows, dists, counts)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        B = q.shape[0]
        flags = {"auto": N.QUERY_AUTO, "stream": N.QUERY_FORCE_STREAM, "tensor": N.QUERY_FORCE_TENSOR}[regime]
        rows = np.empty((B, k), dtype=np.int64)
        dists = np.empty((B, k), dtype=np.float32)
        counts = np.empty(B, dtype=np.int32)
        N.check(self._lib.rag_store_query_fused(self._h, exchange.handle, B, _ptr(q), int(k), int(mask_slot), flags,
                                                int(row_base), _ptr(rows), _ptr(dists), _ptr(counts)))
        return rows, dists, counts

    def submit(self, queries, k: int, mask_slot: int = -1, regime: str = "auto", exchange: "Exchange" = None,
               row_base: int = 0):
        """Start a host-buffer query and return at once (rag_store_query_submit); up to 4 may be in flight.
        Returns a ticket for collect().  With `exchange` this is the fused multi-GPU search."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        flags = {"auto": N.QUERY_AUTO, "stream": N.QUERY_FORCE_STREAM, "tensor": N.QUERY_FORCE_TENSOR}[regime]
        t = C.c_int32(-1)
        N.check(self._lib.rag_store_query_submit(self._h, exchange.handle if exchange is not None else None, q.shape[0],
                                                 _ptr(q), int(k), int(mask_slot), flags, int(row_base), C.byref(t)))
        return (t.value, q.shape[0], int(k))

    def collect(self, ticket):
        """Wait for a submitted query; returns numpy (rows, dists, counts)."""
        t, B, k = ticket
        rows = np.empty((B, k), dtype=np.int64)
        dists = np.empty((B, k), dtype=np.float32)
        counts = np.empty(B, dtype=np.int32)
        N.check(self._lib.rag_store_query_wait(self._h, int(t), _ptr(rows), _ptr(dists), _ptr(counts)))
        return rows, dists, counts

    def fused_ok(self, exchange: "Exchange", B: int, k: int, regime: str = "auto") -> bool:
        flags = {"auto": N.QUERY_AUTO, "stream": N.QUERY_FORCE_STREAM, "tensor": N.QUERY_FORCE_TENSOR}[regime]
        return bool(self._lib.rag_store_fused_ok(self._h, exchange.handle, int(B), int(k), flags))

    def last_query_info(self):
        ms, regime, launches = C.c_float(), C.c_int32(), C.c_int32()
        N.check(self._lib.rag_store_last_query_info(self._h, C.byref(ms), C.byref(regime), C.byref(launches)))
        return {"kernel_ms": ms.value, "regime": {0: None, 1: "stream", 2: "tensor"}[regime.value],
                "launches": launches.value}

    def set_f32_shadow(self, kind: str = "auto"):
        """fp32 stores: pin the bf16 shadow the tensor regime contracts ("hi": bf16(x) filter, "hilo": split
        precision) or hand the choice back to the store ("auto"); include/rag_b200.h, rag_store_set_f32_shadow."""
        kinds = {"auto": N.F32_SHADOW_AUTO, "hi": N.F32_SHADOW_HI, "hilo": N.F32_SHADOW_HILO}
        if kind not in kinds:
            raise ValueError(f"shadow kind must be one of {sorted(kinds)}, got {kind!r}")
        N.check(self._lib.rag_store_set_f32_shadow(self._h, kinds[kind]))

    def f32_tensor_info(self):
        """Shadow kind in force and how many tensor-regime queries of this fp32 store were served / had to be
        re-run on the exact stream kernel because the guard could not certify them."""
        kind, q, r = C.c_int32(), C.c_int64(), C.c_int64()
        N.check(self._lib.rag_store_f32_tensor_info(self._h, C.byref(kind), C.byref(q), C.byref(r)))
        return {"shadow": {0: None, 1: "hi", 2: "hilo"}[kind.value], "queries": q.value, "reruns": r.value}


class ShardedDeviceStore:
    """One corpus over several devices of THIS process (include/rag_b200.h, rag_sharded_*): same
    surface as DeviceStore with GLOBAL rows.  `devices` may repeat a device (logical shards)."""

    def __init__(self, dim: int, dtype: str = "f32", space: str = "l2", devices=(0,), capacity_hint: int = 0,
                 rerank=None):
        if space not in N.SPACES:
            raise ValueError(f"unknown space {space!r}; expected one of {sorted(N.SPACES)}")
        if dtype not in N.DTYPES:
            raise ValueError(f"unknown dtype {dtype!r}; expected f32 or bf16")
        self._lib = N.load()
        devs = np.ascontiguousarray(list(devices), dtype=np.int32)
        if devs.size == 0:
            raise ValueError("a sharded store needs at least one device")
        rerank = _default_rerank() if rerank is None else bool(rerank)
        h = C.c_void_p()
        N.check(self._lib.rag_sharded_create(int(dim), N.DTYPES[dtype], N.SPACES[space], int(devs.size), _ptr(devs),
                                             int(capacity_hint), 0 if rerank else N.STORE_NO_RERANK, C.byref(h)))
        self._h = h
        self.dim, self.space, self.devices = int(dim), space, [int(d) for d in devs]
        self.device = self.devices[0]
        self.dtype = "bf16" if N.DTYPES[dtype] == N.DTYPE_BF16 else "f32"
        self.rerank = bool(self._lib.rag_store_has_rerank(self._lib.rag_sharded_shard(self._h, 0)))
        self.fused = bool(self._lib.rag_sharded_fused(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rag_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def count(self) -> int:
        return int(self._lib.rag_sharded_count(self._h))

    def rows(self) -> int:
        return int(self._lib.rag_sharded_rows(self._h))

    def shards(self) -> int:
        return int(self._lib.rag_sharded_shards(self._h))

    def shard_counts(self):
        return [int(self._lib.rag_store_count(self._lib.rag_sharded_shard(self._h, g))) for g in range(self.shards())]

    def kernel_launches(self) -> int:
        return sum(int(self._lib.rag_store_kernel_launches(self._lib.rag_sharded_shard(self._h, g)))
                   for g in range(self.shards()))

    def is_live(self, row: int) -> bool:
        return bool(self._lib.rag_sharded_is_live(self._h, int(row)))

    def reserve(self, rows: int):
        N.check(self._lib.rag_sharded_reserve(self._h, int(rows)))

    def flush(self):
        N.check(self._lib.rag_sharded_flush(self._h))

    def upsert(self, vectors, rows=None) -> np.ndarray:
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        if v.ndim != 2 or v.shape[1] != self.dim:
            raise ValueError(f"vectors must be [n, {self.dim}], got {v.shape}")
        n = v.shape[0]
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        if r is not None and r.shape != (n,):
            raise ValueError("rows must have one entry per vector")
        out = np.empty(n, dtype=np.int64)
        N.check(self._lib.rag_sharded_upsert(self._h, n, _ptr(v), _ptr(r), _ptr(out)))
        return out

    def delete(self, rows):
        r = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1)
        N.check(self._lib.rag_sharded_delete(self._h, r.shape[0], _ptr(r)))

    def fetch(self, rows, exact: bool = False) -> np.ndarray:
        r = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1)
        out = np.empty((r.shape[0], self.dim), dtype=np.float32)
        N.check(self._lib.rag_sharded_fetch(self._h, r.shape[0], _ptr(r), _ptr(out), 1 if exact else 0))
        return out

    def set_mask(self, slot: int, passing: np.ndarray):
        b = np.ascontiguousarray(passing, dtype=bool).reshape(-1)
        nbits = b.shape[0]
        packed = np.packbits(b, bitorder="little")
        pad = (-packed.shape[0]) % 8
        if pad:
            packed = np.concatenate([packed, np.zeros(pad, dtype=np.uint8)])
        words = packed.view(np.uint64) if packed.shape[0] else np.zeros(0, dtype=np.uint64)
        N.check(self._lib.rag_sharded_set_mask(self._h, int(slot), _ptr(words) if nbits else None, nbits))

    def patch_mask(self, slot: int, rows, passing):
        r = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1)
        p = np.ascontiguousarray(passing, dtype=np.uint8).reshape(-1)
        if r.shape != p.shape:
            raise ValueError("rows and passing must have the same length")
        N.check(self._lib.rag_sharded_patch_mask(self._h, int(slot), r.shape[0], _ptr(r), _ptr(p)))

    def clear_mask(self, slot: int):
        N.check(self._lib.rag_sharded_clear_mask(self._h, int(slot)))

    def query(self, queries, k: int, mask_slot: int = -1, regime: str = "auto"):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [B, {self.dim}], got {q.shape}")
        B = q.shape[0]
        flags = {"auto": N.QUERY_AUTO, "stream": N.QUERY_FORCE_STREAM, "tensor": N.QUERY_FORCE_TENSOR}[regime]
        rows = np.empty((B, k), dtype=np.int64)
        dists = np.empty((B, k), dtype=np.float32)
        counts = np.empty(B, dtype=np.int32)
        N.check(self._lib.rag_sharded_query(self._h, B, _ptr(q), int(k), int(mask_slot), flags,
                                            _ptr(rows), _ptr(dists), _ptr(counts)))
        return rows, dists, counts

    def last_query_info(self):
        ms, regime, launches, path = C.c_float(), C.c_int32(), C.c_int32(), C.c_int32()
        N.check(self._lib.rag_sharded_last_query_info(self._h, C.byref(ms), C.byref(regime), C.byref(launches),
                                                      C.byref(path)))
        return {"kernel_ms": ms.value, "regime": {0: None, 1: "stream", 2: "tensor"}[regime.value],
                "launches": launches.value, "path": {0: None, 1: "fused", 2: "gather"}[path.value]}


def merge_keys_device(device: int, G: int, B: int, k: int, keys_ptr: int, out_keys_ptr: int = 0,
                      out_rows_ptr: int = 0, out_dists_ptr: int = 0, out_counts_ptr: int = 0, stream: int = 0):
    """Cross-shard merge of G x B x k candidate keys on the device."""
    lib = N.load()
    vp = lambda p: C.c_void_p(int(p)) if p else None
    N.check(lib.rag_merge_keys_dev(int(device), int(G), int(B), int(k), vp(keys_ptr), vp(out_keys_ptr),
                                   vp(out_rows_ptr), vp(out_dists_ptr), vp(out_counts_ptr), vp(stream)))


class Exchange:
    """This rank's peer-mapped exchange buffer for the fused multi-GPU search
    (include/rag_b200.h, "fused cross-shard exchange").  `all_gather_bytes` is any
    callable that takes this rank's 64-byte handle and returns the world's handles
    concatenated in rank order (torch.distributed in sharded.py)."""

    def __init__(self, device: int, rank: int, world: int, all_gather_bytes, slot_keys: int = 8192):
        self._lib = N.load()
        h = C.c_void_p()
        self._h = None
        self.rank, self.world, self.slot_keys = rank, world, slot_keys
        mine = C.create_string_buffer(N.EXCHANGE_HANDLE_BYTES)
        rc = self._lib.rag_exchange_create(int(device), int(rank), int(world), int(slot_keys), C.byref(h))
        msg = N.last_error() if rc != 0 else ""
        if rc == 0:
            self._h = h
            rc = self._lib.rag_exchange_handle(self._h, mine)
            msg = N.last_error() if rc != 0 else ""
        # the all-gather is collective: take part even if creating / exporting the buffer failed, then report
        everyone = bytes(all_gather_bytes(mine.raw))
        try:
            if rc != 0:
                raise N.EngineError(f"[rag_b200 {rc}] {msg}")
            if len(everyone) != world * N.EXCHANGE_HANDLE_BYTES:
                raise ValueError("all_gather_bytes must return world x 64 bytes")
            N.check(self._lib.rag_exchange_connect(self._h, everyone))
        except Exception:
            self.close()
            raise

    @property
    def handle(self):
        return self._h

    def timed_out(self) -> bool:
        v = C.c_int32()
        N.check(self._lib.rag_exchange_status(self._h, C.byref(v)))
        return bool(v.value)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rag_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
