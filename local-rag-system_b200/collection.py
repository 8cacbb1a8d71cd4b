"""Chroma-compatible host layer over the device engine.

Mirrors the slice of the `chromadb` Python API the reference calls
(SURVEY.md 8a/8b): PersistentClient(path) -> get_or_create_collection(name,
embedding_function, metadata) -> Collection.{add, upsert, query, get, delete,
count}.  Reference call sites: api/app.py:87-91, 209-225, 264-271, 301-315,
539-566; api/routes/system.py:33; scripts/build_index.py:15-17, 89-96;
scripts/query_local.py:21-34; scripts/ingest_docs_to_chroma.py:9-31.

What lives here (host): string ids <-> dense rows, documents, metadata (row
dicts + typed columns for `where`), result assembly, the embedding-function
hook.  What lives on the device: vectors, live bitmap, filter bitmaps, all
distance arithmetic and top-k selection.  Nothing here computes a distance.
"""
from __future__ import annotations

import json
import logging
import os
import threading
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _native as N
from .engine import DeviceStore, ShardedDeviceStore
from .where import (MetadataColumns, evaluate_where_document, kind_of, match_document, match_record,
                    validate_where, validate_where_document)

logger = logging.getLogger("local_rag_system_b200")

DEFAULT_INCLUDE = ("metadatas", "documents", "distances")
_VALID_INCLUDE = {"metadatas", "documents", "distances", "embeddings", "uris", "data"}


class _RWLock:
    """Writer-preferring reader/writer lock (queries share, writes exclude)."""

    def __init__(self):
        self._cv = threading.Condition(threading.Lock())
        self._readers = 0
        self._writer = False
        self._waiting_writers = 0

    def acquire_read(self):
        with self._cv:
            while self._writer or self._waiting_writers:
                self._cv.wait()
            self._readers += 1

    def release_read(self):
        with self._cv:
            self._readers -= 1
            if self._readers == 0:
                self._cv.notify_all()

    def acquire_write(self):
        with self._cv:
            self._waiting_writers += 1
            while self._writer or self._readers:
                self._cv.wait()
            self._waiting_writers -= 1
            self._writer = True

    def release_write(self):
        with self._cv:
            self._writer = False
            self._cv.notify_all()


class _Shared:
    def __init__(self, lock):
        self.lock = lock

    def __enter__(self):
        self.lock.acquire_read()

    def __exit__(self, *a):
        self.lock.release_read()


class _Exclusive(_Shared):
    def __enter__(self):
        self.lock.acquire_write()

    def __exit__(self, *a):
        self.lock.release_write()


def _grow_obj(arr: np.ndarray, n: int) -> np.ndarray:
    if n <= arr.shape[0]:
        return arr
    new = np.empty(max(n, 2 * arr.shape[0], 64), dtype=object)
    new[:arr.shape[0]] = arr
    return new


class _CollectionState:
    """Process-wide state of one collection (shared by every client handle that
    names the same path + collection, as the reference constructs several
    PersistentClients per process: api/app.py:89, 218, 267, 303)."""

    def __init__(self, name: str, metadata: Optional[dict], path: Optional[str]):
        self.name = name
        self.path = path
        self.apply_metadata(metadata)
        self.store: Optional[DeviceStore] = None
        self.dim: Optional[int] = None
        self.row_of: Dict[str, int] = {}
        self.ids = np.empty(0, dtype=object)        # row -> id (None when dead)
        self.docs = np.empty(0, dtype=object)
        self.metas = np.empty(0, dtype=object)
        self.columns = MetadataColumns()
        self.rw = _RWLock()
        # where -> device mask slot cache
        self.mask_mu = threading.Lock()
        self.mask_slots: Dict[str, int] = {}       # canonical where -> slot
        self.mask_pred: Dict[str, tuple] = {}      # canonical where -> (where, where_document)
        self.mask_valid: Dict[int, bool] = {}      # slot -> the device bitmap reflects every write so far
        self.mask_pins: Dict[int, int] = {}
        self.mask_lru: List[str] = []
        self.meta_version = 0
        self.mask_uploads = 0                       # full bitmap evaluations + uploads (tests / diagnostics)
        self.mask_patches = 0                       # incremental updates
        self.journal = None                         # persistence hook (persist.py)

    def apply_metadata(self, metadata: Optional[dict]):
        """Collection metadata carries the engine knobs: Chroma's `hnsw:space`
        (l2 | cosine | ip; default l2 as in the reference, api/app.py:91) plus
        b200:dtype (f32 | bf16), b200:device, b200:devices ("0-7" / "0,2,4" / [0, 1]: one
        collection sharded over several GPUs of this process), b200:capacity,
        b200:rerank ("f32" default | "none": bf16 stores keep / drop the fp32 re-ranking plane)."""
        self.metadata = dict(metadata) if metadata else None
        md = self.metadata or {}
        self.space = str(md.get("hnsw:space", os.environ.get("RAG_B200_SPACE", "l2")))
        if self.space not in N.SPACES:
            raise ValueError(f"hnsw:space must be one of {sorted(N.SPACES)}, got {self.space!r}")
        self.dtype = str(md.get("b200:dtype", os.environ.get("RAG_B200_DTYPE", "f32")))
        if self.dtype not in N.DTYPES:
            raise ValueError(f"b200:dtype must be f32 or bf16, got {self.dtype!r}")
        self.device = int(md.get("b200:device", os.environ.get("RAG_B200_DEVICE", "0")))
        self.devices = _parse_devices(md.get("b200:devices", os.environ.get("RAG_B200_DEVICES")))
        self.capacity_hint = int(md.get("b200:capacity", 0))
        rr = md.get("b200:rerank", os.environ.get("RAG_B200_RERANK"))
        self.rerank = None if rr is None else str(rr).lower() not in ("0", "none", "off", "false")

    def ensure_store(self, dim: int):
        if self.store is None:
            if self.devices and len(self.devices) > 1:
                self.store = ShardedDeviceStore(dim, self.dtype, self.space, self.devices, self.capacity_hint, self.rerank)
            else:
                dev = self.devices[0] if self.devices else self.device
                self.store = DeviceStore(dim, self.dtype, self.space, dev, self.capacity_hint, self.rerank)
            self.dim = dim
        elif dim != self.dim:
            raise ValueError(f"Embedding dimension {dim} does not match collection dimensionality {self.dim}")

    def n_rows(self) -> int:
        return 0 if self.store is None else self.store.rows()


def _parse_devices(spec):
    """"0-7" | "0,2,4" | "0-3,6" | [0, 1] | 3 -> list of device ordinals (None when unset)."""
    if spec is None or spec == "":
        return None
    if isinstance(spec, (int, np.integer)):
        return [int(spec)]
    if isinstance(spec, (list, tuple)):
        return [int(d) for d in spec]
    out: List[int] = []
    for part in str(spec).split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-", 1)
            out.extend(range(int(lo), int(hi) + 1))
        else:
            out.append(int(part))
    if not out:
        raise ValueError(f"b200:devices must name at least one device, got {spec!r}")
    return out


def _validate_metadata(m):
    if m is None:
        return None
    if not isinstance(m, dict):
        raise ValueError(f"Expected metadata to be a dict or None, got {m!r}")
    for k, v in m.items():
        if not isinstance(k, str):
            raise ValueError(f"Expected metadata key to be a str, got {k!r}")
        kind_of(v)   # raises ValueError on list / dict / None values, as Chroma does
    return dict(m) if m else None


class Collection:
    """Handle on a collection.  Handles created with and without an embedding
    function share the same underlying state."""

    def __init__(self, state: _CollectionState, embedding_function=None):
        self._s = state
        self._ef = embedding_function
        self.name = state.name
        self.metadata = state.metadata

    # ------------------------------------------------------------------ utils
    def _embed(self, texts: Sequence[str]) -> np.ndarray:
        if self._ef is None:
            raise ValueError("You must provide an embedding function to compute embeddings "
                             "(pass embeddings=/query_embeddings= or create the collection with embedding_function=)")
        out = self._ef(list(texts))
        emb = np.asarray(out, dtype=np.float32)
        if emb.ndim != 2 or emb.shape[0] != len(texts):
            raise ValueError("embedding function returned a wrong shape")
        return emb

    @staticmethod
    def _as_list(x):
        if x is None:
            return None
        if isinstance(x, (str, bytes, dict)):
            return [x]
        return list(x)

    def _prepare_records(self, ids, embeddings, metadatas, documents):
        ids = self._as_list(ids)
        if not ids:
            raise ValueError("Expected ids to be a non-empty list")
        for i in ids:
            if not isinstance(i, str):
                raise ValueError(f"Expected ID to be a str, got {i!r}")
        if len(set(ids)) != len(ids):
            dup = sorted({i for i in ids if ids.count(i) > 1})
            raise ValueError(f"Expected IDs to be unique, found duplicates of: {', '.join(dup[:5])}")
        n = len(ids)
        documents = self._as_list(documents)
        metadatas = self._as_list(metadatas)
        if documents is not None and len(documents) != n:
            raise ValueError(f"Number of documents {len(documents)} must match number of ids {n}")
        if metadatas is not None and len(metadatas) != n:
            raise ValueError(f"Number of metadatas {len(metadatas)} must match number of ids {n}")
        if embeddings is None:
            if documents is None:
                raise ValueError("You must provide embeddings or documents")
            emb = self._embed(documents)
        else:
            emb = np.asarray(embeddings, dtype=np.float32)
            if emb.ndim == 1 and n == 1:
                emb = emb[None, :]
            if emb.ndim != 2 or emb.shape[0] != n:
                raise ValueError(f"Number of embeddings {emb.shape[0] if emb.ndim else 0} must match number of ids {n}")
        if emb.shape[1] == 0:
            raise ValueError("Expected each embedding to be a non-empty list")
        metas = [_validate_metadata(m) for m in metadatas] if metadatas is not None else [None] * n
        docs = list(documents) if documents is not None else [None] * n
        return ids, np.ascontiguousarray(emb), metas, docs

    _PATCH_MAX_ROWS = 4096      # larger writes re-evaluate the predicate instead of patching row by row

    def _write_rows(self, s: _CollectionState, rows: np.ndarray, ids, metas, docs):
        top = int(rows.max()) + 1
        s.ids, s.docs, s.metas = _grow_obj(s.ids, top), _grow_obj(s.docs, top), _grow_obj(s.metas, top)
        for r, i, m, d in zip(rows.tolist(), ids, metas, docs):
            s.columns.set_row(r, s.metas[r], m)
            s.ids[r], s.docs[r], s.metas[r] = i, d, m
            s.row_of[i] = r
        s.meta_version += 1
        self._patch_masks(s, rows)

    def _patch_masks(self, s: _CollectionState, rows: np.ndarray):
        """Keep the cached `where` bitmaps valid across a write: only the written rows' bits change
        (api/app.py interleaves col.add with filtered /search; re-evaluating the predicate over every
        row and re-uploading N/8 bytes per query after each write is what this avoids)."""
        with s.mask_mu:
            if not s.mask_slots:
                return
            if rows.size > self._PATCH_MAX_ROWS:
                for slot in s.mask_slots.values():
                    s.mask_valid[slot] = False
                return
            rl = rows.tolist()
            for key, slot in s.mask_slots.items():
                if not s.mask_valid.get(slot):
                    continue
                where, wd = s.mask_pred[key]
                passing = np.fromiter((match_record(where, s.metas[r]) and match_document(wd, s.docs[r]) for r in rl),
                                      dtype=np.uint8, count=len(rl))
                s.store.patch_mask(slot, rows, passing)
                s.mask_patches += 1

    # ----------------------------------------------------------------- writes
    def add(self, ids, embeddings=None, metadatas=None, documents=None, uris=None, images=None):
        """Insert new records; ids that already exist are skipped with a warning
        (Chroma semantics).  Reference: api/app.py:221, ingest_docs_to_chroma.py:31."""
        ids, emb, metas, docs = self._prepare_records(ids, embeddings, metadatas, documents)
        s = self._s
        with _Exclusive(s.rw):
            s.ensure_store(emb.shape[1])
            keep = [j for j, i in enumerate(ids) if i not in s.row_of]
            if len(keep) != len(ids):
                skipped = [i for i in ids if i in s.row_of]
                logger.warning("Add of existing embedding ID: %s", ", ".join(skipped[:8]))
            if not keep:
                return
            sel_ids = [ids[j] for j in keep]
            rows = s.store.upsert(emb[keep], None)
            self._write_rows(s, rows, sel_ids, [metas[j] for j in keep], [docs[j] for j in keep])
            if s.journal is not None:
                s.journal.log("add", sel_ids, emb[keep], [metas[j] for j in keep], [docs[j] for j in keep])

    def upsert(self, ids, embeddings=None, metadatas=None, documents=None, uris=None, images=None):
        """Insert, or update in place.  Reference: scripts/build_index.py:92-96.  For an id that
        exists the vector is replaced and -- as Chroma's metadata segment does for an UPSERT record
        (SqliteMetadataSegment._update_metadata: keys given are inserted-or-replaced, keys not given
        stay; the document is the key "chroma:document") -- metadata keys are MERGED and the document is
        kept when the call passes none."""
        had_docs = documents is not None
        ids, emb, metas, docs = self._prepare_records(ids, embeddings, metadatas, documents)
        s = self._s
        with _Exclusive(s.rw):
            s.ensure_store(emb.shape[1])
            want = np.array([s.row_of.get(i, -1) for i in ids], dtype=np.int64)
            for j, r in enumerate(want.tolist()):
                if r < 0:
                    continue
                if s.metas[r]:
                    merged = dict(s.metas[r])
                    merged.update(metas[j] or {})
                    metas[j] = merged
                if not had_docs:
                    docs[j] = s.docs[r]
            rows = s.store.upsert(emb, want)
            self._write_rows(s, rows, ids, metas, docs)
            if s.journal is not None:
                s.journal.log("upsert", ids, emb, metas, docs)

    def update(self, ids, embeddings=None, metadatas=None, documents=None):
        """Update fields of existing records; unknown ids are ignored with a warning."""
        ids_l = self._as_list(ids)
        s = self._s
        with _Exclusive(s.rw):
            known = [j for j, i in enumerate(ids_l) if i in s.row_of]
            if len(known) != len(ids_l):
                logger.warning("Update of nonexisting embedding ID(s)")
            if not known:
                return
            documents = self._as_list(documents)
            metadatas = self._as_list(metadatas)
            emb = None
            if embeddings is not None:
                emb = np.asarray(embeddings, dtype=np.float32)[known]
            elif documents is not None:
                emb = self._embed([documents[j] for j in known])
            rows = np.array([s.row_of[ids_l[j]] for j in known], dtype=np.int64)
            if emb is not None:
                s.ensure_store(emb.shape[1])
                s.store.upsert(emb, rows)
            for n_, j in enumerate(known):
                r = int(rows[n_])
                if metadatas is not None:
                    merged = dict(s.metas[r] or {})
                    merged.update(_validate_metadata(metadatas[j]) or {})
                    s.columns.set_row(r, s.metas[r], merged)
                    s.metas[r] = merged or None
                if documents is not None:
                    s.docs[r] = documents[j]
            s.meta_version += 1
            self._patch_masks(s, rows)
            if s.journal is not None:
                vec = s.store.fetch(rows, exact=True) if emb is None else emb
                s.journal.log("upsert", [ids_l[j] for j in known], vec, [s.metas[int(r)] for r in rows],
                              [s.docs[int(r)] for r in rows])

    def delete(self, ids=None, where=None, where_document=None):
        """Delete by ids and/or predicate.  Reference: api/app.py:269, 306, 311.
        Unknown ids only warn; a predicate matching nothing is a no-op."""
        ids = self._as_list(ids)
        validate_where(where)
        validate_where_document(where_document)
        if ids is None and not where and not where_document:
            raise ValueError("You must provide either ids, where, or where_document to delete.")
        s = self._s
        with _Exclusive(s.rw):
            if s.store is None:
                return []
            n = s.n_rows()
            if ids is not None:
                rows = [s.row_of[i] for i in ids if i in s.row_of]
                if len(rows) != len(ids):
                    logger.warning("Delete of nonexisting embedding ID(s)")
                cand = np.zeros(n, dtype=bool)
                cand[rows] = True
            else:
                cand = np.zeros(n, dtype=bool)
                cand[list(s.row_of.values())] = True
            if where:
                cand &= s.columns.evaluate(where, n)
            if where_document:
                cand &= evaluate_where_document(where_document, s.docs, n)
            victims = np.nonzero(cand)[0]
            if victims.size == 0:
                return []
            gone = [s.ids[r] for r in victims.tolist()]
            s.store.delete(victims)
            for r, i in zip(victims.tolist(), gone):
                s.columns.set_row(r, s.metas[r], None)
                s.ids[r] = s.docs[r] = s.metas[r] = None
                del s.row_of[i]
            if s.journal is not None:
                s.journal.log("delete", gone, None, None, None)
            return gone

    # ------------------------------------------------------------------ reads
    def count(self) -> int:
        """Live records; O(1).  Reference: api/routes/system.py:33."""
        s = self._s
        return 0 if s.store is None else s.store.count()

    def _mask_slot(self, s: _CollectionState, where, where_document) -> int:
        """Resolve a predicate to a pinned device mask slot (compile + upload on a miss; a hit costs a
        dictionary look-up: writes patch the cached bitmaps in place, see _patch_masks)."""
        key = json.dumps([where or None, where_document or None], sort_keys=True, ensure_ascii=False, default=str)
        with s.mask_mu:
            slot = s.mask_slots.get(key)
            if slot is not None and s.mask_valid.get(slot):
                s.mask_pins[slot] = s.mask_pins.get(slot, 0) + 1
                s.mask_lru.remove(key)
                s.mask_lru.append(key)
                return slot
            if slot is None:
                used = set(s.mask_slots.values())
                free = [i for i in range(N.MAX_MASK_SLOTS) if i not in used]
                if free:
                    slot = free[0]
                else:
                    victim = next((k_ for k_ in s.mask_lru if s.mask_pins.get(s.mask_slots[k_], 0) == 0), None)
                    if victim is None:
                        raise RuntimeError("all filter slots are in use by concurrent queries")
                    slot = s.mask_slots.pop(victim)
                    s.mask_pred.pop(victim, None)
                    s.mask_lru.remove(victim)
            elif s.mask_pins.get(slot, 0) != 0:
                raise RuntimeError("filter slot busy")   # cannot happen: writers exclude readers
            n = s.n_rows()
            passing = s.columns.evaluate(where, n) if where else np.ones(n, dtype=bool)
            if where_document:
                passing &= evaluate_where_document(where_document, s.docs, n)
            s.store.set_mask(slot, passing)
            s.mask_uploads += 1
            s.mask_slots[key] = slot
            s.mask_pred[key] = (where or None, where_document or None)
            s.mask_valid[slot] = True
            if key in s.mask_lru:
                s.mask_lru.remove(key)
            s.mask_lru.append(key)
            s.mask_pins[slot] = s.mask_pins.get(slot, 0) + 1
            return slot

    def _unpin(self, s, slot):
        with s.mask_mu:
            s.mask_pins[slot] -= 1

    def query(self, query_embeddings=None, query_texts=None, query_images=None, query_uris=None,
              n_results: int = 10, where=None, where_document=None, include=DEFAULT_INCLUDE,
              regime: str = "auto") -> Dict[str, Any]:
        """Exact top-n_results per query.  Reference: api/app.py:544-549,
        scripts/query_local.py:29-34.  Returns Chroma's columnar dict: lists (one
        per query) of lists ascending by distance, each of length
        min(n_results, live records passing the filter)."""
        include = list(include)
        for inc in include:
            if inc not in _VALID_INCLUDE:
                raise ValueError(f"Expected include item to be one of {sorted(_VALID_INCLUDE)}, got {inc}")
        if not isinstance(n_results, (int, np.integer)) or isinstance(n_results, bool) or n_results <= 0:
            raise ValueError(f"Number of requested results {n_results!r} must be a positive integer")
        validate_where(where)
        validate_where_document(where_document)
        if query_embeddings is None:
            if query_texts is None:
                raise ValueError("You must provide one of query_embeddings or query_texts")
            q = self._embed(self._as_list(query_texts))
        else:
            q = np.asarray(query_embeddings, dtype=np.float32)
            if q.ndim == 1:
                q = q[None, :]
            if q.ndim != 2 or q.shape[0] == 0:
                raise ValueError("Expected query_embeddings to be a non-empty list of embeddings")
        B = q.shape[0]
        s = self._s
        with _Shared(s.rw):
            cnt = self.count()
            out: Dict[str, Any] = {"ids": [[] for _ in range(B)], "distances": None, "metadatas": None,
                                   "embeddings": None, "documents": None, "uris": None, "data": None,
                                   "included": include}
            for col in ("distances", "metadatas", "documents", "embeddings"):
                if col in include:
                    out[col] = [[] for _ in range(B)]
            if s.store is None or cnt == 0:
                return out
            if q.shape[1] != s.dim:
                raise ValueError(f"Embedding dimension {q.shape[1]} does not match collection dimensionality {s.dim}")
            k = int(n_results)
            if k > cnt:
                logger.warning("Number of requested results %d is greater than number of elements in index %d, "
                               "updating n_results = %d", k, cnt, cnt)
                k = cnt
            if k > N.MAX_K:
                raise ValueError(f"n_results is limited to {N.MAX_K} by the device top-k kernels")
            slot = -1
            if where or where_document:
                slot = self._mask_slot(s, where, where_document)
            try:
                rows, dists, counts = s.store.query(q, k, slot, regime)
            finally:
                if slot >= 0:
                    self._unpin(s, slot)
            # ---- result assembly (vectorised gathers over object arrays) ----
            for b in range(B):
                r = rows[b, :counts[b]]
                out["ids"][b] = s.ids[r].tolist()
                if out["distances"] is not None:
                    out["distances"][b] = dists[b, :counts[b]].astype(np.float64).tolist()
                if out["metadatas"] is not None:
                    out["metadatas"][b] = s.metas[r].tolist()
                if out["documents"] is not None:
                    out["documents"][b] = s.docs[r].tolist()
                if out["embeddings"] is not None:
                    out["embeddings"][b] = s.store.fetch(r, exact=True).tolist() if r.size else []
            return out

    def get(self, ids=None, where=None, limit=None, offset=None, where_document=None,
            include=("metadatas", "documents")) -> Dict[str, Any]:
        """Fetch records by id and/or predicate, in insertion (row) order."""
        include = list(include)
        validate_where(where)
        validate_where_document(where_document)
        ids = self._as_list(ids)
        s = self._s
        with _Shared(s.rw):
            n = s.n_rows()
            if ids is not None:
                rows = np.array(sorted(s.row_of[i] for i in set(ids) if i in s.row_of), dtype=np.int64)
            else:
                rows = np.array(sorted(s.row_of.values()), dtype=np.int64)
            if rows.size and (where or where_document):
                ok = np.ones(n, dtype=bool)
                if where:
                    ok &= s.columns.evaluate(where, n)
                if where_document:
                    ok &= evaluate_where_document(where_document, s.docs, n)
                rows = rows[ok[rows]]
            if offset:
                rows = rows[int(offset):]
            if limit is not None:
                rows = rows[:int(limit)]
            out = {"ids": s.ids[rows].tolist() if rows.size else [], "embeddings": None, "metadatas": None,
                   "documents": None, "uris": None, "data": None, "included": include}
            if "metadatas" in include:
                out["metadatas"] = s.metas[rows].tolist() if rows.size else []
            if "documents" in include:
                out["documents"] = s.docs[rows].tolist() if rows.size else []
            if "embeddings" in include:
                out["embeddings"] = s.store.fetch(rows, exact=True).tolist() if rows.size else []
            return out

    def peek(self, limit: int = 10):
        return self.get(limit=limit, include=("metadatas", "documents", "embeddings"))

    def modify(self, name=None, metadata=None):
        if metadata is not None:
            self._s.metadata = dict(metadata)
            self.metadata = self._s.metadata

    # ------------------------------------------------------------ engine access
    @property
    def device_store(self) -> Optional[DeviceStore]:
        return self._s.store

    def rows_of(self, ids: Sequence[str]) -> List[int]:
        return [self._s.row_of.get(i, -1) for i in ids]


# ------------------------------------------------------------------------------
# clients
# ------------------------------------------------------------------------------
_REGISTRY: Dict[tuple, _CollectionState] = {}
_REGISTRY_MU = threading.Lock()


class Client:
    """In-process client.  `path=None` gives an ephemeral namespace."""

    def __init__(self, path: Optional[str] = None, settings=None, tenant: str = "default_tenant",
                 database: str = "default_database"):
        self._path = os.path.abspath(path) if path is not None else f"<ephemeral:{id(self)}>"
        self._persistent = path is not None

    def _key(self, name):
        return (self._path, name)

    def heartbeat(self) -> int:
        import time
        return int(time.time_ns())

    def get_or_create_collection(self, name: str, metadata: Optional[dict] = None,
                                 embedding_function=None, **_ignored) -> Collection:
        with _REGISTRY_MU:
            st = _REGISTRY.get(self._key(name))
            if st is None:
                st = _CollectionState(name, metadata, self._path if self._persistent else None)
                if self._persistent:
                    from .persist import attach_journal
                    attach_journal(st, self._path)
                _REGISTRY[self._key(name)] = st
        return Collection(st, embedding_function)

    def create_collection(self, name: str, metadata: Optional[dict] = None, embedding_function=None,
                          get_or_create: bool = False, **_ignored) -> Collection:
        with _REGISTRY_MU:
            exists = self._key(name) in _REGISTRY
        if exists and not get_or_create:
            raise ValueError(f"Collection {name} already exists")
        return self.get_or_create_collection(name, metadata, embedding_function)

    def get_collection(self, name: str, embedding_function=None, **_ignored) -> Collection:
        with _REGISTRY_MU:
            st = _REGISTRY.get(self._key(name))
        if st is None:
            if self._persistent:
                from .persist import collection_exists_on_disk
                if collection_exists_on_disk(self._path, name):
                    return self.get_or_create_collection(name, None, embedding_function)
            raise ValueError(f"Collection {name} does not exist.")
        return Collection(st, embedding_function)

    def list_collections(self):
        with _REGISTRY_MU:
            return [Collection(st) for (p, _), st in _REGISTRY.items() if p == self._path]

    def delete_collection(self, name: str):
        with _REGISTRY_MU:
            st = _REGISTRY.pop(self._key(name), None)
        if st is None:
            raise ValueError(f"Collection {name} does not exist.")
        if st.journal is not None:
            st.journal.drop()
        if st.store is not None:
            st.store.close()

    def reset(self):
        with _REGISTRY_MU:
            for key in [k for k in _REGISTRY if k[0] == self._path]:
                st = _REGISTRY.pop(key)
                if st.store is not None:
                    st.store.close()
        return True


class PersistentClient(Client):
    """chromadb.PersistentClient(path=...) -- api/app.py:89.  Every client
    opened on the same path shares collection state within the process."""

    def __init__(self, path: str = "./chroma", settings=None, tenant: str = "default_tenant",
                 database: str = "default_database"):
        super().__init__(path=path, settings=settings)


class EphemeralClient(Client):
    def __init__(self, settings=None, tenant: str = "default_tenant", database: str = "default_database"):
        super().__init__(path=None, settings=settings)


def _reset_registry_for_tests():
    with _REGISTRY_MU:
        for st in _REGISTRY.values():
            if st.store is not None:
                st.store.close()
            if st.journal is not None:
                st.journal.close()
        _REGISTRY.clear()
