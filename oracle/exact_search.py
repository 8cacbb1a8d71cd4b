"""CPU oracle for the dense-retrieval hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import this module.  The product (local-rag-system_b200/)
never does: it fails loudly when the CUDA library is missing.

What it restates
----------------
The reference (akak0487521/Local-RAG-System) has no retrieval arithmetic of
its own: every call goes into the un-vendored dependency chromadb==0.5.3
(requirements.txt:6; brings chroma-hnswlib 0.7.3).  That package is absent
from /root/reference and cannot be installed here (no network, no wheel), so
this file restates its *published* algorithm for the path and anchors it on
the reference's own call sites:

  collection.query  api/app.py:544-549, scripts/query_local.py:29-34
  collection.add    api/app.py:221, scripts/ingest_docs_to_chroma.py:31
  collection.upsert scripts/build_index.py:92-96, scripts/bulk_import.py:66-70
  collection.delete api/app.py:269, 306, 311
  collection.count  api/routes/system.py:33

Distance definitions (hnswlib spaces as Chroma exposes them via `hnsw:space`):
  l2      d = sum((a-b)^2)            (squared, no sqrt)   <- the reference's space
  ip      d = 1 - sum(a*b)
  cosine  d = 1 - sum(a*b) / (|a| |b|)
Results are the k smallest distances, ascending; ties broken by insertion
(row) order.  At the reference's shipped scale (25 rows < batch_size 100)
Chroma itself answers by exact brute force, so exact search IS the reference
behaviour there.

PARITY STATUS: **parity unpinned for distances/ranking** -- no reference test
pins a retrieval result (tests/test_kb_crud.py mocks Chroma) and Chroma cannot
run here.  What IS pinned (tests/test_oracle_golden.py):
  * upsert-replace semantics against Chroma's own materialised metadata
    segment shipped in vector_store/chroma.sqlite3 (37 WAL records -> 25 ids,
    last write wins) -- a genuine reference output;
  * the known answers SURVEY.md 8c lists for the shipped vectors.
"""
from __future__ import annotations

import numpy as np

SPACES = ("l2", "cosine", "ip")


# --------------------------------------------------------------------------
# numeric helpers
# --------------------------------------------------------------------------
def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round-to-nearest-even) -> fp32.  Same rounding the device
    store applies when a collection is created with dtype bf16."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return (rounded & 0xFFFFFFFF).astype(np.uint32).view(np.float32).reshape(x.shape)


def normalise_rows(x: np.ndarray) -> np.ndarray:
    """Row L2-normalisation in fp32 with an fp32 sum of squares (what hnswlib's
    cosine space does on insert and on query).  Zero rows stay zero."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n2 = np.einsum("ij,ij->i", x, x, dtype=np.float32)
    inv = np.where(n2 > 0, 1.0 / np.sqrt(np.maximum(n2, np.float32(1e-30))), 0.0).astype(np.float32)
    return (x * inv[:, None]).astype(np.float32)


def prepare_corpus(space: str, x: np.ndarray, dtype: str = "f32") -> np.ndarray:
    """The values the store holds for rows `x`: normalised for cosine, then
    rounded to bf16 if the collection stores bf16 (returned upcast to fp32)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if space == "cosine":
        x = normalise_rows(x)
    if dtype == "bf16":
        x = round_to_bf16(x)
    return x


def prepare_queries(space: str, q: np.ndarray, dtype: str = "f32") -> np.ndarray:
    """Queries get exactly the same treatment as corpus rows."""
    return prepare_corpus(space, q, dtype)


def distance_matrix(space: str, q: np.ndarray, x: np.ndarray, acc=np.float64) -> np.ndarray:
    """[B,N] distances between prepared queries and prepared corpus rows,
    accumulated in `acc` (fp64 by default: the exact reference)."""
    if space not in SPACES:
        raise ValueError(f"unknown space {space!r}")
    q = q.astype(acc)
    x = x.astype(acc)
    if space == "l2":
        # sum((a-b)^2) evaluated directly (Chroma brute force: norm(x-y)**2)
        out = np.empty((q.shape[0], x.shape[0]), dtype=acc)
        for i in range(q.shape[0]):
            d = x - q[i]
            out[i] = np.einsum("ij,ij->i", d, d)
        return out
    return 1.0 - q @ x.T


def topk_stable(dist: np.ndarray, k: int, valid: np.ndarray | None = None):
    """k smallest of each row of `dist`, ascending, ties by row index.
    `valid` is a boolean [N] mask of rows that may be returned.
    Returns a list (one per query) of (rows int64[], dists float[])."""
    out = []
    n = dist.shape[1]
    cand = np.arange(n, dtype=np.int64) if valid is None else np.nonzero(valid)[0].astype(np.int64)
    for i in range(dist.shape[0]):
        d = dist[i, cand]
        order = np.lexsort((cand, d))[:k]
        out.append((cand[order], d[order]))
    return out


def exact_search(space: str, queries: np.ndarray, corpus: np.ndarray, k: int,
                 valid: np.ndarray | None = None, dtype: str = "f32",
                 acc=np.float64, chunk: int = 262144):
    """Exact brute-force top-k, chunked over corpus rows so that large corpora
    fit in memory.  Returns (rows [B,k'] int64, dists [B,k'] float64) lists."""
    q = prepare_queries(space, np.atleast_2d(queries), dtype)
    n = corpus.shape[0]
    best_rows = [np.empty(0, np.int64) for _ in range(q.shape[0])]
    best_d = [np.empty(0, acc) for _ in range(q.shape[0])]
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = prepare_corpus(space, corpus[s:e], dtype)
        d = distance_matrix(space, q, x, acc)
        v = None if valid is None else valid[s:e]
        part = topk_stable(d, k, v)
        for i, (r, dd) in enumerate(part):
            rows = np.concatenate([best_rows[i], r + s])
            ds = np.concatenate([best_d[i], dd])
            order = np.lexsort((rows, ds))[:k]
            best_rows[i], best_d[i] = rows[order], ds[order]
    return best_rows, best_d


def fast_topk_f32(space: str, q: np.ndarray, x: np.ndarray, k: int, chunk: int = 131072):
    """BLAS-threaded fp32 exact search (Q @ X.T + argpartition).  This is the
    'numpy exact' CPU baseline BASELINE.md names; used by bench.py's
    cpu_baseline / --impl reference legs, and checked against exact_search in
    tests.  `q`, `x` must already be prepared (normalised / rounded)."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    B = q.shape[0]
    rows_acc = np.empty((B, 0), np.int64)
    d_acc = np.empty((B, 0), np.float32)
    qn = np.einsum("ij,ij->i", q, q) if space == "l2" else None
    for s in range(0, x.shape[0], chunk):
        xc = x[s:s + chunk]
        dots = q @ xc.T
        if space == "l2":
            d = qn[:, None] + np.einsum("ij,ij->i", xc, xc)[None, :] - 2.0 * dots
        else:
            d = 1.0 - dots
        kk = min(k, d.shape[1])
        if kk < d.shape[1]:
            idx = np.argpartition(d, kk - 1, axis=1)[:, :kk]
        else:
            idx = np.broadcast_to(np.arange(d.shape[1]), d.shape).copy()
        rows_acc = np.concatenate([rows_acc, idx + s], axis=1)
        d_acc = np.concatenate([d_acc, np.take_along_axis(d, idx, axis=1)], axis=1)
        if rows_acc.shape[1] > 4 * k:
            o = np.argsort(d_acc, axis=1, kind="stable")[:, :k]
            rows_acc = np.take_along_axis(rows_acc, o, axis=1)
            d_acc = np.take_along_axis(d_acc, o, axis=1)
    o = np.lexsort((rows_acc, d_acc), axis=1)[:, :k]
    return np.take_along_axis(rows_acc, o, axis=1), np.take_along_axis(d_acc, o, axis=1)


# --------------------------------------------------------------------------
# `where` / `where_document` evaluation -- one record at a time, plain Python
# (independent of the vectorised compiler the product uses)
# --------------------------------------------------------------------------
_CMP = ("$eq", "$ne", "$gt", "$gte", "$lt", "$lte", "$in", "$nin")


def _same_type(a, b) -> bool:
    # Chroma stores str / int / float / bool in separate typed columns and a
    # predicate only ever looks at the column of its operand's type.
    if isinstance(a, bool) or isinstance(b, bool):
        return isinstance(a, bool) and isinstance(b, bool)
    return type(a) is type(b)


def _cmp(op: str, have, want) -> bool:
    present = have is not None
    if op == "$eq":
        return present and _same_type(have, want) and have == want
    if op == "$ne":   # rows lacking the key (or holding another type) do match
        return not (present and _same_type(have, want) and have == want)
    if op == "$in":
        return present and any(_same_type(have, w) and have == w for w in want)
    if op == "$nin":
        return not (present and any(_same_type(have, w) and have == w for w in want))
    if isinstance(want, (bool, str)) or not isinstance(want, (int, float)):
        raise ValueError(f"Expected operand of {op} to be an int or a float, got {want!r}")
    if not present or isinstance(have, (bool, str)) or not _same_type(have, want):
        return False
    return {"$gt": have > want, "$gte": have >= want, "$lt": have < want, "$lte": have <= want}[op]


def where_matches(where, meta: dict) -> bool:
    """Chroma `where` grammar: {k: v} (implicit $eq), {k: {$op: v}},
    {"$and": [...]}, {"$or": [...]}.  Deviation kept from SURVEY.md 8b: a dict
    with several keys is an implicit $and (api/app.py:540-542 builds one)."""
    if not where:
        return True
    meta = meta or {}
    ok = True
    for key, cond in where.items():
        if key == "$and":
            ok = ok and all(where_matches(w, meta) for w in cond)
        elif key == "$or":
            ok = ok and any(where_matches(w, meta) for w in cond)
        elif isinstance(cond, dict):
            for op, want in cond.items():
                if op not in _CMP:
                    raise ValueError(f"unknown where operator {op!r}")
                ok = ok and _cmp(op, meta.get(key), want)
        else:
            ok = ok and _cmp("$eq", meta.get(key), cond)
    return ok


def where_document_matches(wd, doc) -> bool:
    if not wd:
        return True
    ok = True
    for key, cond in wd.items():
        if key == "$and":
            ok = ok and all(where_document_matches(w, doc) for w in cond)
        elif key == "$or":
            ok = ok and any(where_document_matches(w, doc) for w in cond)
        elif key == "$contains":
            ok = ok and (doc is not None and cond in doc)
        elif key == "$not_contains":
            ok = ok and not (doc is not None and cond in doc)
        else:
            raise ValueError(f"unknown where_document operator {key!r}")
    return ok


# --------------------------------------------------------------------------
# collection model: the add / upsert / delete / query / get / count semantics
# --------------------------------------------------------------------------
class OracleCollection:
    """Dict-and-list model of a Chroma collection (semantics per SURVEY.md 8b).
    Rows are numbered in insertion order; a deleted row's number is never
    reused here (the product may reuse rows; ties are therefore compared by
    distance only in the parity tests)."""

    def __init__(self, space: str = "l2", dtype: str = "f32"):
        assert space in SPACES
        self.space, self.dtype = space, dtype
        self.ids: list = []          # row -> id (None when deleted)
        self.vecs: list = []         # row -> raw fp32 vector
        self.metas: list = []
        self.docs: list = []
        self.row_of: dict = {}
        self.dim = None

    # -- writes ------------------------------------------------------------
    def _check(self, ids, embeddings):
        if len(set(ids)) != len(ids):
            raise ValueError("duplicate ids in one call")
        emb = np.asarray(embeddings, dtype=np.float32)
        if emb.ndim != 2 or emb.shape[0] != len(ids):
            raise ValueError("embeddings must be [len(ids), dim]")
        if self.dim is None:
            self.dim = emb.shape[1]
        if emb.shape[1] != self.dim:
            raise ValueError(f"dimension {emb.shape[1]} != collection dimension {self.dim}")
        return emb

    def add(self, ids, embeddings, metadatas=None, documents=None):
        emb = self._check(ids, embeddings)
        for i, id_ in enumerate(ids):
            if id_ in self.row_of:      # existing id: skipped (Chroma warns)
                continue
            self.row_of[id_] = len(self.ids)
            self.ids.append(id_)
            self.vecs.append(emb[i].copy())
            self.metas.append(dict(metadatas[i]) if metadatas and metadatas[i] else None)
            self.docs.append(documents[i] if documents else None)

    def upsert(self, ids, embeddings, metadatas=None, documents=None):
        emb = self._check(ids, embeddings)
        for i, id_ in enumerate(ids):
            m = dict(metadatas[i]) if metadatas and metadatas[i] else None
            d = documents[i] if documents else None
            if id_ in self.row_of:
                # update in place.  Chroma's metadata segment handles an UPSERT record for an existing id
                # with _update_metadata: the keys given are inserted-or-replaced, keys not given stay, and
                # the document is just the key "chroma:document" -- so metadata MERGES and a call without
                # documents keeps the old one [dep: chromadb 0.5.3 segment/impl/metadata/sqlite.py].  The
                # shipped chroma.sqlite3 cannot tell merge from replace (its 12 re-upserts carry the same
                # key set), which tests/test_oracle_golden.py states.
                r = self.row_of[id_]
                if self.metas[r]:
                    m = {**self.metas[r], **(m or {})}
                if not documents:
                    d = self.docs[r]
                self.vecs[r], self.metas[r], self.docs[r] = emb[i].copy(), m, d
            else:
                self.row_of[id_] = len(self.ids)
                self.ids.append(id_)
                self.vecs.append(emb[i].copy())
                self.metas.append(m)
                self.docs.append(d)

    def delete(self, ids=None, where=None, where_document=None):
        if ids is None and not where and not where_document:
            raise ValueError("delete needs ids, where or where_document")
        victims = []
        for id_, r in self.row_of.items():
            if ids is not None and id_ not in ids:
                continue
            if where_matches(where, self.metas[r]) and where_document_matches(where_document, self.docs[r]):
                victims.append(id_)
        for id_ in victims:
            r = self.row_of.pop(id_)
            self.ids[r] = None
        return victims

    # -- reads -------------------------------------------------------------
    def count(self) -> int:
        return len(self.row_of)

    def vector_of(self, id_):
        return self.vecs[self.row_of[id_]]

    def _valid(self, where, where_document):
        v = np.zeros(len(self.ids), dtype=bool)
        for id_, r in self.row_of.items():
            v[r] = where_matches(where, self.metas[r]) and where_document_matches(where_document, self.docs[r])
        return v

    def query(self, query_embeddings, n_results=10, where=None, where_document=None, acc=np.float64):
        q = np.atleast_2d(np.asarray(query_embeddings, dtype=np.float32))
        if self.dim is not None and q.shape[1] != self.dim:
            raise ValueError("query dimension mismatch")
        out = {"ids": [], "distances": [], "metadatas": [], "documents": [], "rows": []}
        if not self.ids:
            for _ in range(q.shape[0]):
                for key in out:
                    out[key].append([])
            return out
        valid = self._valid(where, where_document)
        x = np.stack(self.vecs)
        rows, dists = exact_search(self.space, q, x, n_results, valid, self.dtype, acc)
        for r, d in zip(rows, dists):
            out["rows"].append([int(i) for i in r])
            out["ids"].append([self.ids[i] for i in r])
            out["distances"].append([float(v) for v in d])
            out["metadatas"].append([self.metas[i] for i in r])
            out["documents"].append([self.docs[i] for i in r])
        return out

    def get(self, ids=None, where=None, where_document=None, limit=None, offset=None):
        rows = []
        if ids is not None:
            rows = [self.row_of[i] for i in ids if i in self.row_of]
            rows.sort()
        else:
            rows = sorted(self.row_of.values())
        rows = [r for r in rows
                if where_matches(where, self.metas[r]) and where_document_matches(where_document, self.docs[r])]
        if offset:
            rows = rows[offset:]
        if limit is not None:
            rows = rows[:limit]
        return {"ids": [self.ids[r] for r in rows], "metadatas": [self.metas[r] for r in rows],
                "documents": [self.docs[r] for r in rows], "rows": rows}
