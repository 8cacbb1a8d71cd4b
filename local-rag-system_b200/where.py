"""`where` / `where_document` compiler: Chroma filter grammar -> row bitmap.

Serves Collection.query(where=...) and Collection.delete(where=...)
(reference call sites api/app.py:540-548, 269, 311).  In Chroma the metadata
segment turns the predicate into SQL over typed value columns and hands the
vector index an allowed-id set (a Python callback per visited candidate in
hnswlib).  Here metadata lives in typed numpy columns (one per key, dictionary-
encoded strings) and a predicate is evaluated for all rows at once; the boolean
result is bit-packed and parked on the device, where the scan kernels AND it
with the live bitmap 32 rows at a time (K5 in SURVEY.md 2.4).

Grammar: {k: v} (implicit $eq) | {k: {$eq|$ne|$gt|$gte|$lt|$lte: v}} |
{k: {$in|$nin: [v..]}} | {"$and": [..]} | {"$or": [..]}.
A dict with several keys is an implicit $and -- a deliberate deviation from
chromadb>=0.4.16 (which rejects it), kept because api/app.py:540-542 builds
exactly such a dict for namespace + canonicality.
Typing follows Chroma's typed columns: a predicate only ever matches values
of its operand's own type (str / int / float / bool); $ne and $nin also match
rows that lack the key.
"""
from __future__ import annotations

import numpy as np

K_NONE, K_STR, K_INT, K_FLOAT, K_BOOL = 0, 1, 2, 3, 4
_RANGE_OPS = ("$gt", "$gte", "$lt", "$lte")
_OPS = ("$eq", "$ne", "$in", "$nin") + _RANGE_OPS


def kind_of(v) -> int:
    if isinstance(v, bool):
        return K_BOOL
    if isinstance(v, str):
        return K_STR
    if isinstance(v, (int, np.integer)):
        return K_INT
    if isinstance(v, (float, np.floating)):
        return K_FLOAT
    raise ValueError(f"Expected metadata value to be a str, int, float or bool, got {v!r}")


def validate_where(where):
    """Raise ValueError for anything outside the grammar (before touching data)."""
    if where is None:
        return
    if not isinstance(where, dict):
        raise ValueError(f"Expected where to be a dict, got {where!r}")
    for key, cond in where.items():
        if key in ("$and", "$or"):
            if not isinstance(cond, (list, tuple)) or len(cond) == 0:
                raise ValueError(f"Expected {key} to hold a non-empty list, got {cond!r}")
            for w in cond:
                validate_where(w)
        elif isinstance(key, str) and key.startswith("$"):
            raise ValueError(f"Expected where key to be a metadata key or $and/$or, got {key}")
        elif isinstance(cond, dict):
            if not cond:
                raise ValueError(f"Expected operator expression for {key!r}")
            for op, want in cond.items():
                if op not in _OPS:
                    raise ValueError(f"Expected where operator to be one of {_OPS}, got {op}")
                if op in ("$in", "$nin"):
                    if not isinstance(want, (list, tuple)) or len(want) == 0:
                        raise ValueError(f"Expected {op} operand to be a non-empty list, got {want!r}")
                    kinds = {kind_of(w) for w in want}
                    if len(kinds) != 1:
                        raise ValueError(f"Expected {op} operands to share one type, got {want!r}")
                elif op in _RANGE_OPS:
                    if kind_of(want) not in (K_INT, K_FLOAT):
                        raise ValueError(f"Expected operand of {op} to be an int or a float, got {want!r}")
                else:
                    kind_of(want)
        else:
            kind_of(cond)


def validate_where_document(wd):
    if wd is None:
        return
    if not isinstance(wd, dict):
        raise ValueError(f"Expected where_document to be a dict, got {wd!r}")
    for key, cond in wd.items():
        if key in ("$and", "$or"):
            if not isinstance(cond, (list, tuple)) or len(cond) == 0:
                raise ValueError(f"Expected {key} to hold a non-empty list")
            for w in cond:
                validate_where_document(w)
        elif key in ("$contains", "$not_contains"):
            if not isinstance(cond, str) or cond == "":
                raise ValueError(f"Expected {key} operand to be a non-empty str, got {cond!r}")
        else:
            raise ValueError(f"Expected where_document operator $contains/$not_contains/$and/$or, got {key}")


class _Column:
    """One metadata key over all rows: kind tag + typed value arrays."""

    def __init__(self):
        self.n = 0
        self.kind = np.zeros(0, dtype=np.uint8)
        self.ival = np.zeros(0, dtype=np.int64)     # int value, bool (0/1) or string code
        self.fval = np.zeros(0, dtype=np.float64)
        self.codes: dict = {}                       # string -> code

    def _grow(self, n):
        if n <= self.kind.shape[0]:
            return
        cap = max(n, 2 * self.kind.shape[0], 64)
        for name in ("kind", "ival", "fval"):
            old = getattr(self, name)
            new = np.zeros(cap, dtype=old.dtype)
            new[:old.shape[0]] = old
            setattr(self, name, new)

    def set(self, row: int, value):
        self._grow(row + 1)
        self.n = max(self.n, row + 1)
        k = kind_of(value)
        self.kind[row] = k
        if k == K_STR:
            code = self.codes.get(value)
            if code is None:
                code = len(self.codes)
                self.codes[value] = code
            self.ival[row] = code
        elif k == K_FLOAT:
            self.fval[row] = float(value)
        else:
            self.ival[row] = int(value)

    def clear(self, row: int):
        if row < self.n:
            self.kind[row] = K_NONE

    # ---- predicates, each returning bool[n_rows] ----
    def _pad(self, a, n_rows):
        a = a[:min(self.n, n_rows)]
        if a.shape[0] < n_rows:
            a = np.concatenate([a, np.zeros(n_rows - a.shape[0], dtype=bool)])
        return a

    def eq(self, want, n_rows):
        k = kind_of(want)
        kind, m = self.kind[:self.n], None
        if k == K_STR:
            code = self.codes.get(want)
            m = np.zeros(self.n, dtype=bool) if code is None else (kind == K_STR) & (self.ival[:self.n] == code)
        elif k == K_FLOAT:
            m = (kind == K_FLOAT) & (self.fval[:self.n] == float(want))
        else:
            m = (kind == k) & (self.ival[:self.n] == int(want))
        return self._pad(m, n_rows)

    def isin(self, wants, n_rows):
        m = np.zeros(n_rows, dtype=bool)
        for w in wants:
            m |= self.eq(w, n_rows)
        return m

    def rng(self, op, want, n_rows):
        k = kind_of(want)
        kind = self.kind[:self.n]
        vals = self.fval[:self.n] if k == K_FLOAT else self.ival[:self.n]
        w = float(want) if k == K_FLOAT else int(want)
        cmp = {"$gt": vals > w, "$gte": vals >= w, "$lt": vals < w, "$lte": vals <= w}[op]
        return self._pad((kind == k) & cmp, n_rows)


class MetadataColumns:
    """Columnar mirror of the per-row metadata dicts, kept in step by the
    collection on every add / upsert."""

    def __init__(self):
        self.cols: dict = {}

    def set_row(self, row: int, old_meta, new_meta):
        if old_meta:
            for key in old_meta:
                if not new_meta or key not in new_meta:
                    self.cols[key].clear(row)
        if new_meta:
            for key, v in new_meta.items():
                col = self.cols.get(key)
                if col is None:
                    col = self.cols[key] = _Column()
                col.set(row, v)

    def evaluate(self, where, n_rows: int) -> np.ndarray:
        """bool[n_rows]: rows whose metadata satisfies `where` (already validated)."""
        if not where:
            return np.ones(n_rows, dtype=bool)
        out = np.ones(n_rows, dtype=bool)
        for key, cond in where.items():
            if key == "$and":
                for w in cond:
                    out &= self.evaluate(w, n_rows)
            elif key == "$or":
                acc = np.zeros(n_rows, dtype=bool)
                for w in cond:
                    acc |= self.evaluate(w, n_rows)
                out &= acc
            else:
                col = self.cols.get(key)
                ops = cond.items() if isinstance(cond, dict) else (("$eq", cond),)
                for op, want in ops:
                    if col is None:
                        m = np.zeros(n_rows, dtype=bool)
                        positive = op not in ("$ne", "$nin")
                        out &= m if positive else ~m
                    elif op == "$eq":
                        out &= col.eq(want, n_rows)
                    elif op == "$ne":
                        out &= ~col.eq(want, n_rows)
                    elif op == "$in":
                        out &= col.isin(want, n_rows)
                    elif op == "$nin":
                        out &= ~col.isin(want, n_rows)
                    else:
                        out &= col.rng(op, want, n_rows)
        return out


def evaluate_where_document(wd, docs, n_rows: int) -> np.ndarray:
    """bool[n_rows] for a validated where_document over the row->document array."""
    if not wd:
        return np.ones(n_rows, dtype=bool)
    out = np.ones(n_rows, dtype=bool)
    for key, cond in wd.items():
        if key == "$and":
            for w in cond:
                out &= evaluate_where_document(w, docs, n_rows)
        elif key == "$or":
            acc = np.zeros(n_rows, dtype=bool)
            for w in cond:
                acc |= evaluate_where_document(w, docs, n_rows)
            out &= acc
        else:
            has = np.fromiter(((d is not None and cond in d) for d in docs[:n_rows]), dtype=bool, count=n_rows)
            out &= has if key == "$contains" else ~has
    return out


# ---- one record at a time (incremental mask maintenance) -----------------------------------------
def _same_kind(a, b) -> bool:
    try:
        return kind_of(a) == kind_of(b)
    except ValueError:
        return False


def match_record(where, meta) -> bool:
    """Does ONE record's metadata dict satisfy a validated `where`?  Same semantics as
    MetadataColumns.evaluate (typed comparison; $ne / $nin also match records lacking the key); used to
    patch the bits of the few rows a write touched instead of re-evaluating every row."""
    if not where:
        return True
    meta = meta or {}
    for key, cond in where.items():
        if key == "$and":
            if not all(match_record(w, meta) for w in cond):
                return False
        elif key == "$or":
            if not any(match_record(w, meta) for w in cond):
                return False
        else:
            have = meta.get(key)
            ops = cond.items() if isinstance(cond, dict) else (("$eq", cond),)
            for op, want in ops:
                if op in ("$eq", "$ne"):
                    hit = have is not None and _same_kind(have, want) and have == want
                    ok = hit if op == "$eq" else not hit
                elif op in ("$in", "$nin"):
                    hit = have is not None and any(_same_kind(have, w) and have == w for w in want)
                    ok = hit if op == "$in" else not hit
                else:
                    if have is None or not _same_kind(have, want):
                        ok = False
                    else:
                        ok = {"$gt": have > want, "$gte": have >= want, "$lt": have < want, "$lte": have <= want}[op]
                if not ok:
                    return False
    return True


def match_document(wd, doc) -> bool:
    if not wd:
        return True
    for key, cond in wd.items():
        if key == "$and":
            if not all(match_document(w, doc) for w in cond):
                return False
        elif key == "$or":
            if not any(match_document(w, doc) for w in cond):
                return False
        else:
            has = doc is not None and cond in doc
            if has != (key == "$contains"):
                return False
    return True
