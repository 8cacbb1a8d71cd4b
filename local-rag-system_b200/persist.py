"""Persistence for PersistentClient(path): a write-ahead journal + importer for the
reference's on-disk Chroma format (SURVEY.md 5 "checkpoint/resume", 8f-2).

Chroma 0.5.3 appends every write to sqlite table `embeddings_queue`
(seq_id, operation, id, fp32 vector blob, JSON metadata incl. "chroma:document")
and replays it on start-up.  The reference ships such a file
(vector_store/chroma.sqlite3: 37 UPSERT records -> 25 live ids).  We do the
same with our own file `<path>/rag_b200.sqlite3`, and -- when that file does
not exist yet -- import `<path>/chroma.sqlite3` read-only so that
PersistentClient(path="./vector_store") serves the shipped index as is
(scripts/query_local.py, GET /health).  Chroma's file is never written to.
"""
from __future__ import annotations

import json
import logging
import os
import sqlite3
import threading

import numpy as np

logger = logging.getLogger("local_rag_system_b200")

JOURNAL_FILE = "rag_b200.sqlite3"
CHROMA_FILE = "chroma.sqlite3"
_CHROMA_OPS = {0: "add", 1: "update", 2: "upsert", 3: "delete"}


class Journal:
    def __init__(self, db_path: str, collection: str):
        self.db_path, self.collection = db_path, collection
        self._mu = threading.Lock()
        self._con = sqlite3.connect(db_path, check_same_thread=False)
        self._con.execute("PRAGMA journal_mode=WAL")
        self._con.execute("CREATE TABLE IF NOT EXISTS collections(name TEXT PRIMARY KEY, metadata TEXT)")
        # names dropped with delete_collection(): never re-import them from a Chroma file lying in the same directory
        self._con.execute("CREATE TABLE IF NOT EXISTS dropped(name TEXT PRIMARY KEY)")
        self._con.execute("CREATE TABLE IF NOT EXISTS wal(seq INTEGER PRIMARY KEY AUTOINCREMENT, collection TEXT, "
                          "op TEXT, id TEXT, vector BLOB, metadata TEXT, document TEXT)")
        self._con.commit()
        self.enabled = True

    def register(self, metadata):
        with self._mu:
            self._con.execute("INSERT OR IGNORE INTO collections(name, metadata) VALUES (?, ?)",
                              (self.collection, json.dumps(metadata) if metadata else None))
            self._con.commit()

    def was_dropped(self) -> bool:
        return self._con.execute("SELECT 1 FROM dropped WHERE name=?", (self.collection,)).fetchone() is not None

    def stored_metadata(self):
        row = self._con.execute("SELECT metadata FROM collections WHERE name=?", (self.collection,)).fetchone()
        return (json.loads(row[0]) if row and row[0] else None), row is not None

    def log(self, op, ids, vectors, metadatas, documents):
        """One transaction per API call, as Chroma does."""
        if not self.enabled:
            return
        recs = []
        for j, i in enumerate(ids):
            vec = None if vectors is None else np.ascontiguousarray(vectors[j], dtype="<f4").tobytes()
            md = None if not metadatas or metadatas[j] is None else json.dumps(metadatas[j], ensure_ascii=False)
            doc = None if not documents else documents[j]
            recs.append((self.collection, op, i, vec, md, doc))
        with self._mu:
            self._con.executemany("INSERT INTO wal(collection, op, id, vector, metadata, document) VALUES (?,?,?,?,?,?)", recs)
            self._con.commit()

    def records(self):
        cur = self._con.execute("SELECT op, id, vector, metadata, document FROM wal WHERE collection=? ORDER BY seq",
                                (self.collection,))
        for op, i, vec, md, doc in cur:
            yield (op, i, None if vec is None else np.frombuffer(vec, dtype="<f4"),
                   json.loads(md) if md else None, doc)

    def drop(self):
        with self._mu:
            self._con.execute("DELETE FROM wal WHERE collection=?", (self.collection,))
            self._con.execute("DELETE FROM collections WHERE name=?", (self.collection,))
            self._con.execute("INSERT OR IGNORE INTO dropped(name) VALUES (?)", (self.collection,))
            self._con.commit()
        self.close()

    def close(self):
        try:
            self._con.close()
        except Exception:
            pass


def read_chroma_wal(path: str, collection: str):
    """Yield (op, id, vector, metadata, document) from a Chroma 0.5.x sqlite file,
    plus the collection's metadata dict.  Read-only, immutable open."""
    db = os.path.join(path, CHROMA_FILE)
    con = sqlite3.connect(f"file:{db}?mode=ro&immutable=1", uri=True)
    try:
        row = con.execute("SELECT id FROM collections WHERE name=?", (collection,)).fetchone()
        if row is None:
            return None, []
        coll_id = row[0]
        meta = {}
        for key, s, i, f, b in con.execute(
                "SELECT key, str_value, int_value, float_value, bool_value FROM collection_metadata "
                "WHERE collection_id=?", (coll_id,)):
            meta[key] = s if s is not None else i if i is not None else f if f is not None else bool(b)
        recs = []
        for op, eid, blob, enc, md in con.execute(
                "SELECT operation, id, vector, encoding, metadata FROM embeddings_queue WHERE topic LIKE ? ORDER BY seq_id",
                (f"%{coll_id}",)):
            vec = None
            if blob is not None:
                if enc not in (None, "FLOAT32"):
                    raise ValueError(f"unsupported Chroma vector encoding {enc!r}")
                vec = np.frombuffer(blob, dtype="<f4")
            m = json.loads(md) if md else None
            doc = m.pop("chroma:document", None) if m else None
            recs.append((_CHROMA_OPS[op], eid, vec, m or None, doc))
        return meta or None, recs
    finally:
        con.close()


def collection_exists_on_disk(path: str, name: str) -> bool:
    ours = os.path.join(path, JOURNAL_FILE)
    if os.path.exists(ours):
        con = sqlite3.connect(f"file:{ours}?mode=ro", uri=True)
        try:
            if con.execute("SELECT 1 FROM collections WHERE name=?", (name,)).fetchone():
                return True
        except sqlite3.Error:
            pass
        finally:
            con.close()
    if os.path.exists(os.path.join(path, CHROMA_FILE)):
        try:
            meta, recs = read_chroma_wal(path, name)
            return meta is not None or bool(recs)
        except sqlite3.Error:
            return False
    return False


def _replay(state, records):
    """Apply journal records through the normal Collection write path (journal off)."""
    from .collection import Collection
    col = Collection(state, None)
    batch_op, ids, vecs, metas, docs = None, [], [], [], []

    def flush():
        nonlocal batch_op, ids, vecs, metas, docs
        if not ids:
            return
        if batch_op == "delete":
            col.delete(ids=ids)
        elif batch_op == "add":
            col.add(ids=ids, embeddings=np.stack(vecs), metadatas=metas, documents=docs)
        else:
            col.upsert(ids=ids, embeddings=np.stack(vecs), metadatas=metas, documents=docs)
        batch_op, ids, vecs, metas, docs = None, [], [], [], []

    for op, i, vec, md, doc in records:
        if op == "update" and vec is None:
            # Chroma logs metadata / document-only updates without a vector: they go through
            # Collection.update (merge the keys, keep the stored vector), one at a time
            flush()
            col.update(ids=[i], metadatas=[md] if md else None, documents=[doc] if doc is not None else None)
            continue
        op = "upsert" if op == "update" else op
        if op != batch_op or i in ids or len(ids) >= 4096:
            flush()
            batch_op = op
        ids.append(i)
        vecs.append(vec)
        metas.append(md)
        docs.append(doc)
    flush()


# the keys that decide what the journal's vectors MEAN; a caller cannot change them for an existing collection
_PINNED_KEYS = ("hnsw:space", "b200:dtype")


def attach_journal(state, path: str):
    """Called under the registry lock when a collection is first opened on `path`."""
    ours = os.path.join(path, JOURNAL_FILE)
    journal = None
    try:
        os.makedirs(path, exist_ok=True)
        journal = Journal(ours, state.name)
    except (OSError, sqlite3.Error) as e:
        logger.warning("persist dir %s is not writable (%s): collection %s will be in-memory only", path, e, state.name)

    recs, known = [], False
    if journal is not None:
        stored_md, known = journal.stored_metadata()
        if known:
            if stored_md and not state.metadata:
                state.apply_metadata(stored_md)
            elif stored_md:
                # the journal was written under the stored space / dtype: replaying it under another one would
                # silently change every distance, so the stored values win (with a warning)
                md = dict(state.metadata)
                for key in _PINNED_KEYS:
                    if key in stored_md and md.get(key, stored_md[key]) != stored_md[key]:
                        logger.warning("collection %s exists with %s=%r; ignoring the requested %r", state.name, key,
                                       stored_md[key], md.get(key))
                    if key in stored_md:
                        md[key] = stored_md[key]
                state.apply_metadata(md)
            recs = list(journal.records())
    imported = False
    dropped = journal.was_dropped() if journal is not None else False
    if not known and not dropped and os.path.exists(os.path.join(path, CHROMA_FILE)):
        try:
            md, recs = read_chroma_wal(path, state.name)
            if md and not state.metadata:
                state.apply_metadata(md)
            imported = bool(recs)
        except (sqlite3.Error, ValueError) as e:
            logger.warning("could not import %s: %s", os.path.join(path, CHROMA_FILE), e)
            recs = []
    if recs:
        _replay(state, recs)
    if journal is not None:
        journal.register(state.metadata)
        if imported:   # carry the imported records into our own journal so later writes build on them
            journal.log("upsert", *_snapshot(state))
        state.journal = journal


def _snapshot(state):
    rows = sorted(state.row_of.values())
    if not rows:
        return [], None, None, None
    vec = state.store.fetch(np.array(rows, dtype=np.int64), exact=True)
    return ([state.ids[r] for r in rows], vec, [state.metas[r] for r in rows], [state.docs[r] for r in rows])
