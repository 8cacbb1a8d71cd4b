"""TEST DOUBLE for DeviceStore, built on the oracle (test infrastructure only).

Lets the CPU-only suite exercise the host layer (Collection semantics, `where`
compilation, result assembly, journal replay, the reference app through the
chromadb shim) in a container without a GPU.  It is injected by monkeypatching
`collection.DeviceStore` inside tests; the product never imports it."""
import numpy as np

from oracle.exact_search import exact_search, prepare_corpus


class FakeDeviceStore:
    def __init__(self, dim, dtype="f32", space="l2", device=0, capacity_hint=0, rerank=None):
        self.dim, self.dtype, self.space, self.device = dim, dtype, space, device
        self.rerank = False
        self.vec = np.zeros((0, dim), np.float32)
        self.live = np.zeros(0, bool)
        self.free = []
        self.masks = {}
        self.launches = 0

    def close(self):
        pass

    def count(self):
        return int(self.live.sum())

    def rows(self):
        return self.vec.shape[0]

    def upsert(self, vectors, rows=None):
        v = np.asarray(vectors, np.float32)
        out = np.empty(v.shape[0], np.int64)
        for i in range(v.shape[0]):
            r = -1 if rows is None else int(rows[i])
            if r < 0:
                while self.free:
                    c = self.free.pop()
                    if not self.live[c]:
                        r = c
                        break
                if r < 0:
                    r = self.vec.shape[0]
                    self.vec = np.concatenate([self.vec, np.zeros((1, self.dim), np.float32)])
                    self.live = np.concatenate([self.live, [False]])
            self.vec[r] = prepare_corpus(self.space, v[i:i + 1], self.dtype)[0]
            self.live[r] = True
            out[i] = r
        return out

    def delete(self, rows):
        for r in np.asarray(rows).reshape(-1).tolist():
            if 0 <= r < self.live.shape[0] and self.live[r]:
                self.live[r] = False
                self.free.append(r)

    def fetch(self, rows, exact=False):
        return self.vec[np.asarray(rows, np.int64)].copy()

    def flush(self):
        pass

    def patch_mask(self, slot, rows, passing):
        m = self.masks[slot]
        rows = np.asarray(rows, np.int64).reshape(-1)
        if rows.size and rows.max() >= m.shape[0]:
            m = np.concatenate([m, np.zeros(int(rows.max()) + 1 - m.shape[0], bool)])
        m[rows] = np.asarray(passing).astype(bool)
        self.masks[slot] = m

    def set_mask(self, slot, passing):
        self.masks[slot] = np.asarray(passing, bool).copy()

    def clear_mask(self, slot):
        self.masks.pop(slot, None)

    def query(self, queries, k, mask_slot=-1, regime="auto"):
        q = np.atleast_2d(np.asarray(queries, np.float32))
        valid = self.live.copy()
        if mask_slot >= 0:
            m = self.masks[mask_slot]
            mm = np.zeros_like(valid)
            mm[:min(len(m), len(valid))] = m[:len(valid)]
            valid &= mm
        # stored rows are already prepared; the space's normalisation is idempotent on them
        rows, dists = exact_search(self.space, q, self.vec, k, valid, self.dtype)
        B = q.shape[0]
        R = np.full((B, k), -1, np.int64)
        D = np.full((B, k), np.inf, np.float32)
        C = np.zeros(B, np.int32)
        for b in range(B):
            n = len(rows[b])
            R[b, :n], D[b, :n], C[b] = rows[b], dists[b], n
        return R, D, C
