// C ABI of one device store (include/rag_b200.h): device-resident corpus + search.
// Host-side bookkeeping only; all arithmetic is in the kernels.  There is no CPU
// fallback: without a usable sm_100 device every entry point that would compute
// returns RAG_ENODEV / RAG_ECUDA.
//
// Ordering of reads and writes (no device-wide synchronisation on the write path):
//   * host side: a reader/writer lock.  Synchronous queries hold it shared for the whole call,
//     asynchronous ones (rag_store_query_dev, caller's stream) only while they launch.
//   * device side: writes run on the store's admin stream.  Before the first kernel of a write the
//     admin stream waits for a marker recorded (by the writer) on the stream of every asynchronous
//     reader context that launched since the previous write; behind the last kernel of a write `ev_write` is recorded and `write_seq` bumped.
//     A reader whose context has not yet seen that sequence number makes its stream wait for
//     `ev_write` before launching.  Rows, bitmaps and masks therefore never change under a running
//     search, and a search launched after a write returns sees it.
//   * small writes (<= 64 rows) are parked in a pinned host block and reach the device as ONE copy +
//     ONE launch when the next read arrives or the block is full (PendingWrites).
#include "../../include/rag_b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "store_internal.h"
#include "tensor_regime.h"

using namespace rag;

// ------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int rag::fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ------------------------------------------------------------------------------
// per-caller scratch
// ------------------------------------------------------------------------------
int QueryCtx::ensure_tickets() {
  if (d_tickets) return RAG_OK;
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_tickets), kMaxTickets * sizeof(unsigned int)));
  CUDA_TRY(cudaMemsetAsync(d_tickets, 0, kMaxTickets * sizeof(unsigned int), stream));
  return RAG_OK;
}
int QueryCtx::ensure_host(size_t bytes) {
  if (bytes <= h_bytes) return RAG_OK;
  if (h_pin) {
    if (stream) cudaStreamSynchronize(stream);
    cudaFreeHost(h_pin);
  }
  h_pin = nullptr; h_bytes = 0;
  size_t want = align_up(std::max(bytes, (size_t)1 << 16), 4096);
  CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&h_pin), want));
  h_bytes = want;
  return RAG_OK;
}
int QueryCtx::ensure_dev(size_t bytes) {
  if (bytes <= d_bytes) return RAG_OK;
  if (d_buf) {
    if (stream) cudaStreamSynchronize(stream);
    cudaFree(d_buf);
  }
  d_buf = nullptr; d_bytes = 0;
  size_t want = align_up(std::max(bytes, (size_t)1 << 20), 1 << 20);
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_buf), want));
  d_bytes = want;
  return RAG_OK;
}
int QueryCtx::ensure_events() {
  if (!ev0) CUDA_TRY(cudaEventCreate(&ev0));
  if (!ev1) CUDA_TRY(cudaEventCreate(&ev1));
  if (!ev_done) CUDA_TRY(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
  return RAG_OK;
}
void QueryCtx::destroy() {
  if (stream && own_stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
  if (h_pin) cudaFreeHost(h_pin);
  if (d_buf) cudaFree(d_buf);
  if (d_tickets) cudaFree(d_tickets);
  if (h_redo) cudaFreeHost(h_redo);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (ev_done) cudaEventDestroy(ev_done);
  stream = nullptr; h_pin = nullptr; d_buf = nullptr; d_tickets = nullptr; h_redo = nullptr; ev0 = ev1 = ev_done = nullptr;
  h_bytes = d_bytes = 0;
}

namespace {

struct RdLock {
  pthread_rwlock_t* l;
  explicit RdLock(pthread_rwlock_t* x) : l(x) { pthread_rwlock_rdlock(l); }
  ~RdLock() { pthread_rwlock_unlock(l); }
};
struct WrLock {
  pthread_rwlock_t* l;
  explicit WrLock(pthread_rwlock_t* x) : l(x) { pthread_rwlock_wrlock(l); }
  ~WrLock() { pthread_rwlock_unlock(l); }
};

int acquire_ctx(rag_store* s, QueryCtx** out) {
  std::unique_lock<std::mutex> g(s->pool_mu);
  for (;;) {
    if (!s->pool_free.empty()) {
      *out = s->pool_free.back();
      s->pool_free.pop_back();
      return RAG_OK;
    }
    if (s->pool_created < rag_store::kMaxPool) {
      s->pool_created++;
      g.unlock();
      QueryCtx* c = new (std::nothrow) QueryCtx();
      if (!c) return fail(RAG_ENOMEM, "out of host memory");
      cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
      c->own_stream = (e == cudaSuccess);
      if (e != cudaSuccess || c->ensure_events() != RAG_OK) {
        c->destroy();
        delete c;
        g.lock();
        s->pool_created--;
        return fail(RAG_ECUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
      }
      *out = c;
      return RAG_OK;
    }
    s->pool_cv.wait(g);
  }
}

void release_ctx(rag_store* s, QueryCtx* c) {
  {
    std::lock_guard<std::mutex> g(s->pool_mu);
    s->pool_free.push_back(c);
  }
  s->pool_cv.notify_one();
}

struct CtxGuard {
  rag_store* s;
  QueryCtx* c;
  ~CtxGuard() { if (c) release_ctx(s, c); }
};

inline bool h_is_live(const rag_store* s, int64_t row) {
  return row >= 0 && row < s->rows && ((s->h_live[row >> 5] >> (row & 31)) & 1u);
}

// ---- write-side ordering (write lock held) -----------------------------------------------
// the admin stream waits for every asynchronous search launched since the previous write
int writer_wait_for_readers(rag_store* s) {
  std::lock_guard<std::mutex> lg(s->dev_mu);
  for (auto& kv : s->dev_ctx) {
    QueryCtx* c = kv.second;
    if (!c->launched) continue;
    c->launched = false;
    // the marker goes behind everything submitted to the reader's stream so far (no reader is launching:
    // the write lock is held).  Recording it here rather than after every search keeps back-to-back
    // searches free of stream markers, which would undo their programmatic overlap.
    if (c->ensure_events() != RAG_OK) continue;
    if (cudaEventRecord(c->ev_done, c->stream) != cudaSuccess) { (void)cudaGetLastError(); continue; }   // caller destroyed its stream
    CUDA_TRY(cudaStreamWaitEvent(s->admin.stream, c->ev_done, 0));
  }
  return RAG_OK;
}
// every later search is ordered behind what the admin stream holds now
int writer_publish(rag_store* s) {
  CUDA_TRY(cudaEventRecord(s->ev_write, s->admin.stream));
  s->write_seq.fetch_add(1, std::memory_order_release);
  return RAG_OK;
}

// grow device arrays to hold at least `need` rows (write lock held)
int grow(rag_store* s, int64_t need) {
  if (need <= s->capacity) return RAG_OK;
  if (need > 0xFFFFFFF0ll) return fail(RAG_EINVAL, "a store holds at most 2^32-16 rows");
  int64_t cap = std::max<int64_t>(need, std::max<int64_t>(1024, s->capacity * 2));
  cap = (cap + 127) / 128 * 128;       // whole 128-row tiles: the tensor regime reads all the live words of a tile
  void* nv = nullptr;
  float* nn = nullptr;
  float* nx = nullptr;
  uint32_t* nl = nullptr;
  uint32_t* nm[RAG_MAX_MASK_SLOTS] = {};
  const size_t per_row = s->row_bytes + (s->exact_elems ? (size_t)s->exact_elems * sizeof(float) : 0);
  auto alloc_rows = [&](int64_t c) -> cudaError_t {
    cudaError_t e = cudaMalloc(&nv, (size_t)c * s->row_bytes);
    if (e == cudaSuccess && s->exact_elems) {
      e = cudaMalloc(reinterpret_cast<void**>(&nx), (size_t)c * s->exact_elems * sizeof(float));
      if (e != cudaSuccess) { cudaFree(nv); nv = nullptr; }
    }
    if (e != cudaSuccess) (void)cudaGetLastError();
    return e;
  };
  cudaError_t e = alloc_rows(cap);
  if (e != cudaSuccess && cap > need) {   // doubling did not fit: take exactly what is needed
    cap = (need + 127) / 128 * 128;
    e = alloc_rows(cap);
  }
  if (e != cudaSuccess) return fail(RAG_ENOMEM, "cudaMalloc of %zu bytes for the corpus failed: %s", (size_t)cap * per_row, cudaGetErrorString(e));
  e = cudaMalloc(reinterpret_cast<void**>(&nn), (size_t)cap * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&nl), (size_t)cap / 32 * sizeof(uint32_t));
  for (int i = 0; e == cudaSuccess && i < RAG_MAX_MASK_SLOTS; ++i)
    if (s->d_masks[i]) e = cudaMalloc(reinterpret_cast<void**>(&nm[i]), (size_t)cap / 32 * sizeof(uint32_t));
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(nv); if (nx) cudaFree(nx); if (nn) cudaFree(nn); if (nl) cudaFree(nl);
    for (uint32_t* m : nm) if (m) cudaFree(m);
    return fail(RAG_ENOMEM, "cudaMalloc for side arrays failed: %s", cudaGetErrorString(e));
  }
  cudaStream_t st = s->admin.stream;
  int rc = writer_wait_for_readers(s);        // nobody may still read the old arrays when they are freed below
  if (rc != RAG_OK) return rc;
  CUDA_TRY(cudaMemsetAsync(nl, 0, (size_t)cap / 32 * sizeof(uint32_t), st));
  if (s->rows > 0) {
    CUDA_TRY(cudaMemcpyAsync(nv, s->d_vectors, (size_t)s->rows * s->row_bytes, cudaMemcpyDeviceToDevice, st));
    if (nx) CUDA_TRY(cudaMemcpyAsync(nx, s->d_exact, (size_t)s->rows * s->exact_elems * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(nn, s->d_norms2, (size_t)s->rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(nl, s->d_live, (size_t)((s->rows + 31) / 32) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  }
  for (int i = 0; i < RAG_MAX_MASK_SLOTS; ++i) {
    if (!nm[i]) continue;
    CUDA_TRY(cudaMemsetAsync(nm[i], 0, (size_t)cap / 32 * sizeof(uint32_t), st));
    if (s->mask_words[i] > 0)
      CUDA_TRY(cudaMemcpyAsync(nm[i], s->d_masks[i], (size_t)s->mask_words[i] * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  if (s->d_vectors) cudaFree(s->d_vectors);
  if (s->d_exact) cudaFree(s->d_exact);
  if (s->d_norms2) cudaFree(s->d_norms2);
  if (s->d_live) cudaFree(s->d_live);
  for (int i = 0; i < RAG_MAX_MASK_SLOTS; ++i)
    if (nm[i]) { cudaFree(s->d_masks[i]); s->d_masks[i] = nm[i]; }
  s->d_vectors = nv; s->d_exact = nx; s->d_norms2 = nn; s->d_live = nl;
  if (s->d_shadow) { cudaFree(s->d_shadow); s->d_shadow = nullptr; }
  s->capacity = cap;
  s->h_live.resize((size_t)cap / 32, 0u);
  return writer_publish(s);
}

void unclaim(rag_store* s, std::vector<int64_t>& claimed) {
  for (int64_t c : claimed) {
    s->h_live[(size_t)(c >> 5)] &= ~(1u << (c & 31));
    s->live--;
    s->free_rows.push_back(c);
  }
  claimed.clear();
}

// Assign destination rows for an upsert (write lock held).  Free rows handed out are marked live at once
// (so that a duplicate entry of the free list cannot hand the same row out twice) and recorded in
// `claimed`, which the caller gives back with unclaim() if the write fails.
int assign_rows(rag_store* s, int64_t n, const int64_t* rows, std::vector<int64_t>& dst, std::vector<int64_t>& claimed) {
  dst.resize((size_t)n);
  int64_t hwm = s->rows;
  for (int64_t i = 0; i < n; ++i) {
    int64_t r = rows ? rows[i] : -1;
    if (r >= 0) {
      if (r >= hwm) {
        if (s->external_rows) {
          hwm = r + 1;
        } else {
          unclaim(s, claimed);
          return fail(RAG_EINVAL, "upsert row %lld is beyond the store's %lld rows", (long long)r, (long long)hwm);
        }
      }
    } else {
      r = -1;
      while (!s->free_rows.empty()) {
        int64_t c = s->free_rows.back();
        s->free_rows.pop_back();
        if (c < s->rows && !h_is_live(s, c)) { r = c; break; }
      }
      if (r >= 0) {
        s->h_live[(size_t)(r >> 5)] |= 1u << (r & 31);
        s->live++;
        claimed.push_back(r);
      } else {
        r = hwm++;
      }
    }
    dst[(size_t)i] = r;
  }
  return RAG_OK;
}

bool contiguous(const int64_t* v, int64_t n) {
  for (int64_t i = 1; i < n; ++i)
    if (v[i] != v[0] + i) return false;
  return true;
}

void mark_live(rag_store* s, const std::vector<int64_t>& dst) {
  const int64_t n = (int64_t)dst.size();
  if (n > 64 && dst[0] >= s->rows && contiguous(dst.data(), n)) {     // bulk append: whole bitmap words at a time
    const int64_t lo = dst[0], hi = lo + n;
    for (int64_t r = lo; r < hi;) {
      const int64_t w = r >> 5, b = r & 31;
      const int64_t take = std::min<int64_t>(32 - b, hi - r);
      s->h_live[(size_t)w] |= (take == 32 ? 0xFFFFFFFFu : ((1u << take) - 1u) << b);
      r += take;
    }
    s->live += n;
    s->rows = hi;
    return;
  }
  for (int64_t r : dst) {
    if (r >= s->rows) s->rows = r + 1;
    uint32_t& w = s->h_live[(size_t)(r >> 5)];
    uint32_t bit = 1u << (r & 31);
    if (!(w & bit)) { w |= bit; s->live++; }
  }
}

// launch the upsert kernel for fp32 vectors already on the device (write lock held, admin stream).
// d_rows: destination rows on the device, or nullptr for the contiguous run starting at row0.
int launch_upsert_rows(rag_store* s, const float* d_src, int64_t n, const int64_t* d_rows, int64_t row0) {
  UpsertArgs a{};
  a.src = d_src;
  a.n = n;
  a.dim = s->dim;
  a.row_elems = s->row_elems;
  a.dtype = s->dtype;
  a.normalise = (s->space == RAG_SPACE_COSINE);
  a.vectors = s->d_vectors;
  a.norms2 = s->d_norms2;
  a.max_norm2 = s->d_max_norm2;
  a.live = s->d_live;
  a.exact = s->d_exact;
  a.exact_elems = s->exact_elems;
  a.shadow = s->d_shadow;
  a.shadow_kind = s->shadow_kind;
  a.lo_max2 = s->d_max_norm2 + 2;
  a.rows = d_rows;
  a.row0 = row0;
  CUDA_TRY(launch_upsert(a, s->admin.stream));
  s->launches++;
  return RAG_OK;
}

// pending small writes -> device: one copy, one launch (write lock held)
int flush_pending_locked(rag_store* s) {
  PendingWrites& p = s->pending;
  if (p.n == 0) return RAG_OK;
  NvtxRange nvtx("rag:flush_writes");
  QueryCtx& c = s->admin;
  const size_t row_in = (size_t)s->dim * sizeof(float);
  const size_t vec_b = align_up((size_t)p.n * row_in, 256);
  const size_t rows_b = (size_t)p.n * sizeof(int64_t);
  int rc = c.ensure_dev(vec_b + rows_b);
  if (rc != RAG_OK) return rc;
  rc = writer_wait_for_readers(s);
  if (rc != RAG_OK) return rc;
  // the parked vectors sit at the front of the pinned block, their destination rows behind the LAST slot:
  // move the row list next to the vectors so that one copy carries both
  int64_t* h_rows = reinterpret_cast<int64_t*>(p.h + (size_t)PendingWrites::kMaxRows * row_in);
  const bool contig = contiguous(h_rows, p.n);
  const int64_t row0 = h_rows[0];
  size_t bytes = (size_t)p.n * row_in;
  if (!contig) {
    memmove(p.h + vec_b, h_rows, rows_b);
    bytes = vec_b + rows_b;
  }
  const int64_t n = p.n;
  p.n = 0;
  p.slot_of.clear();
  s->pending_n.store(0, std::memory_order_release);
  CUDA_TRY(cudaMemcpyAsync(c.d_buf, p.h, bytes, cudaMemcpyHostToDevice, c.stream));
  CUDA_TRY(cudaEventRecord(p.ev_h2d, c.stream));
  p.in_flight = true;
  rc = launch_upsert_rows(s, reinterpret_cast<const float*>(c.d_buf), n,
                          contig ? nullptr : reinterpret_cast<const int64_t*>(c.d_buf + vec_b), row0);
  if (rc != RAG_OK) return rc;
  return writer_publish(s);
}

void fill_empty(int B, int k, int64_t* out_rows, float* out_dists, int32_t* out_counts) {
  const float inf = __builtin_inff();
  for (int64_t i = 0; i < (int64_t)B * k; ++i) {
    if (out_rows) out_rows[i] = -1;
    if (out_dists) out_dists[i] = inf;
  }
  if (out_counts) for (int b = 0; b < B; ++b) out_counts[b] = 0;
}

// fp32 store about to be searched by the tensor regime: make sure its bf16 shadow (of kind s->shadow_kind) exists
// (read lock held: no writer is active; concurrent readers serialise on shadow_mu)
int ensure_shadow(rag_store* s, cudaStream_t st) {
  std::lock_guard<std::mutex> g(s->shadow_mu);
  if (s->d_shadow) return RAG_OK;
  __nv_bfloat16* sh = nullptr;
  const size_t bytes = (size_t)s->capacity * (s->shadow_kind == kShadowHiLo ? 2 : 1) * s->row_elems * sizeof(__nv_bfloat16);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&sh), bytes);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(RAG_ENOMEM, "cudaMalloc of %zu bytes for the bf16 shadow failed: %s", bytes, cudaGetErrorString(e)); }
  e = cudaMemsetAsync(s->d_max_norm2 + 2, 0, sizeof(float), st);       // the bound restarts from the rows held now
  if (e == cudaSuccess)
    e = launch_split_rows(reinterpret_cast<const float*>(s->d_vectors), s->row_elems, 0, s->rows, sh, s->shadow_kind,
                          s->d_max_norm2 + 2, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { (void)cudaGetLastError(); cudaFree(sh); return fail(RAG_ECUDA, "building the bf16 shadow failed: %s", cudaGetErrorString(e)); }
  s->d_shadow = sh;
  s->launches++;
  return RAG_OK;
}

// statistics of the fp32 tensor regime: the re-run count of a context's previous batch has come back by now (it was
// copied behind that batch; a context runs one batch at a time).  Feeds the totals and the hi-only policy window.
void harvest_redo(rag_store* s, QueryCtx* c) {
  if (c->redo_batch <= 0 || c->h_redo == nullptr) return;
  const int seen = *reinterpret_cast<volatile int*>(c->h_redo);
  const int redo = seen < 0 ? 0 : (seen > c->redo_batch ? c->redo_batch : seen);
  s->f32_tensor_queries += c->redo_batch;
  s->f32_tensor_reruns += redo;
  s->last_batch_reruns.store(redo, std::memory_order_relaxed);
  if (c->redo_kind == kShadowHi && !s->shadow_pinned) {
    const int64_t q = (s->hi_window_queries += c->redo_batch);
    const int64_t r = (s->hi_window_reruns += redo);
    // more than 1/8 of >= 64 queries could not be certified: the rows are packed closer than the bf16 filter can tell
    // apart; the hi/lo split (error 1e-6 instead of 1e-3) serves such a store better.  Decided per window of 4096.
    if (q >= 64 && r * 8 > q && tensor::supported(s->dtype, s->row_elems, 1, s->space, 0, kShadowHiLo))
      s->want_shadow_kind.store(kShadowHiLo, std::memory_order_release);
    if (q >= 4096) { s->hi_window_queries = 0; s->hi_window_reruns = 0; }
  }
  c->redo_batch = 0;
}

// scratch of one search: prepared queries | per-CTA partial lists [grid_x][B][k_scan] | redo list |
// un-rounded queries | merged keys | tensor-regime scratch
struct ScratchLayout {
  size_t off_q, off_partial, off_redo, off_qexact, off_merged, off_tensor, total;
};
ScratchLayout scratch_layout(const rag_store* s, int B, int k, int grid_x) {
  const int ks = scan_k(s, k);
  ScratchLayout L{};
  size_t off = 0;
  L.off_q = off; off += align_up((size_t)B * s->row_elems * sizeof(float), 256);
  L.off_partial = off; off += align_up((size_t)grid_x * B * ks * sizeof(uint64_t), 256);
  L.off_redo = off; off += align_up((size_t)(B + 1) * sizeof(int), 256);      // redo count + list (split-precision tensor regime)
  L.off_qexact = off; off += align_up((size_t)B * std::max(s->exact_elems, 4) * sizeof(float), 256);
  L.off_merged = off; off += align_up((size_t)B * ks * sizeof(uint64_t), 256);
  // (fp32 stores: whichever shadow kind the store has or may move to)
  L.off_tensor = off;
  off += std::max(tensor::scratch_bytes(s->dtype, s->row_elems, B, k, s->sm_count, s->exact_elems ? 1 : 0, kShadowHi),
                  tensor::scratch_bytes(s->dtype, s->row_elems, B, k, s->sm_count, s->exact_elems ? 1 : 0, kShadowHiLo));
  L.total = off;
  return L;
}

}  // namespace

static bool fused_merge_enabled() {
  static const bool on = !(getenv("RAG_B200_FUSED_MERGE") && atoi(getenv("RAG_B200_FUSED_MERGE")) == 0);
  return on;
}

bool rag::direct_host_ok(const rag_store* s, int B, int k, int regime) {
  static const bool on = !(getenv("RAG_B200_DIRECT_HOST") && atoi(getenv("RAG_B200_DIRECT_HOST")) == 0);
  return on && regime == 1 && fused_merge_enabled() && (int64_t)B * k <= 4096 &&
         scan_stream_groups(B, s->dtype, s->row_elems, scan_k(s, k)) == 1;
}

int rag::wait_host_flag(const volatile uint32_t* flag, uint32_t seq, cudaStream_t st) {
  for (uint32_t spins = 1;; ++spins) {
    if (*flag == seq) { std::atomic_thread_fence(std::memory_order_acquire); return RAG_OK; }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
    if ((spins & 0x3FFFu) == 0) {          // now and then: has the stream died, or finished without the signal?
      const cudaError_t e = cudaStreamQuery(st);
      if (e == cudaSuccess) {
        if (*flag == seq) { std::atomic_thread_fence(std::memory_order_acquire); return RAG_OK; }
        return fail(RAG_ECUDA, "the search finished without raising its completion signal");
      }
      if (e != cudaErrorNotReady) { (void)cudaGetLastError(); return fail(RAG_ECUDA, "search failed: %s", cudaGetErrorString(e)); }
    }
  }
}

int rag::scan_k(const rag_store* s, int k) {
  if (!s->exact_elems) return k;
  return k <= 10 ? 16 : k + 16;        // slack for the exact re-ranking (same rule as tensor::candidates_kept)
}

size_t rag::search_scratch_bytes(const rag_store* s, int B, int k, int grid_x) {
  return scratch_layout(s, B, k, grid_x).total;
}

// Decide the kernel regime for a batch.
int rag::choose_regime(const rag_store* s, int B, int k, int flags) {
  if (flags == RAG_QUERY_FORCE_STREAM) return 1;
  const bool tensor_ok = tensor::supported(s->dtype, s->row_elems, k, s->space, s->exact_elems ? 1 : 0, s->shadow_kind);
  if (flags == RAG_QUERY_FORCE_TENSOR) return tensor_ok ? 2 : -1;
  if (!tensor_ok) return 1;
  // the stream kernel reads the corpus once per group of <= 8 queries; the tensor kernel
  // once per 128 (HBM-bound up to ~256 queries, tensor-bound beyond)
  if (s->dtype == RAG_DTYPE_F32) {
    if (B > tensor::kStreamMaxBatchF32) return 2;
    const bool big = (size_t)s->rows * s->row_bytes >= tensor::kF32TensorAlwaysBytes;
    return (big && k <= 16 && s->shadow_kind == kShadowHi) ? 2 : 1;       // half the bytes per query (tensor_regime.h)
  }
  return (B > tensor::kStreamMaxBatch) ? 2 : 1;
}

int rag::flush_if_pending(rag_store* s) {
  const int want = s->want_shadow_kind.load(std::memory_order_acquire);
  const bool switch_shadow = want != 0 && want != s->shadow_kind;
  if (s->pending_n.load(std::memory_order_acquire) == 0 && !switch_shadow) return RAG_OK;
  WrLock g(&s->lock);
  CUDA_TRY(cudaSetDevice(s->device));
  if (switch_shadow && want != s->shadow_kind) {
    // like a write: searches in flight may still read the old shadow; the next tensor-regime search rebuilds it
    int rc = writer_wait_for_readers(s);
    if (rc != RAG_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(s->admin.stream));
    if (s->d_shadow) { cudaFree(s->d_shadow); s->d_shadow = nullptr; }
    s->shadow_kind = want;
    rc = writer_publish(s);
    if (rc != RAG_OK) return rc;
  }
  return flush_pending_locked(s);
}

// shard-local search on device buffers.  Emits keys and/or rows; everything is asynchronous on c->stream.
int rag::search_device(rag_store* s, QueryCtx* c, unsigned char* scratch, int B, const float* d_queries_raw, int k,
                       int mask_slot, int regime, RowMap rows_map, const SearchOut& out, bool timed,
                       rag_exchange* xchg, uint32_t xchg_epoch, bool forced_tensor) {
  cudaStream_t st = c->stream;
  NvtxRange nvtx(regime == 2 ? "rag:search:tensor" : (xchg ? "rag:search:stream+exchange" : "rag:search:stream"));
  int rc = c->ensure_events();
  if (rc != RAG_OK) return rc;
  {   // order this stream behind the last write it has not seen yet
    const uint64_t seq = s->write_seq.load(std::memory_order_acquire);
    if (c->seen_write != seq) {
      CUDA_TRY(cudaStreamWaitEvent(st, s->ev_write, 0));
      c->seen_write = seq;
    }
  }
  const bool rerank = s->exact_elems != 0;
  const int ks = scan_k(s, k);
  const int grid_x = scan_stream_grid_x(s->sm_count, s->rows);
  const ScratchLayout L = scratch_layout(s, B, k, grid_x);
  float* d_q = reinterpret_cast<float*>(scratch + L.off_q);
  uint64_t* d_partial = reinterpret_cast<uint64_t*>(scratch + L.off_partial);
  int* d_redo = reinterpret_cast<int*>(scratch + L.off_redo);       // [0] count, [1..B] query indices
  float* d_qexact = reinterpret_cast<float*>(scratch + L.off_qexact);
  uint64_t* d_merged = reinterpret_cast<uint64_t*>(scratch + L.off_merged);
  unsigned char* d_tensor = scratch + L.off_tensor;

  const uint32_t* filter = nullptr;
  int64_t fwords = 0;
  if (mask_slot >= 0) {
    filter = s->d_masks[mask_slot];
    fwords = s->mask_words[mask_slot];
  }
  int launches = 0;
  int S = 0;
  tensor::Result tres{};
  // (a multi-shard front end decides the regime on ONE shard; another shard's shadow may have moved to a kind that
  // cannot take this k: answer exactly on the stream kernel rather than fail)
  if (regime == 2 && !forced_tensor &&
      !tensor::supported(s->dtype, s->row_elems, k, s->space, s->exact_elems ? 1 : 0, s->shadow_kind))
    regime = 1;
  if (regime == 2 && s->dtype == RAG_DTYPE_F32) {
    // the tensor regime contracts a bf16 shadow of the rows (half as many bytes again, or as many for the hi/lo
    // split).  If the device cannot hold it, an AUTO query is still answered -- exactly, by the stream kernel.
    harvest_redo(s, c);
    const int rcs = ensure_shadow(s, st);
    if (rcs == RAG_ENOMEM && !forced_tensor) regime = 1;
    else if (rcs != RAG_OK) return rcs;
  }
  // common arguments of the stream kernel
  ScanArgs sa{};
  sa.vectors = s->d_vectors; sa.dtype = s->dtype; sa.row_elems = s->row_elems;
  sa.cpr = (int)(s->row_bytes / 16);
  sa.n_rows = s->rows; sa.live = s->d_live; sa.filter = filter; sa.filter_words = fwords;
  sa.B = B; sa.l2 = (s->space == RAG_SPACE_L2);
  sa.grid_x = grid_x;
  sa.dim = s->dim; sa.normalise = (s->space == RAG_SPACE_COSINE);
  sa.rows_map = rows_map;
  sa.out_keys = out.keys; sa.out_rows = out.rows; sa.out_dists = out.dists; sa.out_counts = out.counts;
  const uint64_t* merge_src = nullptr;
  int k_lists = k;              // length of the candidate lists the generic tail below sees
  if (regime == 2) {
    if (timed) CUDA_TRY(cudaEventRecord(c->ev0, st));
    tensor::Problem p{};
    p.max_norm2 = s->d_max_norm2; p.lo_max2 = s->d_max_norm2 + 2;
    p.vectors = s->d_vectors; p.shadow = s->d_shadow; p.shadow_kind = (s->dtype == RAG_DTYPE_F32) ? s->shadow_kind : kShadowNone; p.norms2 = s->d_norms2; p.min_norm2 = s->d_max_norm2 + 1; p.n_rows = s->rows; p.row_elems = s->row_elems;
    p.dim = s->dim; p.dtype = s->dtype; p.space = s->space;
    p.live = s->d_live; p.filter = filter; p.filter_words = fwords;
    p.dense = (filter == nullptr && s->live == s->rows) ? 1 : 0;
    p.rerank = rerank ? 1 : 0; p.exact_elems = s->exact_elems;
    p.queries_raw = d_queries_raw; p.B = B; p.k = k;
    p.scratch = d_tensor; p.sm_count = s->sm_count;
    cudaError_t e = tensor::launch(p, st, &tres, &launches);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(RAG_ECUDA, "tensor-regime launch failed: %s", cudaGetErrorString(e)); }
    merge_src = tres.partial;
    S = tres.S;
    k_lists = tres.k_kept;
    d_merged = tres.merged;
    if (timed) CUDA_TRY(cudaEventRecord(c->ev1, st));
  } else {
    const bool fused = fused_merge_enabled();
    if (B > QueryCtx::kMaxTickets) return fail(RAG_EINVAL, "batch %d exceeds %d", B, QueryCtx::kMaxTickets);
    sa.k = ks; sa.k_out = k;
    sa.partial = d_partial;
    sa.round_bf16 = (s->dtype == RAG_DTYPE_BF16);
    if (fused) {
      rc = c->ensure_tickets();
      if (rc != RAG_OK) return rc;
      sa.done = c->d_tickets;
      sa.queries = nullptr; sa.queries_raw = d_queries_raw;
      if (rerank) { sa.exact = s->d_exact; sa.exact_elems = s->exact_elems; }
      if (out.query_flag != nullptr) { sa.query_flag = out.query_flag; sa.query_seq = out.query_seq; }
      if (out.done_flag != nullptr && direct_host_ok(s, B, k, 1)) {
        sa.done_flag = out.done_flag; sa.done_seq = out.done_seq;
        if (out.armed) *out.armed = true;
      }
    } else {          // unfused debugging mode: separate prep, merge and re-ranking kernels
      PrepArgs pa{};
      pa.src = d_queries_raw; pa.B = B; pa.dim = s->dim; pa.row_elems = s->row_elems;
      pa.normalise = (s->space == RAG_SPACE_COSINE);
      pa.round_bf16 = (s->dtype == RAG_DTYPE_BF16);
      pa.q_f32 = d_q; pa.q_bf16 = nullptr; pa.q_norm2 = nullptr;
      pa.q_exact = rerank ? d_qexact : nullptr; pa.exact_elems = s->exact_elems;
      CUDA_TRY(launch_prep_queries(pa, st));
      launches++;
      sa.done = nullptr; sa.queries = d_q; sa.queries_raw = nullptr;
      sa.k_out = ks;
    }
    if (xchg != nullptr) {
      if (!fused) return fail(RAG_EINVAL, "the fused exchange needs the fused merge (RAG_B200_FUSED_MERGE=0 is set)");
      sa.xchg_peers = xchg->d_peers; sa.xchg_rank = xchg->rank; sa.xchg_world = xchg->world;
      sa.xchg_epoch = xchg_epoch; sa.xchg_slot_keys = xchg->slot_keys;
    }
    if (timed) CUDA_TRY(cudaEventRecord(c->ev0, st));
    CUDA_TRY(launch_scan_stream(sa, s->sm_count, st, &launches));
    if (timed) CUDA_TRY(cudaEventRecord(c->ev1, st));
    if (fused) {   // the scan kernel merged across CTAs, re-ranked and emitted the result itself
      s->launches += launches;
      s->last_launches = launches;
      s->last_regime = regime;
      c->launched = true;
      return RAG_OK;
    }
    merge_src = d_partial;
    S = grid_x;
    k_lists = ks;
    tres.q_exact = d_qexact;
  }
  // ---- generic tail: merge the S lists per query; re-score the winners where ranking was approximate ----
  const bool split = (regime == 2 && s->dtype == RAG_DTYPE_F32);       // contracted through a bf16 shadow (either kind)
  const bool refine = rerank || (regime == 2 && (s->space == RAG_SPACE_L2 || split));
  MergeArgs ma{};
  ma.keys = merge_src; ma.S = S; ma.B = B; ma.k = k_lists; ma.rows_map = rows_map;
  ma.out_keys = out.keys; ma.out_rows = out.rows; ma.out_dists = out.dists; ma.out_counts = out.counts;
  if (refine) {   // merge to scratch keys first (local rows), then re-score the winners exactly
    ma.rows_map = RowMap{}; ma.out_keys = d_merged; ma.out_rows = nullptr; ma.out_dists = nullptr; ma.out_counts = nullptr;
  }
  {
    NvtxRange nvtx_m("rag:merge");
    CUDA_TRY(launch_merge(ma, st));
  }
  launches++;
  if (refine) {
    NvtxRange nvtx_r("rag:rerank");
    RefineArgs ra{};
    ra.keys = d_merged; ra.B = B; ra.k = k; ra.k_in = k_lists;
    if (rerank) {       // the un-rounded fp32 plane and the un-rounded queries
      ra.vectors = s->d_exact; ra.queries = tres.q_exact; ra.dtype = RAG_DTYPE_F32; ra.row_elems = s->exact_elems;
    } else {
      ra.vectors = s->d_vectors; ra.queries = tres.q_f32; ra.dtype = s->dtype; ra.row_elems = s->row_elems;
    }
    ra.l2 = (s->space == RAG_SPACE_L2) ? 1 : 0; ra.rows_map = rows_map;
    ra.out_keys = out.keys; ra.out_rows = out.rows; ra.out_dists = out.dists; ra.out_counts = out.counts;
    if (split) {
      CUDA_TRY(cudaMemsetAsync(d_redo, 0, sizeof(int), st));
      ra.guard_rel = kGuardRel; ra.q_norm2 = tres.q_norm2; ra.x_max_norm2 = s->d_max_norm2;
      ra.q_lo_norm2 = tres.q_lo_norm2; ra.x_lo_max2 = s->d_max_norm2 + 2;      // hi-only shadow: the bf16 rounding it ignores
      ra.redo_count = d_redo; ra.redo_list = d_redo + 1;
    }
    if (split && tres.filt) {      // hi-only filter: every row within 2 eps of the approximate k-th best is re-scored
      RefineFilterArgs fa{};
      fa.r = ra; fa.extra = tres.extra; fa.extra_cnt = tres.extra_cnt; fa.S = tres.S; fa.cap = tres.extra_cap;
      CUDA_TRY(launch_refine_filter(fa, st));
    } else {
      CUDA_TRY(launch_refine(ra, st));
    }
    launches++;
    if (split) {
      // statistics: how many queries the guard sends to the re-run (read when this context searches again)
      if (c->h_redo == nullptr && cudaMallocHost(reinterpret_cast<void**>(&c->h_redo), 64) != cudaSuccess) { (void)cudaGetLastError(); c->h_redo = nullptr; }
      if (c->h_redo != nullptr) {
        CUDA_TRY(cudaMemcpyAsync(c->h_redo, d_redo, sizeof(int), cudaMemcpyDeviceToHost, st));
        c->redo_batch = B; c->redo_kind = s->shadow_kind;
      }
      // queries whose top-k the approximate ranking could not certify are re-run on the exact fp32
      // stream kernel; the launch covers the worst case and exits at once when the list is empty
      rc = c->ensure_tickets();
      if (rc != RAG_OK) return rc;
      if (B > QueryCtx::kMaxTickets) return fail(RAG_EINVAL, "batch %d exceeds %d", B, QueryCtx::kMaxTickets);
      sa.k = k; sa.k_out = k;
      sa.partial = d_partial;
      sa.done = c->d_tickets;
      sa.queries = nullptr; sa.queries_raw = d_queries_raw;
      sa.round_bf16 = 0;
      sa.q_count = d_redo; sa.q_index = d_redo + 1;
      // The launch covers the worst case (every query fails): grid_x x ceil(B / 8) CTAs that exit at once when their
      // group holds no failed query -- 38k CTAs = 46 us at B = 1024 for nothing.  While the guard certifies everything
      // (the last harvested batch had no re-run) a large batch gets an eighth of the CTAs per group: the empty launch
      // costs ~6 us, and should queries fail after all, >= 8 failing groups still fill the machine (fewer run at a
      // fraction of the HBM rate, once: the count that comes back restores the full grid for the next batch).
      if (scan_stream_groups(B, s->dtype, s->row_elems, k) > 8 && s->last_batch_reruns.load(std::memory_order_relaxed) == 0)
        sa.grid_x = std::max(1, grid_x / 8);
      CUDA_TRY(launch_scan_stream(sa, s->sm_count, st, &launches));
    }
  }
  s->launches += launches;
  s->last_launches = launches;
  s->last_regime = regime;
  c->launched = true;
  return RAG_OK;
}

// scratch of the asynchronous API: one context per caller stream
int rag::dev_ctx_for(rag_store* s, void* stream, QueryCtx** out) {
  std::lock_guard<std::mutex> lg(s->dev_mu);
  auto it = s->dev_ctx.find(stream);
  if (it != s->dev_ctx.end()) { *out = it->second; return RAG_OK; }
  QueryCtx* c = new (std::nothrow) QueryCtx();
  if (!c) return fail(RAG_ENOMEM, "out of host memory");
  c->stream = reinterpret_cast<cudaStream_t>(stream);
  c->own_stream = false;
  s->dev_ctx[stream] = c;
  *out = c;
  return RAG_OK;
}

// largest query batch one search_device call may take (bounds the partial buffer)
int rag::batch_limit(const rag_store* s, int k) {
  const int grid_x = scan_stream_grid_x(s->sm_count, s->rows);
  const size_t budget = (size_t)512 << 20;
  int64_t lim = (int64_t)(budget / ((size_t)grid_x * scan_k(s, k) * sizeof(uint64_t)));
  if (lim < 1) lim = 1;
  if (lim > 4096) lim = 4096;
  return (int)lim;
}

int rag::check_query_args(const rag_store* s, int B, const void* q, int k, int mask_slot) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (B <= 0) return fail(RAG_EINVAL, "query batch must be >= 1, got %d", B);
  if (!q) return fail(RAG_EINVAL, "queries is NULL");
  const int kmax = RAG_MAX_K - (s->exact_elems ? 16 : 0);
  if (k < 1 || k > kmax) return fail(RAG_EINVAL, "k must be in [1, %d], got %d", kmax, k);
  if (mask_slot >= RAG_MAX_MASK_SLOTS) return fail(RAG_EINVAL, "mask slot %d out of range", mask_slot);
  if (mask_slot >= 0 && !s->mask_set[mask_slot]) return fail(RAG_EINVAL, "mask slot %d is not set", mask_slot);
  return RAG_OK;
}

// ------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------
extern "C" {

const char* rag_last_error(void) { return g_err; }
int rag_abi_version(void) { return RAG_B200_ABI_VERSION; }

int rag_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return n;
}

int rag_store_create_ex(int dim, int dtype, int space, int device, int64_t capacity_hint, int flags, rag_store** out) {
  if (!out) return fail(RAG_EINVAL, "out is NULL");
  *out = nullptr;
  if (dim <= 0 || dim > 65536) return fail(RAG_EINVAL, "dim must be in [1, 65536], got %d", dim);
  if (dtype != RAG_DTYPE_F32 && dtype != RAG_DTYPE_BF16) return fail(RAG_EINVAL, "unknown dtype %d", dtype);
  if (space < RAG_SPACE_L2 || space > RAG_SPACE_IP) return fail(RAG_EINVAL, "unknown space %d", space);
  if (flags & ~RAG_STORE_NO_RERANK) return fail(RAG_EINVAL, "unknown store flags 0x%x", flags);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return fail(RAG_ENODEV, "no CUDA device available (%s); this engine has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= ndev) return fail(RAG_EINVAL, "device %d out of range (%d visible)", device, ndev);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(RAG_ENODEV, "device %d is sm_%d%d; this build targets sm_100a (B200) only", device, prop.major, prop.minor);
  CUDA_TRY(cudaSetDevice(device));

  rag_store* s = new (std::nothrow) rag_store();
  if (!s) return fail(RAG_ENOMEM, "out of host memory");
  s->dim = dim; s->dtype = dtype; s->space = space; s->device = device;
  const int epc = (dtype == RAG_DTYPE_BF16) ? 8 : 4;    // elements per 16-byte chunk
  s->row_elems = (dim + epc - 1) / epc * epc;
  s->row_bytes = (size_t)s->row_elems * (dtype == RAG_DTYPE_BF16 ? 2 : 4);
  s->exact_elems = (dtype == RAG_DTYPE_BF16 && !(flags & RAG_STORE_NO_RERANK)) ? (dim + 3) / 4 * 4 : 0;
  s->sm_count = prop.multiProcessorCount;
  pthread_rwlock_init(&s->lock, nullptr);
  e = cudaStreamCreateWithFlags(&s->admin.stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete s; return fail(RAG_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  s->admin.own_stream = true;
  e = cudaEventCreateWithFlags(&s->ev_write, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->pending.ev_h2d, cudaEventDisableTiming);
  if (e == cudaSuccess)
    e = cudaMallocHost(reinterpret_cast<void**>(&s->pending.h),
                       (size_t)PendingWrites::kMaxRows * ((size_t)dim * sizeof(float) + sizeof(int64_t)) + 1024);
  if (dtype == RAG_DTYPE_F32) {
    s->shadow_kind = tensor::default_shadow_kind(s->row_elems);
    s->shadow_pinned = getenv("RAG_B200_F32_SHADOW") != nullptr;
  }
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->d_max_norm2), 4 * sizeof(float));
  if (e == cudaSuccess) {
    const float init[4] = {0.0f, __builtin_inff(), 0.0f, 0.0f};      // [0] running max, [1] running min of |stored row|^2, [2] max |x - bf16(x)|^2
    e = cudaMemcpy(s->d_max_norm2, init, sizeof(init), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) { (void)cudaGetLastError(); rag_store_destroy(s); return fail(RAG_ENOMEM, "store set-up failed: %s", cudaGetErrorString(e)); }
  int rc = grow(s, std::max<int64_t>(capacity_hint, 1024));
  if (rc != RAG_OK) { rag_store_destroy(s); return rc; }
  *out = s;
  return RAG_OK;
}

int rag_store_create(int dim, int dtype, int space, int device, int64_t capacity_hint, rag_store** out) {
  return rag_store_create_ex(dim, dtype, space, device, capacity_hint, 0, out);
}

int rag_store_destroy(rag_store* s) {
  if (!s) return RAG_OK;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();
  if (s->pipe) {
    for (PipeSlot& ps : s->pipe->slot) {
      if (ps.h) cudaFreeHost(ps.h);
      if (ps.d) cudaFree(ps.d);
      if (ps.ev) cudaEventDestroy(ps.ev);
    }
  }
  for (QueryCtx* c : s->pool_free) { c->destroy(); delete c; }
  for (auto& kv : s->dev_ctx) { kv.second->destroy(); delete kv.second; }
  if (s->pipe) {
    if (s->pipe->stream) cudaStreamDestroy(s->pipe->stream);
    if (s->pipe->copy_stream) cudaStreamDestroy(s->pipe->copy_stream);
    delete s->pipe;
  }
  s->admin.destroy();
  if (s->ev_write) cudaEventDestroy(s->ev_write);
  if (s->pending.ev_h2d) cudaEventDestroy(s->pending.ev_h2d);
  if (s->pending.h) cudaFreeHost(s->pending.h);
  for (int i = 0; i < RAG_MAX_MASK_SLOTS; ++i) if (s->d_masks[i]) cudaFree(s->d_masks[i]);
  if (s->d_vectors) cudaFree(s->d_vectors);
  if (s->d_exact) cudaFree(s->d_exact);
  if (s->d_norms2) cudaFree(s->d_norms2);
  if (s->d_max_norm2) cudaFree(s->d_max_norm2);
  if (s->d_shadow) cudaFree(s->d_shadow);
  if (s->d_live) cudaFree(s->d_live);
  pthread_rwlock_destroy(&s->lock);
  delete s;
  return RAG_OK;
}

int rag_store_reserve(rag_store* s, int64_t rows) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  WrLock g(&s->lock);
  CUDA_TRY(cudaSetDevice(s->device));
  return grow(s, rows);
}

int rag_store_flush(rag_store* s) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  WrLock g(&s->lock);
  CUDA_TRY(cudaSetDevice(s->device));
  int rc = flush_pending_locked(s);
  if (rc != RAG_OK) return rc;
  CUDA_TRY(cudaStreamSynchronize(s->admin.stream));
  return RAG_OK;
}

static int upsert_impl(rag_store* s, int64_t n, const float* vectors, bool on_device, const int64_t* rows, int64_t* out_rows) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (n < 0) return fail(RAG_EINVAL, "n < 0");
  if (n == 0) return RAG_OK;
  if (!vectors) return fail(RAG_EINVAL, "vectors is NULL");
  NvtxRange nvtx("rag:upsert");
  WrLock g(&s->lock);
  CUDA_TRY(cudaSetDevice(s->device));
  std::vector<int64_t> dst, claimed;
  int rc = assign_rows(s, n, rows, dst, claimed);
  if (rc != RAG_OK) return rc;
  int64_t top = s->rows;
  for (int64_t r : dst) top = std::max(top, r + 1);
  rc = grow(s, top);
  if (rc != RAG_OK) { unclaim(s, claimed); return rc; }
  const size_t row_in = (size_t)s->dim * sizeof(float);
  PendingWrites& p = s->pending;

  if (!on_device && n <= PendingWrites::kSmallCall) {
    // ---- small write: park it on the host; the next read (or a full block) sends it down ----
    if (p.n + n > PendingWrites::kMaxRows) {
      rc = flush_pending_locked(s);
      if (rc != RAG_OK) { unclaim(s, claimed); return rc; }
    }
    if (p.in_flight) {            // the previous flush must have left the pinned block before it is rewritten
      CUDA_TRY(cudaEventSynchronize(p.ev_h2d));
      p.in_flight = false;
    }
    int64_t* h_rows = reinterpret_cast<int64_t*>(p.h + (size_t)PendingWrites::kMaxRows * row_in);
    for (int64_t i = 0; i < n; ++i) {
      const int64_t r = dst[(size_t)i];
      int64_t slot;
      auto it = p.slot_of.find(r);
      if (it != p.slot_of.end()) slot = it->second;          // a second write to the same row replaces the first
      else { slot = p.n++; p.slot_of.emplace(r, slot); h_rows[slot] = r; }
      memcpy(p.h + (size_t)slot * row_in, vectors + (size_t)i * s->dim, row_in);
    }
    s->pending_n.store(p.n, std::memory_order_release);
    mark_live(s, dst);
    if (out_rows) memcpy(out_rows, dst.data(), (size_t)n * sizeof(int64_t));
    return RAG_OK;
  }

  // ---- bulk write: straight to the device (behind anything parked, which may touch the same rows) ----
  rc = flush_pending_locked(s);
  if (rc == RAG_OK) rc = writer_wait_for_readers(s);
  if (rc != RAG_OK) { unclaim(s, claimed); return rc; }
  QueryCtx& c = s->admin;
  const bool contig = contiguous(dst.data(), n);
  auto fail_cleanup = [&](int code) {
    // host bookkeeping never saw the batch; drop whatever live bits the device already set for it
    char keep[sizeof(g_err)];
    memcpy(keep, g_err, sizeof(keep));
    unclaim(s, claimed);
    std::vector<int64_t> dead;
    for (int64_t r : dst) if (!h_is_live(s, r)) dead.push_back(r);
    (void)cudaStreamSynchronize(c.stream);
    (void)cudaGetLastError();
    if (!dead.empty() && c.ensure_dev(dead.size() * sizeof(int64_t)) == RAG_OK &&
        cudaMemcpyAsync(c.d_buf, dead.data(), dead.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream) == cudaSuccess) {
      (void)launch_clear_live(s->d_live, reinterpret_cast<const int64_t*>(c.d_buf), (int64_t)dead.size(), c.stream);
      (void)cudaStreamSynchronize(c.stream);
    }
    (void)cudaGetLastError();
    (void)writer_publish(s);
    memcpy(g_err, keep, sizeof(keep));
    return code;
  };
  if (on_device) {
    int64_t* d_rows = nullptr;
    if (!contig) {
      rc = c.ensure_dev((size_t)n * sizeof(int64_t));
      if (rc != RAG_OK) return fail_cleanup(rc);
      d_rows = reinterpret_cast<int64_t*>(c.d_buf);
      if (cudaMemcpyAsync(d_rows, dst.data(), (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream) != cudaSuccess)
        return fail_cleanup(fail(RAG_ECUDA, "copy of the destination rows failed"));
    }
    rc = c.ensure_events();
    if (rc == RAG_OK) (void)cudaEventRecord(c.ev0, c.stream);
    if (rc == RAG_OK) rc = launch_upsert_rows(s, vectors, n, d_rows, dst[0]);
    if (rc == RAG_OK) (void)cudaEventRecord(c.ev1, c.stream);
    // the caller's device buffer is only guaranteed to live until this call returns
    if (rc == RAG_OK) {
      cudaError_t e = cudaStreamSynchronize(c.stream);
      if (e != cudaSuccess) rc = fail(RAG_ECUDA, "upsert kernel failed: %s", cudaGetErrorString(e));
      else if (cudaEventElapsedTime(&s->last_upsert_ms, c.ev0, c.ev1) != cudaSuccess) { (void)cudaGetLastError(); s->last_upsert_ms = 0.0f; }
    }
    if (rc != RAG_OK) return fail_cleanup(rc);
  } else {
    // stage through two pinned halves of <= 32 MB: filling half i+1 overlaps the copy + kernel of half i
    int64_t per = std::max<int64_t>(1, (int64_t)(((size_t)32 << 20) / (row_in + sizeof(int64_t))));
    per = std::min<int64_t>(per, n);
    const size_t vec_b = align_up((size_t)per * row_in, 256);
    const size_t half = vec_b + align_up((size_t)per * sizeof(int64_t), 256);
    rc = c.ensure_host(2 * half);
    if (rc == RAG_OK) rc = c.ensure_dev(2 * half);
    if (rc == RAG_OK) rc = c.ensure_events();
    if (rc != RAG_OK) return fail_cleanup(rc);
    if (cudaStreamSynchronize(c.stream) != cudaSuccess)      // an earlier write may still be reading the pinned block
      return fail_cleanup(fail(RAG_ECUDA, "an earlier write failed: %s", cudaGetErrorString(cudaGetLastError())));
    cudaEvent_t ev[2] = {c.ev0, c.ev1};
    bool used[2] = {false, false};
    int hsel = 0;
    for (int64_t i0 = 0; i0 < n; i0 += per, hsel ^= 1) {
      const int64_t m = std::min<int64_t>(per, n - i0);
      if (used[hsel] && cudaEventSynchronize(ev[hsel]) != cudaSuccess) return fail_cleanup(fail(RAG_ECUDA, "upsert staging failed"));
      unsigned char* hp = c.h_pin + (size_t)hsel * half;
      unsigned char* dp = c.d_buf + (size_t)hsel * half;
      memcpy(hp, vectors + (size_t)i0 * s->dim, (size_t)m * row_in);
      size_t bytes = (size_t)m * row_in;
      if (!contig) { memcpy(hp + vec_b, dst.data() + i0, (size_t)m * sizeof(int64_t)); bytes = vec_b + (size_t)m * sizeof(int64_t); }
      if (cudaMemcpyAsync(dp, hp, bytes, cudaMemcpyHostToDevice, c.stream) != cudaSuccess)
        return fail_cleanup(fail(RAG_ECUDA, "host-to-device copy of the vectors failed"));
      rc = launch_upsert_rows(s, reinterpret_cast<const float*>(dp), m,
                              contig ? nullptr : reinterpret_cast<const int64_t*>(dp + vec_b), dst[(size_t)i0]);
      if (rc != RAG_OK) return fail_cleanup(rc);
      (void)cudaEventRecord(ev[hsel], c.stream);
      used[hsel] = true;
    }
  }
  mark_live(s, dst);
  if (out_rows) memcpy(out_rows, dst.data(), (size_t)n * sizeof(int64_t));
  return writer_publish(s);
}

int rag_store_upsert(rag_store* s, int64_t n, const float* vectors, const int64_t* rows, int64_t* out_rows) {
  return upsert_impl(s, n, vectors, false, rows, out_rows);
}
int rag_store_upsert_dev(rag_store* s, int64_t n, const float* vectors_dev, const int64_t* rows, int64_t* out_rows) {
  return upsert_impl(s, n, vectors_dev, true, rows, out_rows);
}

int rag_store_delete(rag_store* s, int64_t n, const int64_t* rows) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (n <= 0) return RAG_OK;
  if (!rows) return fail(RAG_EINVAL, "rows is NULL");
  NvtxRange nvtx("rag:delete");
  WrLock g(&s->lock);
  CUDA_TRY(cudaSetDevice(s->device));
  int rc = flush_pending_locked(s);      // a parked write to a victim must set its live bit BEFORE it is cleared
  if (rc != RAG_OK) return rc;
  std::vector<int64_t> victims;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t r = rows[i];
    if (!h_is_live(s, r)) continue;
    s->h_live[(size_t)(r >> 5)] &= ~(1u << (r & 31));
    s->live--;
    if (!s->external_rows) s->free_rows.push_back(r);
    victims.push_back(r);
  }
  if (victims.empty()) return RAG_OK;
  QueryCtx& c = s->admin;
  const size_t bytes = victims.size() * sizeof(int64_t);
  rc = c.ensure_dev(bytes);
  if (rc == RAG_OK) rc = c.ensure_host(bytes);
  if (rc == RAG_OK) rc = writer_wait_for_readers(s);
  if (rc != RAG_OK) return rc;
  CUDA_TRY(cudaStreamSynchronize(c.stream));               // the pinned block may still feed an earlier copy
  memcpy(c.h_pin, victims.data(), bytes);
  CUDA_TRY(cudaMemcpyAsync(c.d_buf, c.h_pin, bytes, cudaMemcpyHostToDevice, c.stream));
  CUDA_TRY(launch_clear_live(s->d_live, reinterpret_cast<const int64_t*>(c.d_buf), (int64_t)victims.size(), c.stream));
  s->launches++;
  return writer_publish(s);
}

int64_t rag_store_count(const rag_store* s) { return s ? s->live : 0; }
int64_t rag_store_rows(const rag_store* s) { return s ? s->rows : 0; }
int64_t rag_store_capacity(const rag_store* s) { return s ? s->capacity : 0; }
int rag_store_dim(const rag_store* s) { return s ? s->dim : 0; }
int rag_store_dtype(const rag_store* s) { return s ? s->dtype : 0; }
int rag_store_space(const rag_store* s) { return s ? s->space : 0; }
int rag_store_device(const rag_store* s) { return s ? s->device : -1; }
int rag_store_has_rerank(const rag_store* s) { return (s && s->exact_elems) ? 1 : 0; }
int64_t rag_store_kernel_launches(const rag_store* s) { return s ? s->launches.load() : 0; }

int rag_store_is_live(const rag_store* s, int64_t row) {
  if (!s) return 0;
  RdLock g(const_cast<pthread_rwlock_t*>(&s->lock));
  return h_is_live(s, row) ? 1 : 0;
}

static int fetch_impl(rag_store* s, int64_t n, const int64_t* rows, float* out, bool exact) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (n <= 0) return RAG_OK;
  if (!rows || !out) return fail(RAG_EINVAL, "rows/out is NULL");
  WrLock g(&s->lock);   // uses the admin scratch
  CUDA_TRY(cudaSetDevice(s->device));
  for (int64_t i = 0; i < n; ++i)
    if (rows[i] < 0 || rows[i] >= s->rows) return fail(RAG_EINVAL, "fetch row %lld out of range", (long long)rows[i]);
  int rc = flush_pending_locked(s);
  if (rc != RAG_OK) return rc;
  QueryCtx& c = s->admin;
  const size_t rb = align_up((size_t)n * sizeof(int64_t), 256);
  const size_t ob = (size_t)n * s->dim * sizeof(float);
  rc = c.ensure_dev(rb + ob);
  if (rc != RAG_OK) return rc;
  int64_t* d_rows = reinterpret_cast<int64_t*>(c.d_buf);
  float* d_out = reinterpret_cast<float*>(c.d_buf + rb);
  CUDA_TRY(cudaMemcpyAsync(d_rows, rows, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
  if (exact && s->d_exact) CUDA_TRY(launch_fetch(s->d_exact, RAG_DTYPE_F32, s->dim, s->exact_elems, d_rows, n, d_out, c.stream));
  else CUDA_TRY(launch_fetch(s->d_vectors, s->dtype, s->dim, s->row_elems, d_rows, n, d_out, c.stream));
  CUDA_TRY(cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  s->launches++;
  return RAG_OK;
}

int rag_store_fetch(rag_store* s, int64_t n, const int64_t* rows, float* out) { return fetch_impl(s, n, rows, out, false); }
int rag_store_fetch_exact(rag_store* s, int64_t n, const int64_t* rows, float* out) { return fetch_impl(s, n, rows, out, true); }

int rag_store_set_mask(rag_store* s, int slot, const uint64_t* bits, int64_t nbits) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (slot < 0 || slot >= RAG_MAX_MASK_SLOTS) return fail(RAG_EINVAL, "mask slot %d out of range", slot);
  if (nbits < 0 || (nbits > 0 && !bits)) return fail(RAG_EINVAL, "bad mask arguments");
  NvtxRange nvtx("rag:mask");
  WrLock g(&s->lock);
  CUDA_TRY(cudaSetDevice(s->device));
  const int64_t words64 = (nbits + 63) / 64;
  const int64_t cap_words = s->capacity / 32;               // a mask is allocated once per capacity, not per call
  const int64_t words32 = std::min<int64_t>(words64 * 2, cap_words);    // bits beyond the capacity can name no row
  QueryCtx& c = s->admin;
  int rc = writer_wait_for_readers(s);
  if (rc != RAG_OK) return rc;
  if (!s->d_masks[slot]) {
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_masks[slot]), (size_t)cap_words * sizeof(uint32_t)));
    CUDA_TRY(cudaMemsetAsync(s->d_masks[slot], 0, (size_t)cap_words * sizeof(uint32_t), c.stream));
    s->mask_words[slot] = 0;
  }
  if (words32 > 0) {
    rc = c.ensure_host((size_t)words32 * 4);
    if (rc != RAG_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(c.stream));              // the pinned block may still feed an earlier copy
    memcpy(c.h_pin, bits, (size_t)words32 * 4);
    if (nbits < words32 * 32) {                             // bits past nbits do not pass
      uint32_t* w = reinterpret_cast<uint32_t*>(c.h_pin);
      if (nbits % 32) w[nbits / 32] &= (~0u >> (32 - nbits % 32));
      for (int64_t i = (nbits + 31) / 32; i < words32; ++i) w[i] = 0u;
    }
    CUDA_TRY(cudaMemcpyAsync(s->d_masks[slot], c.h_pin, (size_t)words32 * 4, cudaMemcpyHostToDevice, c.stream));
  }
  if (s->mask_words[slot] > words32)                        // a shorter mask replaces a longer one: clear the old tail
    CUDA_TRY(cudaMemsetAsync(s->d_masks[slot] + words32, 0, (size_t)(s->mask_words[slot] - words32) * 4, c.stream));
  s->mask_words[slot] = words32;
  s->mask_set[slot] = true;
  return writer_publish(s);
}

int rag_store_patch_mask(rag_store* s, int slot, int64_t n, const int64_t* rows, const unsigned char* pass) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (slot < 0 || slot >= RAG_MAX_MASK_SLOTS) return fail(RAG_EINVAL, "mask slot %d out of range", slot);
  if (n <= 0) return RAG_OK;
  if (!rows || !pass) return fail(RAG_EINVAL, "rows/pass is NULL");
  NvtxRange nvtx("rag:mask");
  WrLock g(&s->lock);
  if (!s->mask_set[slot]) return fail(RAG_EINVAL, "mask slot %d is not set", slot);
  CUDA_TRY(cudaSetDevice(s->device));
  int64_t top = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (rows[i] < 0 || rows[i] >= s->capacity) return fail(RAG_EINVAL, "mask row %lld out of range", (long long)rows[i]);
    top = std::max(top, rows[i] + 1);
  }
  QueryCtx& c = s->admin;
  const size_t rb = align_up((size_t)n * sizeof(int64_t), 256);
  int rc = c.ensure_dev(rb + (size_t)n);
  if (rc == RAG_OK) rc = c.ensure_host(rb + (size_t)n);
  if (rc == RAG_OK) rc = writer_wait_for_readers(s);
  if (rc != RAG_OK) return rc;
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  memcpy(c.h_pin, rows, (size_t)n * sizeof(int64_t));
  memcpy(c.h_pin + rb, pass, (size_t)n);
  CUDA_TRY(cudaMemcpyAsync(c.d_buf, c.h_pin, rb + (size_t)n, cudaMemcpyHostToDevice, c.stream));
  CUDA_TRY(launch_patch_mask(s->d_masks[slot], reinterpret_cast<const int64_t*>(c.d_buf), c.d_buf + rb, n, c.stream));
  s->launches++;
  s->mask_words[slot] = std::max<int64_t>(s->mask_words[slot], (top + 31) / 32);   // the words in between are zero
  return writer_publish(s);
}

int rag_store_clear_mask(rag_store* s, int slot) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (slot < 0 || slot >= RAG_MAX_MASK_SLOTS) return fail(RAG_EINVAL, "mask slot %d out of range", slot);
  WrLock g(&s->lock);
  CUDA_TRY(cudaSetDevice(s->device));
  if (s->d_masks[slot] && s->mask_words[slot] > 0) {
    int rc = writer_wait_for_readers(s);
    if (rc != RAG_OK) return rc;
    CUDA_TRY(cudaMemsetAsync(s->d_masks[slot], 0, (size_t)s->mask_words[slot] * 4, s->admin.stream));
    rc = writer_publish(s);
    if (rc != RAG_OK) return rc;
  }
  s->mask_words[slot] = 0;
  s->mask_set[slot] = false;
  return RAG_OK;
}

int rag_store_query(rag_store* s, int B, const float* queries, int k, int mask_slot, int flags,
                    int64_t* out_rows, float* out_dists, int32_t* out_counts) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  int rc = flush_if_pending(s);
  if (rc != RAG_OK) return rc;
  RdLock g(&s->lock);
  rc = check_query_args(s, B, queries, k, mask_slot);
  if (rc != RAG_OK) return rc;
  if (!out_rows || !out_dists || !out_counts) return fail(RAG_EINVAL, "output pointer is NULL");
  if (s->live == 0) { fill_empty(B, k, out_rows, out_dists, out_counts); return RAG_OK; }
  CUDA_TRY(cudaSetDevice(s->device));
  QueryCtx* c = nullptr;
  rc = acquire_ctx(s, &c);
  if (rc != RAG_OK) return rc;
  CtxGuard cg{s, c};

  const int lim = batch_limit(s, k);
  float total_ms = 0.0f;
  int total_launches = 0;
  int regime_used = 0;
  for (int b0 = 0; b0 < B; b0 += lim) {
    const int Bc = std::min(lim, B - b0);
    const int regime = choose_regime(s, Bc, k, flags);
    if (regime < 0) return fail(RAG_EINVAL, "tensor regime does not support this store/query (dtype %d, dim %d, k %d)", s->dtype, s->dim, k);
    const int grid_x = scan_stream_grid_x(s->sm_count, s->rows);
    const size_t in_b = align_up((size_t)Bc * s->dim * sizeof(float), 256);
    const size_t rows_b = align_up((size_t)Bc * k * sizeof(int64_t), 256);
    const size_t dist_b = align_up((size_t)Bc * k * sizeof(float), 256);
    const size_t cnt_b = align_up((size_t)Bc * sizeof(int32_t), 256);
    const size_t scr_b = search_scratch_bytes(s, Bc, k, grid_x);
    rc = c->ensure_host(in_b + rows_b + dist_b + cnt_b + 256);      // + the completion flag
    if (rc != RAG_OK) return rc;
    rc = c->ensure_dev(in_b + rows_b + dist_b + cnt_b + scr_b);
    if (rc != RAG_OK) return rc;
    unsigned char* d = c->d_buf;
    float* d_in = reinterpret_cast<float*>(d);
    SearchOut so{};
    so.rows = reinterpret_cast<int64_t*>(d + in_b);
    so.dists = reinterpret_cast<float*>(d + in_b + rows_b);
    so.counts = reinterpret_cast<int32_t*>(d + in_b + rows_b + dist_b);
    unsigned char* scratch = d + in_b + rows_b + dist_b + cnt_b;

    memcpy(c->h_pin, queries + (size_t)b0 * s->dim, (size_t)Bc * s->dim * sizeof(float));
    CUDA_TRY(cudaMemcpyAsync(d_in, c->h_pin, (size_t)Bc * s->dim * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    bool armed = false;
    if (direct_host_ok(s, Bc, k, regime)) {
      // small stream-regime batch: the kernel writes the result straight into the pinned block (mapped host
      // memory) and raises a flag there -- no device-to-host copy, no stream synchronisation
      so.rows = reinterpret_cast<int64_t*>(c->h_pin + in_b);
      so.dists = reinterpret_cast<float*>(c->h_pin + in_b + rows_b);
      so.counts = reinterpret_cast<int32_t*>(c->h_pin + in_b + rows_b + dist_b);
      so.done_flag = reinterpret_cast<uint32_t*>(c->h_pin + in_b + rows_b + dist_b + cnt_b);
      so.done_seq = ++c->signal_seq ? c->signal_seq : ++c->signal_seq;
      so.armed = &armed;
    }
    rc = search_device(s, c, scratch, Bc, d_in, k, mask_slot, regime, RowMap{}, so, true, nullptr, 0u, flags == RAG_QUERY_FORCE_TENSOR);
    if (rc != RAG_OK) return rc;
    regime_used = s->last_regime.load();     // the regime that actually ran (an fp32 store may have fallen back)
    if (armed) {
      rc = wait_host_flag(so.done_flag, so.done_seq, c->stream);
      if (rc != RAG_OK) return rc;
    } else {
      // one D2H for rows + dists + counts (contiguous in the scratch and in the pinned buffer)
      CUDA_TRY(cudaMemcpyAsync(c->h_pin + in_b, d + in_b, rows_b + dist_b + cnt_b, cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    memcpy(out_rows + (size_t)b0 * k, c->h_pin + in_b, (size_t)Bc * k * sizeof(int64_t));
    memcpy(out_dists + (size_t)b0 * k, c->h_pin + in_b + rows_b, (size_t)Bc * k * sizeof(float));
    memcpy(out_counts + b0, c->h_pin + in_b + rows_b + dist_b, (size_t)Bc * sizeof(int32_t));
    harvest_redo(s, c);            // fp32 tensor regime: the guard's re-run count came back with the result
    float ms = 0.0f;
    if (armed) {
      // the flag is raised a moment before the kernel (and the event behind it) completes: leave the timing to
      // rag_store_last_query_info(), which waits for the event only if somebody asks
      s->timing_ctx.store(c, std::memory_order_release);
    } else if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) {
      total_ms += ms;
    } else {
      (void)cudaGetLastError();
    }
    total_launches += s->last_launches.load();
  }
  s->last_kernel_ms = total_ms;
  s->last_launches = total_launches;
  s->last_regime = regime_used;
  return RAG_OK;
}

int rag_store_query_dev(rag_store* s, int B, const float* queries_dev, int k, int mask_slot, int flags,
                        uint32_t row_base, uint64_t* out_keys_dev, int64_t* out_rows_dev, float* out_dists_dev,
                        int32_t* out_counts_dev, void* stream) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  int rc = flush_if_pending(s);
  if (rc != RAG_OK) return rc;
  RdLock g(&s->lock);
  rc = check_query_args(s, B, queries_dev, k, mask_slot);
  if (rc != RAG_OK) return rc;
  if (!out_keys_dev && !out_rows_dev) return fail(RAG_EINVAL, "out_keys_dev and out_rows_dev are both NULL");
  CUDA_TRY(cudaSetDevice(s->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (s->live == 0) {
    if (out_keys_dev) CUDA_TRY(cudaMemsetAsync(out_keys_dev, 0xFF, (size_t)B * k * sizeof(uint64_t), st));
    if (out_rows_dev) CUDA_TRY(cudaMemsetAsync(out_rows_dev, 0xFF, (size_t)B * k * sizeof(int64_t), st));   // -1
    if (out_dists_dev) CUDA_TRY(cudaMemsetAsync(out_dists_dev, 0x7F, (size_t)B * k * sizeof(float), st));     // large, not inf
    if (out_counts_dev) CUDA_TRY(cudaMemsetAsync(out_counts_dev, 0, (size_t)B * sizeof(int32_t), st));
    return RAG_OK;
  }
  if (B > batch_limit(s, k)) return fail(RAG_EINVAL, "batch %d too large for one asynchronous call (limit %d)", B, batch_limit(s, k));
  QueryCtx* c = nullptr;
  rc = dev_ctx_for(s, stream, &c);
  if (rc != RAG_OK) return rc;
  const int regime = choose_regime(s, B, k, flags);
  if (regime < 0) return fail(RAG_EINVAL, "tensor regime does not support this store/query");
  const int grid_x = scan_stream_grid_x(s->sm_count, s->rows);
  rc = c->ensure_dev(search_scratch_bytes(s, B, k, grid_x));   // stream-ordered reuse: same stream, same scratch
  if (rc != RAG_OK) return rc;
  SearchOut so{};
  so.keys = out_keys_dev;
  so.rows = out_rows_dev;
  so.dists = out_dists_dev;
  so.counts = out_counts_dev;
  return search_device(s, c, c->d_buf, B, queries_dev, k, mask_slot, regime, RowMap{row_base, 0u, 1u}, so, false, nullptr, 0u,
                       flags == RAG_QUERY_FORCE_TENSOR);
}

// ---- fused cross-shard exchange (multi-GPU, stream regime) ---------------------------------
int rag_exchange_create(int device, int rank, int world, int64_t slot_keys, rag_exchange** out) {
  if (!out) return fail(RAG_EINVAL, "out is NULL");
  *out = nullptr;
  if (world < 1 || world > kXchgMaxWorld || rank < 0 || rank >= world)
    return fail(RAG_EINVAL, "exchange needs 0 <= rank < world <= %d, got rank %d world %d", kXchgMaxWorld, rank, world);
  if (slot_keys < 1 || slot_keys > (1 << 22)) return fail(RAG_EINVAL, "slot_keys must be in [1, 2^22]");
  CUDA_TRY(cudaSetDevice(device));
  rag_exchange* x = new (std::nothrow) rag_exchange();
  if (!x) return fail(RAG_ENOMEM, "out of host memory");
  x->device = device; x->rank = rank; x->world = world; x->slot_keys = slot_keys;
  x->bytes = xchg_buffer_bytes(world, slot_keys);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&x->d_local), x->bytes);
  if (e == cudaSuccess) e = cudaMemset(x->d_local, 0, x->bytes);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->d_peers), kXchgMaxWorld * sizeof(unsigned char*));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&x->host.stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    if (x->d_local) cudaFree(x->d_local);
    if (x->d_peers) cudaFree(x->d_peers);
    delete x;
    return fail(RAG_ENOMEM, "exchange buffer allocation failed: %s", cudaGetErrorString(e));
  }
  x->host.own_stream = true;
  x->peers.assign((size_t)world, nullptr);
  *out = x;
  return RAG_OK;
}

int rag_exchange_handle(rag_exchange* x, void* out_handle) {
  if (!x || !out_handle) return fail(RAG_EINVAL, "exchange/out is NULL");
  static_assert(sizeof(cudaIpcMemHandle_t) == RAG_EXCHANGE_HANDLE_BYTES, "IPC handle size");
  CUDA_TRY(cudaSetDevice(x->device));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, x->d_local));
  memcpy(out_handle, &h, sizeof(h));
  return RAG_OK;
}

int rag_exchange_connect(rag_exchange* x, const void* handles) {
  if (!x || !handles) return fail(RAG_EINVAL, "exchange/handles is NULL");
  if (x->connected) return fail(RAG_EINVAL, "exchange is already connected");
  CUDA_TRY(cudaSetDevice(x->device));
  for (int g = 0; g < x->world; ++g) {
    if (g == x->rank) { x->peers[(size_t)g] = x->d_local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)g * sizeof(h), sizeof(h));
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->peers[(size_t)g] = static_cast<unsigned char*>(p);
  }
  CUDA_TRY(cudaMemcpy(x->d_peers, x->peers.data(), (size_t)x->world * sizeof(unsigned char*), cudaMemcpyHostToDevice));
  x->connected = true;
  x->ipc = true;
  return RAG_OK;
}

int rag_exchange_status(rag_exchange* x, int* timed_out) {
  if (!x || !timed_out) return fail(RAG_EINVAL, "exchange/out is NULL");
  CUDA_TRY(cudaSetDevice(x->device));
  uint32_t st = 0;
  CUDA_TRY(cudaMemcpy(&st, x->d_local + kXchgStatusOff, sizeof(st), cudaMemcpyDeviceToHost));
  *timed_out = st != 0 ? 1 : 0;
  if (st != 0) CUDA_TRY(cudaMemset(x->d_local + kXchgStatusOff, 0, sizeof(st)));     // read-and-clear
  return RAG_OK;
}

int rag_exchange_destroy(rag_exchange* x) {
  if (!x) return RAG_OK;
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  if (x->ipc)
    for (int g = 0; g < x->world; ++g)
      if (g != x->rank && x->peers[(size_t)g]) cudaIpcCloseMemHandle(x->peers[(size_t)g]);
  x->host.destroy();
  if (x->d_peers) cudaFree(x->d_peers);
  if (x->d_local) cudaFree(x->d_local);
  delete x;
  return RAG_OK;
}

int rag_store_fused_ok(const rag_store* s, const rag_exchange* x, int B, int k, int flags) {
  if (!s || !x || !x->connected || B < 1 || k < 1 || scan_k(s, k) > 128) return 0;
  if (choose_regime(s, B, k, flags) != 1) return 0;
  if ((int64_t)B * k > x->slot_keys) return 0;
  if (scan_stream_groups(B, s->dtype, s->row_elems, scan_k(s, k)) > kXchgMaxGroups) return 0;
  if (B > QueryCtx::kMaxTickets || B > batch_limit(s, k)) return 0;
  return 1;
}

// argument checks shared by the two fused entry points; deterministic across ranks (every rank makes the
// same call on a store of the same shape), so either all ranks pass or none does
static int fused_check(rag_store* s, rag_exchange* x, int B, const void* q, int k, int mask_slot, int flags) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (!x || !x->connected) return fail(RAG_EINVAL, "exchange is NULL or not connected");
  int rc = check_query_args(s, B, q, k, mask_slot);
  if (rc != RAG_OK) return rc;
  if (x->device != s->device) return fail(RAG_EINVAL, "exchange and store live on different devices");
  if (!rag_store_fused_ok(s, x, B, k, flags))
    return fail(RAG_EINVAL, "batch %d / k %d is not served by the fused exchange (use rag_store_query_dev + all-gather)", B, k);
  return RAG_OK;
}

int rag_store_query_fused_dev(rag_store* s, rag_exchange* x, int B, const float* queries_dev, int k, int mask_slot,
                              int flags, uint32_t row_base, int64_t* out_rows_dev, float* out_dists_dev,
                              int32_t* out_counts_dev, void* stream) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  int rc = flush_if_pending(s);
  if (rc != RAG_OK) return rc;
  RdLock g(&s->lock);
  rc = fused_check(s, x, B, queries_dev, k, mask_slot, flags);
  if (rc != RAG_OK) return rc;
  if (!out_rows_dev) return fail(RAG_EINVAL, "out_rows_dev is NULL");
  // The epoch is taken BEFORE anything that can fail on one rank only (allocations): the ranks' epochs
  // stay in step, and a rank that fails below simply never delivers this epoch -- its peers time out on
  // it instead of mis-reading the next one.
  const uint32_t epoch = ++x->epoch;
  CUDA_TRY(cudaSetDevice(s->device));
  QueryCtx* c = nullptr;
  rc = dev_ctx_for(s, stream, &c);
  if (rc != RAG_OK) return rc;
  const int grid_x = scan_stream_grid_x(s->sm_count, s->rows);
  rc = c->ensure_dev(search_scratch_bytes(s, B, k, grid_x));
  if (rc != RAG_OK) return rc;
  SearchOut so{};
  so.rows = out_rows_dev; so.dists = out_dists_dev; so.counts = out_counts_dev;
  // an empty shard still takes part: it publishes empty lists and waits like everyone else
  return search_device(s, c, c->d_buf, B, queries_dev, k, mask_slot, 1, RowMap{row_base, 0u, 1u}, so, false, x, epoch, false);
}

int rag_store_query_fused(rag_store* s, rag_exchange* x, int B, const float* queries, int k, int mask_slot, int flags,
                          uint32_t row_base, int64_t* out_rows, float* out_dists, int32_t* out_counts) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  int rc = flush_if_pending(s);
  if (rc != RAG_OK) return rc;
  RdLock g(&s->lock);
  rc = fused_check(s, x, B, queries, k, mask_slot, flags);
  if (rc != RAG_OK) return rc;
  if (!out_rows || !out_dists || !out_counts) return fail(RAG_EINVAL, "output pointer is NULL");
  const uint32_t epoch = ++x->epoch;      // see rag_store_query_fused_dev
  CUDA_TRY(cudaSetDevice(s->device));
  QueryCtx* c = &x->host;
  // pinned block and its device mirror: queries | rows | dists | counts, then the search scratch
  const size_t in_b = align_up((size_t)B * s->dim * sizeof(float), 256);
  const size_t rows_b = align_up((size_t)B * k * sizeof(int64_t), 256);
  const size_t dist_b = align_up((size_t)B * k * sizeof(float), 256);
  const size_t cnt_b = align_up((size_t)B * sizeof(int32_t), 256);
  const size_t io_b = in_b + rows_b + dist_b + cnt_b;
  rc = c->ensure_host(io_b + 256);         // + the completion flag
  if (rc != RAG_OK) return rc;
  const int grid_x = scan_stream_grid_x(s->sm_count, s->rows);
  rc = c->ensure_dev(io_b + search_scratch_bytes(s, B, k, grid_x));
  if (rc != RAG_OK) return rc;
  unsigned char* d = c->d_buf;
  memcpy(c->h_pin, queries, (size_t)B * s->dim * sizeof(float));
  CUDA_TRY(cudaMemcpyAsync(d, c->h_pin, (size_t)B * s->dim * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  SearchOut so{};
  so.rows = reinterpret_cast<int64_t*>(d + in_b);
  so.dists = reinterpret_cast<float*>(d + in_b + rows_b);
  so.counts = reinterpret_cast<int32_t*>(d + in_b + rows_b + dist_b);
  bool armed = false;
  if (direct_host_ok(s, B, k, 1)) {       // result + completion flag straight into the pinned block (see rag_store_query)
    so.rows = reinterpret_cast<int64_t*>(c->h_pin + in_b);
    so.dists = reinterpret_cast<float*>(c->h_pin + in_b + rows_b);
    so.counts = reinterpret_cast<int32_t*>(c->h_pin + in_b + rows_b + dist_b);
    so.done_flag = reinterpret_cast<uint32_t*>(c->h_pin + io_b);
    so.done_seq = ++c->signal_seq ? c->signal_seq : ++c->signal_seq;
    so.armed = &armed;
  }
  rc = search_device(s, c, d + io_b, B, reinterpret_cast<const float*>(d), k, mask_slot, 1, RowMap{row_base, 0u, 1u}, so,
                     false, x, epoch, false);
  if (rc != RAG_OK) return rc;
  if (armed) {
    rc = wait_host_flag(so.done_flag, so.done_seq, c->stream);
    if (rc != RAG_OK) return rc;
  } else {
    CUDA_TRY(cudaMemcpyAsync(c->h_pin + in_b, d + in_b, rows_b + dist_b + cnt_b, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  memcpy(out_rows, c->h_pin + in_b, (size_t)B * k * sizeof(int64_t));
  memcpy(out_dists, c->h_pin + in_b + rows_b, (size_t)B * k * sizeof(float));
  memcpy(out_counts, c->h_pin + in_b + rows_b + dist_b, (size_t)B * sizeof(int32_t));
  for (int b = 0; b < B; ++b)
    if (out_counts[b] < 0) {        // the kernel's way of saying: a peer's slot was stale
      (void)cudaMemsetAsync(x->d_local + kXchgStatusOff, 0, sizeof(uint32_t), c->stream);
      return fail(RAG_ECUDA, "fused exchange: a peer did not deliver its candidates within 20 s; the result is not valid");
    }
  return RAG_OK;
}

// ---- host-buffer queries in flight ----------------------------------------------------------------
static int pipeline_of(rag_store* s, Pipeline** out) {
  if (s->pipe) { *out = s->pipe; return RAG_OK; }
  Pipeline* p = new (std::nothrow) Pipeline();
  if (!p) return fail(RAG_ENOMEM, "out of host memory");
  cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { if (p->stream) cudaStreamDestroy(p->stream); delete p; return fail(RAG_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  int rc = dev_ctx_for(s, p->stream, &p->ctx);
  if (rc != RAG_OK) { cudaStreamDestroy(p->stream); cudaStreamDestroy(p->copy_stream); delete p; return rc; }
  s->pipe = p;
  *out = p;
  return RAG_OK;
}

int rag_store_query_submit(rag_store* s, rag_exchange* x, int B, const float* queries, int k, int mask_slot, int flags,
                           uint32_t row_base, int* ticket) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (!ticket) return fail(RAG_EINVAL, "ticket is NULL");
  int rc = flush_if_pending(s);
  if (rc != RAG_OK) return rc;
  RdLock g(&s->lock);
  rc = x ? fused_check(s, x, B, queries, k, mask_slot, flags) : check_query_args(s, B, queries, k, mask_slot);
  if (rc != RAG_OK) return rc;
  if (B > batch_limit(s, k)) return fail(RAG_EINVAL, "batch %d too large for one submitted query (limit %d)", B, batch_limit(s, k));
  CUDA_TRY(cudaSetDevice(s->device));
  Pipeline* p = nullptr;
  {
    static std::mutex create_mu;
    std::lock_guard<std::mutex> lg(create_mu);
    rc = pipeline_of(s, &p);
  }
  if (rc != RAG_OK) return rc;
  std::lock_guard<std::mutex> lg(p->mu);
  const int si = (int)(p->next % Pipeline::kSlots);
  PipeSlot& ps = p->slot[si];
  if (ps.busy) return fail(RAG_EINVAL, "%d queries are already in flight; wait for the oldest first", Pipeline::kSlots);
  // (the checks above fail on every rank or on none.)  The epoch is taken before anything that can fail on one
  // rank only -- see rag_store_query_fused_dev
  const uint32_t epoch = x ? ++x->epoch : 0u;
  ps.B = B; ps.k = k;
  ps.in_b = align_up((size_t)B * s->dim * sizeof(float), 256);
  ps.rows_b = align_up((size_t)B * k * sizeof(int64_t), 256);
  ps.dist_b = align_up((size_t)B * k * sizeof(float), 256);
  ps.cnt_b = align_up((size_t)B * sizeof(int32_t), 256);
  const size_t io_b = ps.in_b + ps.rows_b + ps.dist_b + ps.cnt_b + 256;
  if (io_b > ps.bytes) {
    if (ps.h) cudaFreeHost(ps.h);
    if (ps.d) cudaFree(ps.d);
    ps.h = nullptr; ps.d = nullptr; ps.bytes = 0;
    const size_t want = align_up(std::max(io_b, (size_t)1 << 16), 4096);
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&ps.h), want));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ps.d), want));
    ps.bytes = want;
  }
  if (!ps.ev) CUDA_TRY(cudaEventCreateWithFlags(&ps.ev, cudaEventDisableTiming));
  QueryCtx* c = p->ctx;
  const int regime = x ? 1 : choose_regime(s, B, k, flags);
  if (regime < 0) return fail(RAG_EINVAL, "tensor regime does not support this store/query");
  if (s->live == 0 && !x) {          // nothing to search: the answer is known
    fill_empty(B, k, reinterpret_cast<int64_t*>(ps.h + ps.in_b), reinterpret_cast<float*>(ps.h + ps.in_b + ps.rows_b),
               reinterpret_cast<int32_t*>(ps.h + ps.in_b + ps.rows_b + ps.dist_b));
    ps.armed = true; ps.seq = 1;
    *reinterpret_cast<volatile uint32_t*>(ps.h + io_b - 256) = 1u;
    ps.busy = true; p->next++; *ticket = si;
    return RAG_OK;
  }
  const int grid_x = scan_stream_grid_x(s->sm_count, s->rows);
  rc = c->ensure_dev(search_scratch_bytes(s, B, k, grid_x));
  if (rc != RAG_OK) return rc;
  // The queries of a small stream-regime batch go down on the COPY stream, followed by an arrival flag the kernel
  // waits for: no operation then sits between two searches on the pipeline stream, which would undo their
  // programmatic overlap (1.25M x 768 shard, 2 in flight: 0.287 ms per query; plain copy on the pipeline stream 0.304;
  // query in the launch parameters 0.290 -- that variant also made a lone blocking query slower and was removed).
  static const bool side_on = !(getenv("RAG_B200_INFLIGHT_QUERY") && atoi(getenv("RAG_B200_INFLIGHT_QUERY")) == 0);
  const bool fused_stream = regime == 1 && direct_host_ok(s, B, k, regime);
  const bool side = side_on && fused_stream;
  SearchOut so{};
  if (side) {
    uint32_t* h_flag = reinterpret_cast<uint32_t*>(ps.h + io_b - 128);       // pinned source of the flag value
    uint32_t* d_flag = reinterpret_cast<uint32_t*>(ps.d + io_b - 128);
    ps.qseq = ++ps.qseq ? ps.qseq : ++ps.qseq;
    memcpy(ps.h, queries, (size_t)B * s->dim * sizeof(float));
    *h_flag = ps.qseq;
    CUDA_TRY(cudaMemcpyAsync(ps.d, ps.h, (size_t)B * s->dim * sizeof(float), cudaMemcpyHostToDevice, p->copy_stream));
    CUDA_TRY(cudaMemcpyAsync(d_flag, h_flag, sizeof(uint32_t), cudaMemcpyHostToDevice, p->copy_stream));   // stream-ordered behind the queries
    so.query_flag = d_flag;
    so.query_seq = ps.qseq;
  } else {
    memcpy(ps.h, queries, (size_t)B * s->dim * sizeof(float));
    CUDA_TRY(cudaMemcpyAsync(ps.d, ps.h, (size_t)B * s->dim * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  }
  bool armed = false;
  unsigned char* ob = ps.d;
  if (fused_stream) {
    ob = ps.h;
    so.done_flag = reinterpret_cast<uint32_t*>(ps.h + io_b - 256);
    so.done_seq = ++ps.seq ? ps.seq : ++ps.seq;
    so.armed = &armed;
  }
  so.rows = reinterpret_cast<int64_t*>(ob + ps.in_b);
  so.dists = reinterpret_cast<float*>(ob + ps.in_b + ps.rows_b);
  so.counts = reinterpret_cast<int32_t*>(ob + ps.in_b + ps.rows_b + ps.dist_b);
  rc = search_device(s, c, c->d_buf, B, reinterpret_cast<const float*>(ps.d), k, mask_slot, regime, RowMap{row_base, 0u, 1u}, so,
                     false, x, epoch, flags == RAG_QUERY_FORCE_TENSOR);
  if (rc != RAG_OK) return rc;
  if (!armed) {
    CUDA_TRY(cudaMemcpyAsync(ps.h + ps.in_b, ps.d + ps.in_b, ps.rows_b + ps.dist_b + ps.cnt_b, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaEventRecord(ps.ev, c->stream));
  }
  ps.armed = armed;
  ps.busy = true;
  p->next++;
  *ticket = si;
  return RAG_OK;
}

int rag_store_query_wait(rag_store* s, int ticket, int64_t* out_rows, float* out_dists, int32_t* out_counts) {
  if (!s || !s->pipe) return fail(RAG_EINVAL, "no query was submitted on this store");
  if (ticket < 0 || ticket >= Pipeline::kSlots) return fail(RAG_EINVAL, "bad ticket %d", ticket);
  if (!out_rows || !out_dists || !out_counts) return fail(RAG_EINVAL, "output pointer is NULL");
  Pipeline* p = s->pipe;
  PipeSlot& ps = p->slot[ticket];
  if (!ps.busy) return fail(RAG_EINVAL, "ticket %d is not in flight", ticket);
  CUDA_TRY(cudaSetDevice(s->device));
  int rc = RAG_OK;
  if (ps.armed) {
    const size_t io_b = ps.in_b + ps.rows_b + ps.dist_b + ps.cnt_b + 256;
    rc = wait_host_flag(reinterpret_cast<const volatile uint32_t*>(ps.h + io_b - 256), ps.seq, p->stream);
  } else if (cudaEventSynchronize(ps.ev) != cudaSuccess) {
    rc = fail(RAG_ECUDA, "search failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  ps.busy = false;
  if (rc != RAG_OK) return rc;
  memcpy(out_rows, ps.h + ps.in_b, (size_t)ps.B * ps.k * sizeof(int64_t));
  memcpy(out_dists, ps.h + ps.in_b + ps.rows_b, (size_t)ps.B * ps.k * sizeof(float));
  memcpy(out_counts, ps.h + ps.in_b + ps.rows_b + ps.dist_b, (size_t)ps.B * sizeof(int32_t));
  for (int b = 0; b < ps.B; ++b)
    if (out_counts[b] < 0) return fail(RAG_ECUDA, "fused exchange: a peer did not deliver its candidates within 20 s; the result is not valid");
  return RAG_OK;
}

int rag_merge_keys_dev(int device, int G, int B, int k, const uint64_t* keys_dev, uint64_t* out_keys_dev,
                       int64_t* out_rows_dev, float* out_dists_dev, int32_t* out_counts_dev, void* stream) {
  if (G <= 0 || B <= 0 || k < 1 || k > RAG_MAX_K || !keys_dev) return fail(RAG_EINVAL, "bad merge arguments");
  NvtxRange nvtx("rag:merge");
  CUDA_TRY(cudaSetDevice(device));
  MergeArgs ma{};
  ma.keys = keys_dev; ma.S = G; ma.B = B; ma.k = k;
  ma.out_keys = out_keys_dev; ma.out_rows = out_rows_dev; ma.out_dists = out_dists_dev; ma.out_counts = out_counts_dev;
  CUDA_TRY(launch_merge(ma, reinterpret_cast<cudaStream_t>(stream)));
  return RAG_OK;
}

uint64_t rag_key_pack(float dist, uint32_t row) { return make_key(dist, row); }
float rag_key_dist(uint64_t key) { return key_dist(key); }
uint32_t rag_key_row(uint64_t key) { return key_row(key); }

int rag_debug_tensor_stats(uint64_t* out8, int reset) {
  static_assert(sizeof(uint64_t) == sizeof(unsigned long long), "counter width");
  return tensor::read_stats(reinterpret_cast<unsigned long long*>(out8), reset) == 0 ? RAG_OK : fail(RAG_ECUDA, "reading the counters failed");
}

float rag_store_last_upsert_ms(const rag_store* s) { return s ? s->last_upsert_ms : 0.0f; }

int rag_store_f32_tensor_info(const rag_store* s, int* shadow_kind, int64_t* queries, int64_t* reruns) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (shadow_kind) {
    const int want = s->want_shadow_kind.load(std::memory_order_acquire);
    *shadow_kind = (s->dtype == RAG_DTYPE_F32) ? (want ? want : s->shadow_kind) : 0;
  }
  if (queries) *queries = s->f32_tensor_queries.load();
  if (reruns) *reruns = s->f32_tensor_reruns.load();
  return RAG_OK;
}

int rag_store_set_f32_shadow(rag_store* s, int kind) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (s->dtype != RAG_DTYPE_F32) return fail(RAG_EINVAL, "only fp32 stores have a bf16 shadow");
  if (kind != RAG_F32_SHADOW_AUTO && kind != RAG_F32_SHADOW_HI && kind != RAG_F32_SHADOW_HILO)
    return fail(RAG_EINVAL, "unknown shadow kind %d", kind);
  int k = kind;
  if (kind == RAG_F32_SHADOW_AUTO) k = tensor::default_shadow_kind(s->row_elems);
  else if (!tensor::supported(s->dtype, s->row_elems, 1, s->space, 0, kind))
    return fail(RAG_EINVAL, "the tensor regime cannot take a %d-element fp32 row through shadow kind %d", s->row_elems, kind);
  s->shadow_pinned.store(kind != RAG_F32_SHADOW_AUTO);
  s->hi_window_queries = 0; s->hi_window_reruns = 0;
  if (k == 0) return RAG_OK;
  s->want_shadow_kind.store(k, std::memory_order_release);
  return flush_if_pending(s);
}

int rag_debug_ring_plan(int stages_available, int stages_per_tile, int accumulators, int* stages_used, int* one_issuer) {
  if (!stages_used || !one_issuer || stages_per_tile < 1 || stages_available < stages_per_tile + 1 ||
      (accumulators != 2 && accumulators != 4))
    return fail(RAG_EINVAL, "bad ring geometry");
  tensor::plan_ring(stages_available, stages_per_tile, accumulators, stages_used, one_issuer);
  return RAG_OK;
}

int rag_store_last_query_info(const rag_store* s, float* kernel_ms, int* regime, int* launches) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  rag_store* ms_ = const_cast<rag_store*>(s);
  if (QueryCtx* c = ms_->timing_ctx.exchange(nullptr, std::memory_order_acq_rel)) {
    float ms = 0.0f;      // contexts live as long as the store; a concurrent query on the same context only blurs the number
    if (cudaSetDevice(s->device) == cudaSuccess && cudaEventSynchronize(c->ev1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) ms_->last_kernel_ms = ms;
    else (void)cudaGetLastError();
  }
  if (kernel_ms) *kernel_ms = s->last_kernel_ms;
  if (regime) *regime = s->last_regime.load();
  if (launches) *launches = s->last_launches.load();
  return RAG_OK;
}

}  // extern "C"
