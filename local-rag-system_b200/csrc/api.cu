// C ABI of the engine (include/rag_b200.h): device-resident corpus store + search.
// Host-side bookkeeping only; all arithmetic is in the kernels.  There is no CPU
// fallback: without a usable sm_100 device every entry point that would compute
// returns RAG_ENODEV / RAG_ECUDA.
#include "../../include/rag_b200.h"

#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "tensor_regime.h"

using namespace rag;

// ------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      (void)cudaGetLastError();                                                              \
      return fail(_e == cudaErrorMemoryAllocation ? RAG_ENOMEM : RAG_ECUDA, "%s: %s (%s:%d)", \
                  #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                        \
    }                                                                                        \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------------------
// per-query scratch: one stream + pinned staging + device scratch
// ------------------------------------------------------------------------------
struct QueryCtx {
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  unsigned char* h_pin = nullptr;
  size_t h_bytes = 0;
  unsigned char* d_buf = nullptr;
  size_t d_bytes = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  unsigned int* d_tickets = nullptr;     // kMaxTickets zeroed counters, self-resetting (scan kernel)
  static constexpr int kMaxTickets = 4096;

  int ensure_tickets() {
    if (d_tickets) return RAG_OK;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_tickets), kMaxTickets * sizeof(unsigned int)));
    CUDA_TRY(cudaMemsetAsync(d_tickets, 0, kMaxTickets * sizeof(unsigned int), stream));
    return RAG_OK;
  }
  int ensure_host(size_t bytes) {
    if (bytes <= h_bytes) return RAG_OK;
    if (h_pin) cudaFreeHost(h_pin);
    h_pin = nullptr; h_bytes = 0;
    size_t want = align_up(std::max(bytes, (size_t)1 << 16), 4096);
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&h_pin), want));
    h_bytes = want;
    return RAG_OK;
  }
  int ensure_dev(size_t bytes) {
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
  void destroy() {
    if (stream && own_stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
    if (h_pin) cudaFreeHost(h_pin);
    if (d_buf) cudaFree(d_buf);
    if (d_tickets) cudaFree(d_tickets);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  }
};

// ------------------------------------------------------------------------------
// the store
// ------------------------------------------------------------------------------
struct rag_store {
  int dim = 0, dtype = 0, space = 0, device = 0;
  int row_elems = 0;        // dim padded so that a row is a whole number of 16-byte chunks
  size_t row_bytes = 0;
  int sm_count = 0;
  int64_t capacity = 0;     // rows allocated (multiple of 32)
  int64_t rows = 0;         // high-water mark
  int64_t live = 0;
  void* d_vectors = nullptr;
  float* d_norms2 = nullptr;
  float* d_max_norm2 = nullptr;          // [2] largest / smallest |stored row|^2 ever written (error bound of the split regime, l2 rejection bound)
  uint32_t* d_live = nullptr;
  // fp32 stores, tensor regime: bf16 [capacity][hi(row_elems) | lo(row_elems)] split of the rows, built on
  // the first large-batch query, kept in step by upsert, dropped (and rebuilt lazily) when the store grows
  __nv_bfloat16* d_shadow = nullptr;
  std::mutex shadow_mu;
  uint32_t* d_masks[RAG_MAX_MASK_SLOTS] = {};
  int64_t mask_words[RAG_MAX_MASK_SLOTS] = {};
  std::vector<uint32_t> h_live;
  std::vector<int64_t> free_rows;
  pthread_rwlock_t lock;
  // scratch pool for the synchronous API
  std::mutex pool_mu;
  std::condition_variable pool_cv;
  std::vector<QueryCtx*> pool_free;
  int pool_created = 0;
  static constexpr int kMaxPool = 8;
  // scratch for the asynchronous API, one per caller stream
  std::mutex dev_mu;
  std::unordered_map<void*, QueryCtx*> dev_ctx;
  QueryCtx admin;           // upsert / delete / fetch (used under the write lock)
  tensor::Plan* tensor_plan = nullptr;
  std::atomic<int64_t> launches{0};
  std::atomic<int> last_regime{0};
  std::atomic<int> last_launches{0};
  float last_kernel_ms = 0.0f;
};

// peer-mapped buffers for the fused scan + all-gather + merge launch (multi-GPU, one process per GPU)
struct rag_exchange {
  int device = 0, rank = 0, world = 1;
  int64_t slot_keys = 0;
  size_t bytes = 0;
  unsigned char* d_local = nullptr;
  std::vector<unsigned char*> peers;     // [world] base of every rank's buffer as mapped into this process
  unsigned char** d_peers = nullptr;     // the same table on the device
  uint32_t epoch = 0;
  bool connected = false;
};

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
      if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
      if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
      if (e != cudaSuccess) {
        c->destroy();
        delete c;
        g.lock();
        s->pool_created--;
        return fail(RAG_ECUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
      }
      c->own_stream = true;
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

// grow device arrays to hold at least `need` rows (write lock held, device idle)
int grow(rag_store* s, int64_t need) {
  if (need <= s->capacity) return RAG_OK;
  if (need > 0xFFFFFFF0ll) return fail(RAG_EINVAL, "a store holds at most 2^32-16 rows");
  int64_t cap = std::max<int64_t>(need, std::max<int64_t>(1024, s->capacity * 2));
  cap = (cap + 31) / 32 * 32;
  void* nv = nullptr;
  float* nn = nullptr;
  uint32_t* nl = nullptr;
  cudaError_t e = cudaMalloc(&nv, (size_t)cap * s->row_bytes);
  if (e != cudaSuccess && cap > need) {   // doubling did not fit: take exactly what is needed
    (void)cudaGetLastError();
    cap = (need + 31) / 32 * 32;
    e = cudaMalloc(&nv, (size_t)cap * s->row_bytes);
  }
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(RAG_ENOMEM, "cudaMalloc of %zu bytes for the corpus failed: %s", (size_t)cap * s->row_bytes, cudaGetErrorString(e)); }
  e = cudaMalloc(reinterpret_cast<void**>(&nn), (size_t)cap * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&nl), (size_t)cap / 32 * sizeof(uint32_t));
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(nv); if (nn) cudaFree(nn);
    return fail(RAG_ENOMEM, "cudaMalloc for side arrays failed: %s", cudaGetErrorString(e));
  }
  cudaStream_t st = s->admin.stream;
  CUDA_TRY(cudaMemsetAsync(nl, 0, (size_t)cap / 32 * sizeof(uint32_t), st));
  if (s->rows > 0) {
    CUDA_TRY(cudaMemcpyAsync(nv, s->d_vectors, (size_t)s->rows * s->row_bytes, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(nn, s->d_norms2, (size_t)s->rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(nl, s->d_live, (size_t)((s->rows + 31) / 32) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  if (s->d_vectors) cudaFree(s->d_vectors);
  if (s->d_norms2) cudaFree(s->d_norms2);
  if (s->d_live) cudaFree(s->d_live);
  s->d_vectors = nv; s->d_norms2 = nn; s->d_live = nl;
  if (s->d_shadow) { cudaFree(s->d_shadow); s->d_shadow = nullptr; }
  s->capacity = cap;
  s->h_live.resize((size_t)cap / 32, 0u);
  if (s->tensor_plan) tensor::invalidate(s->tensor_plan);
  return RAG_OK;
}

// assign destination rows for an upsert (write lock held)
int assign_rows(rag_store* s, int64_t n, const int64_t* rows, std::vector<int64_t>& dst) {
  dst.resize((size_t)n);
  int64_t hwm = s->rows;
  for (int64_t i = 0; i < n; ++i) {
    int64_t r = rows ? rows[i] : -1;
    if (r >= 0) {
      if (r >= hwm) return fail(RAG_EINVAL, "upsert row %lld is beyond the store's %lld rows", (long long)r, (long long)hwm);
    } else {
      r = -1;
      while (!s->free_rows.empty()) {
        int64_t c = s->free_rows.back();
        s->free_rows.pop_back();
        if (!h_is_live(s, c) && c < s->rows) { r = c; break; }
      }
      if (r < 0) r = hwm++;
    }
    dst[(size_t)i] = r;
  }
  int rc = grow(s, hwm);
  if (rc != RAG_OK) return rc;
  return RAG_OK;
}

void mark_live(rag_store* s, const std::vector<int64_t>& dst) {
  for (int64_t r : dst) {
    if (r >= s->rows) s->rows = r + 1;
    uint32_t& w = s->h_live[(size_t)(r >> 5)];
    uint32_t bit = 1u << (r & 31);
    if (!(w & bit)) { w |= bit; s->live++; }
  }
}

bool contiguous(const std::vector<int64_t>& v) {
  for (size_t i = 1; i < v.size(); ++i)
    if (v[i] != v[0] + (int64_t)i) return false;
  return true;
}

// run the upsert kernel for vectors already on the device (write lock held)
int upsert_device_chunk(rag_store* s, const float* d_src, int64_t n, const int64_t* dst_rows_host, bool contig) {
  QueryCtx& c = s->admin;
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
  a.shadow = s->d_shadow;
  if (contig) {
    a.rows = nullptr;
    a.row0 = dst_rows_host[0];
  } else {
    int64_t* d_rows = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_rows), (size_t)n * sizeof(int64_t)));
    cudaError_t e = cudaMemcpyAsync(d_rows, dst_rows_host, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream);
    if (e == cudaSuccess) { a.rows = d_rows; e = launch_upsert(a, c.stream); }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    cudaFree(d_rows);
    CUDA_TRY(e);
    s->launches++;
    return RAG_OK;
  }
  CUDA_TRY(launch_upsert(a, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  s->launches++;
  return RAG_OK;
}

void fill_empty(int B, int k, int64_t* out_rows, float* out_dists, int32_t* out_counts) {
  const float inf = __builtin_inff();
  for (int64_t i = 0; i < (int64_t)B * k; ++i) {
    if (out_rows) out_rows[i] = -1;
    if (out_dists) out_dists[i] = inf;
  }
  if (out_counts) for (int b = 0; b < B; ++b) out_counts[b] = 0;
}

// Decide the kernel regime for a batch.
int choose_regime(const rag_store* s, int B, int k, int flags) {
  if (flags == RAG_QUERY_FORCE_STREAM) return 1;
  const bool tensor_ok = tensor::supported(s->dtype, s->row_elems, k, s->space);
  if (flags == RAG_QUERY_FORCE_TENSOR) return tensor_ok ? 2 : -1;
  if (!tensor_ok) return 1;
  // the stream kernel reads the corpus once per group of <= 8 queries; the tensor kernel
  // once per 128 (HBM-bound up to ~256 queries, tensor-bound beyond)
  if (s->dtype == RAG_DTYPE_F32) return (B > tensor::kStreamMaxBatchF32) ? 2 : 1;
  return (B > tensor::kStreamMaxBatch) ? 2 : 1;
}

// fp32 store about to be searched by the tensor regime: make sure its bf16 hi/lo shadow exists
// (read lock held: no writer is active; concurrent readers serialise on shadow_mu)
int ensure_shadow(rag_store* s, cudaStream_t st) {
  std::lock_guard<std::mutex> g(s->shadow_mu);
  if (s->d_shadow) return RAG_OK;
  __nv_bfloat16* sh = nullptr;
  const size_t bytes = (size_t)s->capacity * 2 * s->row_elems * sizeof(__nv_bfloat16);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&sh), bytes);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(RAG_ENOMEM, "cudaMalloc of %zu bytes for the split-precision shadow failed: %s", bytes, cudaGetErrorString(e)); }
  e = launch_split_rows(reinterpret_cast<const float*>(s->d_vectors), s->row_elems, 0, s->rows, sh, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { (void)cudaGetLastError(); cudaFree(sh); return fail(RAG_ECUDA, "building the split-precision shadow failed: %s", cudaGetErrorString(e)); }
  s->d_shadow = sh;
  s->launches++;
  return RAG_OK;
}

// shard-local search on device buffers: prep -> scan -> merge.  Emits keys and/or rows.
// `scratch` must hold scratch_bytes(); everything is asynchronous on c->stream.
struct SearchOut {
  uint64_t* keys = nullptr;
  int64_t* rows = nullptr;
  float* dists = nullptr;
  int32_t* counts = nullptr;
};

// stream-regime scratch: prepared queries | per-CTA partial lists [grid_x][B][k] | tickets [B]
size_t search_scratch_bytes(const rag_store* s, int B, int k, int grid_x) {
  size_t q = align_up((size_t)B * s->row_elems * sizeof(float), 256);
  size_t part = align_up((size_t)grid_x * B * k * sizeof(uint64_t), 256);
  size_t st = align_up((size_t)(B + 1) * sizeof(int), 256);      // redo count + list (split-precision tensor regime)
  return q + part + st + tensor::scratch_bytes(s->dtype, s->row_elems, B, k, s->sm_count);
}

int search_device(rag_store* s, QueryCtx* c, unsigned char* scratch, int B, const float* d_queries_raw, int k,
                  int mask_slot, int regime, uint32_t row_base, const SearchOut& out, bool timed,
                  rag_exchange* xchg = nullptr, bool forced_tensor = false) {
  cudaStream_t st = c->stream;
  const int grid_x = scan_stream_grid_x(s->sm_count, s->rows);
  float* d_q = reinterpret_cast<float*>(scratch);
  size_t off = align_up((size_t)B * s->row_elems * sizeof(float), 256);
  uint64_t* d_partial = reinterpret_cast<uint64_t*>(scratch + off);
  off += align_up((size_t)grid_x * B * k * sizeof(uint64_t), 256);
  int* d_redo = reinterpret_cast<int*>(scratch + off);       // [0] count, [1..B] query indices
  off += align_up((size_t)(B + 1) * sizeof(int), 256);
  unsigned char* d_tensor = scratch + off;

  const uint32_t* filter = nullptr;
  int64_t fwords = 0;
  if (mask_slot >= 0) {
    filter = s->d_masks[mask_slot];
    fwords = s->mask_words[mask_slot];
  }
  int launches = 0;
  int S = 0;
  tensor::Result tres{};
  if (regime == 2 && s->dtype == RAG_DTYPE_F32) {
    // the split-precision regime needs the hi/lo shadow (as many bytes again as the rows).  If the device
    // cannot hold it, an AUTO query is still answered -- exactly, by the stream kernel, just slower.
    const int rcs = ensure_shadow(s, st);
    if (rcs == RAG_ENOMEM && !forced_tensor) regime = 1;
    else if (rcs != RAG_OK) return rcs;
  }
  if (regime == 2) {
    if (timed) CUDA_TRY(cudaEventRecord(c->ev0, st));
    tensor::Problem p{};
    p.vectors = s->d_vectors; p.shadow = s->d_shadow; p.norms2 = s->d_norms2; p.min_norm2 = s->d_max_norm2 + 1; p.n_rows = s->rows; p.row_elems = s->row_elems;
    p.dim = s->dim; p.dtype = s->dtype; p.space = s->space;
    p.live = s->d_live; p.filter = filter; p.filter_words = fwords;
    p.dense = (filter == nullptr && s->live == s->rows) ? 1 : 0;
    p.queries_raw = d_queries_raw; p.B = B; p.k = k;
    p.scratch = d_tensor; p.sm_count = s->sm_count;
    cudaError_t e = tensor::launch(s->tensor_plan, p, st, &tres, &launches);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(RAG_ECUDA, "tensor-regime launch failed: %s", cudaGetErrorString(e)); }
    d_partial = const_cast<uint64_t*>(tres.partial);
    S = tres.S;
    if (timed) CUDA_TRY(cudaEventRecord(c->ev1, st));
  } else {
    static const bool fused = !(getenv("RAG_B200_FUSED_MERGE") && atoi(getenv("RAG_B200_FUSED_MERGE")) == 0);
    if (fused) {
      int rc2 = c->ensure_tickets();
      if (rc2 != RAG_OK) return rc2;
      if (B > QueryCtx::kMaxTickets) return fail(RAG_EINVAL, "batch %d exceeds %d", B, QueryCtx::kMaxTickets);
    } else {          // unfused debugging mode: separate prep and merge kernels
      PrepArgs pa{};
      pa.src = d_queries_raw; pa.B = B; pa.dim = s->dim; pa.row_elems = s->row_elems;
      pa.normalise = (s->space == RAG_SPACE_COSINE);
      pa.round_bf16 = (s->dtype == RAG_DTYPE_BF16);
      pa.q_f32 = d_q; pa.q_bf16 = nullptr; pa.q_norm2 = nullptr;
      pa.init_keys = nullptr; pa.init_keys_n = 0; pa.init_zero = nullptr; pa.init_zero_n = 0;
      CUDA_TRY(launch_prep_queries(pa, st));
      launches++;
    }
    ScanArgs sa{};
    sa.vectors = s->d_vectors; sa.dtype = s->dtype; sa.row_elems = s->row_elems;
    sa.cpr = (int)(s->row_bytes / 16);
    sa.n_rows = s->rows; sa.live = s->d_live; sa.filter = filter; sa.filter_words = fwords;
    sa.B = B; sa.k = k; sa.l2 = (s->space == RAG_SPACE_L2);
    sa.grid_x = grid_x;
    sa.partial = d_partial; sa.done = fused ? c->d_tickets : nullptr; sa.merge_keys_cap = 0;
    sa.queries = fused ? nullptr : d_q;
    sa.queries_raw = fused ? d_queries_raw : nullptr;
    sa.dim = s->dim; sa.normalise = (s->space == RAG_SPACE_COSINE); sa.round_bf16 = (s->dtype == RAG_DTYPE_BF16);
    sa.row_base = row_base;
    sa.out_keys = out.keys; sa.out_rows = out.rows; sa.out_dists = out.dists; sa.out_counts = out.counts;
    if (xchg != nullptr) {
      if (!fused) return fail(RAG_EINVAL, "the fused exchange needs the fused merge (RAG_B200_FUSED_MERGE=0 is set)");
      sa.xchg_peers = xchg->d_peers; sa.xchg_rank = xchg->rank; sa.xchg_world = xchg->world;
      sa.xchg_epoch = ++xchg->epoch; sa.xchg_slot_keys = xchg->slot_keys;
    }
    if (timed) CUDA_TRY(cudaEventRecord(c->ev0, st));
    CUDA_TRY(launch_scan_stream(sa, s->sm_count, st, &launches));
    if (timed) CUDA_TRY(cudaEventRecord(c->ev1, st));
    if (fused) {   // the scan kernel merged across CTAs and emitted the result itself
      s->launches += launches;
      s->last_launches = launches;
      s->last_regime = regime;
      return RAG_OK;
    }
    S = grid_x;
  }
  const bool split = (regime == 2 && s->dtype == RAG_DTYPE_F32);
  const int k_kept = (regime == 2) ? tres.k_kept : k;
  MergeArgs ma{};
  ma.keys = d_partial; ma.S = S; ma.B = B; ma.k = k_kept; ma.row_base = row_base;
  ma.out_keys = out.keys; ma.out_rows = out.rows; ma.out_dists = out.dists; ma.out_counts = out.counts;
  const bool refine = (regime == 2 && (s->space == RAG_SPACE_L2 || split));
  if (refine) {   // merge to scratch keys first, then re-score the winners exactly
    ma.row_base = 0; ma.out_keys = tres.merged; ma.out_rows = nullptr; ma.out_dists = nullptr; ma.out_counts = nullptr;
  }
  CUDA_TRY(launch_merge(ma, st));
  launches++;
  if (refine) {
    RefineArgs ra{};
    ra.keys = tres.merged; ra.vectors = s->d_vectors; ra.queries = tres.q_f32;
    ra.dtype = s->dtype; ra.row_elems = s->row_elems; ra.B = B; ra.k = k; ra.k_in = k_kept;
    ra.l2 = (s->space == RAG_SPACE_L2) ? 1 : 0; ra.row_base = row_base;
    ra.out_keys = out.keys; ra.out_rows = out.rows; ra.out_dists = out.dists; ra.out_counts = out.counts;
    if (split) {
      CUDA_TRY(cudaMemsetAsync(d_redo, 0, sizeof(int), st));
      ra.guard_rel = 1.2e-4f; ra.q_norm2 = tres.q_norm2; ra.x_max_norm2 = s->d_max_norm2;
      ra.redo_count = d_redo; ra.redo_list = d_redo + 1;
    }
    CUDA_TRY(launch_refine(ra, st));
    launches++;
    if (split) {
      // queries whose top-k the approximate ranking could not certify are re-run on the exact fp32
      // stream kernel; the launch covers the worst case and exits at once when the list is empty
      int rc2 = c->ensure_tickets();
      if (rc2 != RAG_OK) return rc2;
      if (B > QueryCtx::kMaxTickets) return fail(RAG_EINVAL, "batch %d exceeds %d", B, QueryCtx::kMaxTickets);
      ScanArgs sa{};
      sa.vectors = s->d_vectors; sa.dtype = s->dtype; sa.row_elems = s->row_elems;
      sa.cpr = (int)(s->row_bytes / 16);
      sa.n_rows = s->rows; sa.live = s->d_live; sa.filter = filter; sa.filter_words = fwords;
      sa.B = B; sa.k = k; sa.l2 = (s->space == RAG_SPACE_L2);
      sa.grid_x = grid_x;
      sa.partial = reinterpret_cast<uint64_t*>(scratch + align_up((size_t)B * s->row_elems * sizeof(float), 256));
      sa.done = c->d_tickets; sa.merge_keys_cap = 0;
      sa.queries = nullptr; sa.queries_raw = d_queries_raw;
      sa.dim = s->dim; sa.normalise = (s->space == RAG_SPACE_COSINE); sa.round_bf16 = 0;
      sa.row_base = row_base;
      sa.out_keys = out.keys; sa.out_rows = out.rows; sa.out_dists = out.dists; sa.out_counts = out.counts;
      sa.q_count = d_redo; sa.q_index = d_redo + 1;
      CUDA_TRY(launch_scan_stream(sa, s->sm_count, st, &launches));
    }
  }
  s->launches += launches;
  s->last_launches = launches;
  s->last_regime = regime;
  return RAG_OK;
}

// scratch of the asynchronous API: one context per caller stream
int dev_ctx_for(rag_store* s, void* stream, QueryCtx** out) {
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
int batch_limit(const rag_store* s, int k) {
  const int grid_x = scan_stream_grid_x(s->sm_count, s->rows);
  const size_t budget = (size_t)512 << 20;
  int64_t lim = (int64_t)(budget / ((size_t)grid_x * k * sizeof(uint64_t)));
  if (lim < 1) lim = 1;
  if (lim > 4096) lim = 4096;
  return (int)lim;
}

}  // namespace

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

int rag_store_create(int dim, int dtype, int space, int device, int64_t capacity_hint, rag_store** out) {
  if (!out) return fail(RAG_EINVAL, "out is NULL");
  *out = nullptr;
  if (dim <= 0 || dim > 65536) return fail(RAG_EINVAL, "dim must be in [1, 65536], got %d", dim);
  if (dtype != RAG_DTYPE_F32 && dtype != RAG_DTYPE_BF16) return fail(RAG_EINVAL, "unknown dtype %d", dtype);
  if (space < RAG_SPACE_L2 || space > RAG_SPACE_IP) return fail(RAG_EINVAL, "unknown space %d", space);
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
  s->sm_count = prop.multiProcessorCount;
  pthread_rwlock_init(&s->lock, nullptr);
  e = cudaStreamCreateWithFlags(&s->admin.stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete s; return fail(RAG_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  s->admin.own_stream = true;
  e = cudaMalloc(reinterpret_cast<void**>(&s->d_max_norm2), 2 * sizeof(float));
  if (e == cudaSuccess) {
    const float init[2] = {0.0f, __builtin_inff()};      // [0] running max, [1] running min of |stored row|^2
    e = cudaMemcpy(s->d_max_norm2, init, sizeof(init), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) { (void)cudaGetLastError(); rag_store_destroy(s); return fail(RAG_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); }
  s->tensor_plan = tensor::create_plan();
  int rc = grow(s, std::max<int64_t>(capacity_hint, 1024));
  if (rc != RAG_OK) { rag_store_destroy(s); return rc; }
  *out = s;
  return RAG_OK;
}

int rag_store_destroy(rag_store* s) {
  if (!s) return RAG_OK;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();
  for (QueryCtx* c : s->pool_free) { c->destroy(); delete c; }
  for (auto& kv : s->dev_ctx) { kv.second->destroy(); delete kv.second; }
  s->admin.destroy();
  if (s->tensor_plan) tensor::destroy_plan(s->tensor_plan);
  for (int i = 0; i < RAG_MAX_MASK_SLOTS; ++i) if (s->d_masks[i]) cudaFree(s->d_masks[i]);
  if (s->d_vectors) cudaFree(s->d_vectors);
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
  CUDA_TRY(cudaDeviceSynchronize());
  return grow(s, rows);
}

static int upsert_impl(rag_store* s, int64_t n, const float* vectors, bool on_device, const int64_t* rows, int64_t* out_rows) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (n < 0) return fail(RAG_EINVAL, "n < 0");
  if (n == 0) return RAG_OK;
  if (!vectors) return fail(RAG_EINVAL, "vectors is NULL");
  WrLock g(&s->lock);
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaDeviceSynchronize());   // no reader (sync or async) may be in flight while rows change
  std::vector<int64_t> dst;
  int rc = assign_rows(s, n, rows, dst);
  if (rc != RAG_OK) return rc;
  const bool contig = contiguous(dst);
  if (on_device) {
    rc = upsert_device_chunk(s, vectors, n, dst.data(), contig);
    if (rc != RAG_OK) return rc;
  } else {
    // stage through pinned memory in chunks of <= 64 MB
    const size_t row_in = (size_t)s->dim * sizeof(float);
    int64_t per = std::max<int64_t>(1, (int64_t)(((size_t)64 << 20) / row_in));
    per = std::min<int64_t>(per, n);
    QueryCtx& c = s->admin;
    rc = c.ensure_host((size_t)per * row_in);
    if (rc != RAG_OK) return rc;
    rc = c.ensure_dev((size_t)per * row_in);
    if (rc != RAG_OK) return rc;
    for (int64_t i0 = 0; i0 < n; i0 += per) {
      const int64_t m = std::min<int64_t>(per, n - i0);
      memcpy(c.h_pin, vectors + (size_t)i0 * s->dim, (size_t)m * row_in);
      CUDA_TRY(cudaMemcpyAsync(c.d_buf, c.h_pin, (size_t)m * row_in, cudaMemcpyHostToDevice, c.stream));
      rc = upsert_device_chunk(s, reinterpret_cast<const float*>(c.d_buf), m, dst.data() + i0, contig);
      if (rc != RAG_OK) return rc;
    }
  }
  mark_live(s, dst);
  if (out_rows) memcpy(out_rows, dst.data(), (size_t)n * sizeof(int64_t));
  return RAG_OK;
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
  WrLock g(&s->lock);
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaDeviceSynchronize());
  std::vector<int64_t> victims;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t r = rows[i];
    if (!h_is_live(s, r)) continue;
    s->h_live[(size_t)(r >> 5)] &= ~(1u << (r & 31));
    s->live--;
    s->free_rows.push_back(r);
    victims.push_back(r);
  }
  if (victims.empty()) return RAG_OK;
  QueryCtx& c = s->admin;
  int rc = c.ensure_dev(victims.size() * sizeof(int64_t));
  if (rc != RAG_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(c.d_buf, victims.data(), victims.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
  CUDA_TRY(launch_clear_live(s->d_live, reinterpret_cast<const int64_t*>(c.d_buf), (int64_t)victims.size(), c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  s->launches++;
  return RAG_OK;
}

int64_t rag_store_count(const rag_store* s) { return s ? s->live : 0; }
int64_t rag_store_rows(const rag_store* s) { return s ? s->rows : 0; }
int64_t rag_store_capacity(const rag_store* s) { return s ? s->capacity : 0; }
int rag_store_dim(const rag_store* s) { return s ? s->dim : 0; }
int rag_store_dtype(const rag_store* s) { return s ? s->dtype : 0; }
int rag_store_space(const rag_store* s) { return s ? s->space : 0; }
int rag_store_device(const rag_store* s) { return s ? s->device : -1; }
int64_t rag_store_kernel_launches(const rag_store* s) { return s ? s->launches.load() : 0; }

int rag_store_is_live(const rag_store* s, int64_t row) {
  if (!s) return 0;
  RdLock g(const_cast<pthread_rwlock_t*>(&s->lock));
  return h_is_live(s, row) ? 1 : 0;
}

int rag_store_fetch(rag_store* s, int64_t n, const int64_t* rows, float* out) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (n <= 0) return RAG_OK;
  if (!rows || !out) return fail(RAG_EINVAL, "rows/out is NULL");
  WrLock g(&s->lock);   // uses the admin scratch
  CUDA_TRY(cudaSetDevice(s->device));
  for (int64_t i = 0; i < n; ++i)
    if (rows[i] < 0 || rows[i] >= s->rows) return fail(RAG_EINVAL, "fetch row %lld out of range", (long long)rows[i]);
  QueryCtx& c = s->admin;
  const size_t rb = align_up((size_t)n * sizeof(int64_t), 256);
  const size_t ob = (size_t)n * s->dim * sizeof(float);
  int rc = c.ensure_dev(rb + ob);
  if (rc != RAG_OK) return rc;
  int64_t* d_rows = reinterpret_cast<int64_t*>(c.d_buf);
  float* d_out = reinterpret_cast<float*>(c.d_buf + rb);
  CUDA_TRY(cudaMemcpyAsync(d_rows, rows, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, c.stream));
  CUDA_TRY(launch_fetch(s->d_vectors, s->dtype, s->dim, s->row_elems, d_rows, n, d_out, c.stream));
  CUDA_TRY(cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  s->launches++;
  return RAG_OK;
}

int rag_store_set_mask(rag_store* s, int slot, const uint64_t* bits, int64_t nbits) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (slot < 0 || slot >= RAG_MAX_MASK_SLOTS) return fail(RAG_EINVAL, "mask slot %d out of range", slot);
  if (nbits < 0 || (nbits > 0 && !bits)) return fail(RAG_EINVAL, "bad mask arguments");
  WrLock g(&s->lock);
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaDeviceSynchronize());
  if (s->d_masks[slot]) { cudaFree(s->d_masks[slot]); s->d_masks[slot] = nullptr; s->mask_words[slot] = 0; }
  const int64_t words64 = (nbits + 63) / 64;
  const int64_t words32 = words64 * 2;
  if (words32 == 0) {       // an empty mask: nothing passes
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_masks[slot]), 8));
    CUDA_TRY(cudaMemset(s->d_masks[slot], 0, 8));
    s->mask_words[slot] = 0;
    return RAG_OK;
  }
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_masks[slot]), (size_t)words32 * 4));
  std::vector<uint64_t> tmp(bits, bits + words64);
  if (nbits % 64) tmp[(size_t)words64 - 1] &= (~0ull >> (64 - nbits % 64));   // bits past nbits do not pass
  CUDA_TRY(cudaMemcpy(s->d_masks[slot], tmp.data(), (size_t)words64 * 8, cudaMemcpyHostToDevice));
  s->mask_words[slot] = words32;
  return RAG_OK;
}

int rag_store_clear_mask(rag_store* s, int slot) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (slot < 0 || slot >= RAG_MAX_MASK_SLOTS) return fail(RAG_EINVAL, "mask slot %d out of range", slot);
  WrLock g(&s->lock);
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaDeviceSynchronize());
  if (s->d_masks[slot]) { cudaFree(s->d_masks[slot]); s->d_masks[slot] = nullptr; }
  s->mask_words[slot] = 0;
  return RAG_OK;
}

static int check_query_args(const rag_store* s, int B, const void* q, int k, int mask_slot) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (B <= 0) return fail(RAG_EINVAL, "query batch must be >= 1, got %d", B);
  if (!q) return fail(RAG_EINVAL, "queries is NULL");
  if (k < 1 || k > RAG_MAX_K) return fail(RAG_EINVAL, "k must be in [1, %d], got %d", RAG_MAX_K, k);
  if (mask_slot >= RAG_MAX_MASK_SLOTS) return fail(RAG_EINVAL, "mask slot %d out of range", mask_slot);
  if (mask_slot >= 0 && !s->d_masks[mask_slot]) return fail(RAG_EINVAL, "mask slot %d is not set", mask_slot);
  return RAG_OK;
}

int rag_store_query(rag_store* s, int B, const float* queries, int k, int mask_slot, int flags,
                    int64_t* out_rows, float* out_dists, int32_t* out_counts) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  RdLock g(&s->lock);
  int rc = check_query_args(s, B, queries, k, mask_slot);
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
    regime_used = regime;
    const int grid_x = scan_stream_grid_x(s->sm_count, s->rows);
    const size_t in_b = align_up((size_t)Bc * s->dim * sizeof(float), 256);
    const size_t rows_b = align_up((size_t)Bc * k * sizeof(int64_t), 256);
    const size_t dist_b = align_up((size_t)Bc * k * sizeof(float), 256);
    const size_t cnt_b = align_up((size_t)Bc * sizeof(int32_t), 256);
    const size_t scr_b = search_scratch_bytes(s, Bc, k, grid_x);
    rc = c->ensure_host(in_b + rows_b + dist_b + cnt_b);
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
    rc = search_device(s, c, scratch, Bc, d_in, k, mask_slot, regime, 0u, so, true, nullptr, flags == RAG_QUERY_FORCE_TENSOR);
    if (rc != RAG_OK) return rc;
    regime_used = s->last_regime.load();     // the regime that actually ran (an fp32 store may have fallen back)
    // one D2H for rows + dists + counts (contiguous in the scratch and in the pinned buffer)
    CUDA_TRY(cudaMemcpyAsync(c->h_pin + in_b, d + in_b, rows_b + dist_b + cnt_b, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out_rows + (size_t)b0 * k, c->h_pin + in_b, (size_t)Bc * k * sizeof(int64_t));
    memcpy(out_dists + (size_t)b0 * k, c->h_pin + in_b + rows_b, (size_t)Bc * k * sizeof(float));
    memcpy(out_counts + b0, c->h_pin + in_b + rows_b + dist_b, (size_t)Bc * sizeof(int32_t));
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) total_ms += ms;
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
  RdLock g(&s->lock);
  int rc = check_query_args(s, B, queries_dev, k, mask_slot);
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
  return search_device(s, c, c->d_buf, B, queries_dev, k, mask_slot, regime, row_base, so, false, nullptr, flags == RAG_QUERY_FORCE_TENSOR);
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
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    if (x->d_local) cudaFree(x->d_local);
    delete x;
    return fail(RAG_ENOMEM, "exchange buffer allocation failed: %s", cudaGetErrorString(e));
  }
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
  return RAG_OK;
}

int rag_exchange_status(rag_exchange* x, int* timed_out) {
  if (!x || !timed_out) return fail(RAG_EINVAL, "exchange/out is NULL");
  CUDA_TRY(cudaSetDevice(x->device));
  uint32_t st = 0;
  CUDA_TRY(cudaMemcpy(&st, x->d_local + kXchgStatusOff, sizeof(st), cudaMemcpyDeviceToHost));
  *timed_out = st != 0 ? 1 : 0;
  return RAG_OK;
}

int rag_exchange_destroy(rag_exchange* x) {
  if (!x) return RAG_OK;
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  for (int g = 0; g < x->world; ++g)
    if (g != x->rank && x->peers[(size_t)g]) cudaIpcCloseMemHandle(x->peers[(size_t)g]);
  if (x->d_peers) cudaFree(x->d_peers);
  if (x->d_local) cudaFree(x->d_local);
  delete x;
  return RAG_OK;
}

int rag_store_fused_ok(const rag_store* s, const rag_exchange* x, int B, int k, int flags) {
  if (!s || !x || !x->connected || B < 1 || k < 1 || k > 128) return 0;
  if (choose_regime(s, B, k, flags) != 1) return 0;
  if ((int64_t)B * k > x->slot_keys) return 0;
  if (scan_stream_groups(B, s->dtype, s->row_elems, k) > kXchgMaxGroups) return 0;
  if (B > QueryCtx::kMaxTickets || B > batch_limit(s, k)) return 0;
  return 1;
}

int rag_store_query_fused_dev(rag_store* s, rag_exchange* x, int B, const float* queries_dev, int k, int mask_slot,
                              int flags, uint32_t row_base, int64_t* out_rows_dev, float* out_dists_dev,
                              int32_t* out_counts_dev, void* stream) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (!x || !x->connected) return fail(RAG_EINVAL, "exchange is NULL or not connected");
  RdLock g(&s->lock);
  int rc = check_query_args(s, B, queries_dev, k, mask_slot);
  if (rc != RAG_OK) return rc;
  if (!out_rows_dev) return fail(RAG_EINVAL, "out_rows_dev is NULL");
  if (x->device != s->device) return fail(RAG_EINVAL, "exchange and store live on different devices");
  if (!rag_store_fused_ok(s, x, B, k, flags))
    return fail(RAG_EINVAL, "batch %d / k %d is not served by the fused exchange (use rag_store_query_dev + all-gather)", B, k);
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
  return search_device(s, c, c->d_buf, B, queries_dev, k, mask_slot, 1, row_base, so, false, x);
}

int rag_merge_keys_dev(int device, int G, int B, int k, const uint64_t* keys_dev, uint64_t* out_keys_dev,
                       int64_t* out_rows_dev, float* out_dists_dev, int32_t* out_counts_dev, void* stream) {
  if (G <= 0 || B <= 0 || k < 1 || k > RAG_MAX_K || !keys_dev) return fail(RAG_EINVAL, "bad merge arguments");
  CUDA_TRY(cudaSetDevice(device));
  MergeArgs ma{};
  ma.keys = keys_dev; ma.S = G; ma.B = B; ma.k = k; ma.row_base = 0;
  ma.out_keys = out_keys_dev; ma.out_rows = out_rows_dev; ma.out_dists = out_dists_dev; ma.out_counts = out_counts_dev;
  CUDA_TRY(launch_merge(ma, reinterpret_cast<cudaStream_t>(stream)));
  return RAG_OK;
}

uint64_t rag_key_pack(float dist, uint32_t row) { return make_key(dist, row); }
float rag_key_dist(uint64_t key) { return key_dist(key); }
uint32_t rag_key_row(uint64_t key) { return key_row(key); }

int rag_store_last_query_info(const rag_store* s, float* kernel_ms, int* regime, int* launches) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (kernel_ms) *kernel_ms = s->last_kernel_ms;
  if (regime) *regime = s->last_regime.load();
  if (launches) *launches = s->last_launches.load();
  return RAG_OK;
}

}  // extern "C"
