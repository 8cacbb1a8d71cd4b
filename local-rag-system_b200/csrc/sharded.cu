// K6, single-process form (SURVEY.md 8e "process model"): ONE collection over G device stores inside
// the process that owns the collection -- what the reference's FastAPI worker needs
// (api/app.py:87-91 builds one in-process collection and queries it from worker threads).
//
// Placement: chunks of 2^kStripeShift rows are dealt round-robin to the shards (RowMap, common.cuh), so
// global rows stay dense (the host's id / document / metadata arrays are indexed by them) while every
// shard grows at the same rate.  Keys carry GLOBAL rows: a sharded search returns exactly what one store
// over the union returns, ties included.
//
// Search, small batches (stream regime) on G distinct devices with peer access -- the latency-bound case:
//   the caller's thread and G-1 resident worker threads (one per device, woken through one atomic) each
//   copy the query batch to their device and launch the shard's fused kernel: scan + all-gather of the
//   B x k keys over NVLink peer stores + cross-shard merge in ONE launch per device (scan_stream.cu).
//   The result lands on every device; device 0 copies it back.  No NCCL, no host-side merge.
// Everything else (tensor regime, shards sharing a device, no peer access): every shard searches
//   asynchronously on its own stream and emits keys with global rows, the G lists are copied to shard 0's
//   device (peer copies ordered by events) and merged there by the merge kernel.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <thread>

#include "store_internal.h"

using namespace rag;

namespace {
constexpr uint32_t kStripeShift = 10;                 // 1024-row chunks = 32 bitmap words
constexpr int64_t kStripeRows = 1ll << kStripeShift;
constexpr int kMaxShards = kXchgMaxWorld;

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
}  // namespace

struct rag_sharded {
  int G = 0;
  int dim = 0, dtype = 0, space = 0;
  std::vector<rag_store*> shard;
  std::vector<int> device;
  std::vector<rag_exchange*> xchg;          // fused mode only
  std::vector<QueryCtx> ctx;                // one search context (stream + scratch) per shard
  std::vector<cudaEvent_t> ev_keys;         // gather mode: shard g's keys are ready
  bool fused_ok = false;                    // distinct devices, all pairs peer-mapped
  // global bookkeeping
  pthread_rwlock_t lock;
  int64_t rows = 0;                         // global high-water mark
  int64_t live = 0;
  std::vector<uint32_t> h_live;
  std::vector<int64_t> free_rows;
  // queries
  std::mutex query_mu;                      // one sharded search at a time (the exchange is a lock-step protocol)
  unsigned char* h_q = nullptr;             // pinned, portable: query batch read by every device
  size_t h_q_bytes = 0;
  unsigned char* d_gather = nullptr;        // gather mode: [G][B][k] keys on shard 0's device
  size_t d_gather_bytes = 0;
  // worker threads (fused mode)
  std::vector<std::thread> workers;
  std::atomic<uint64_t> job_seq{0};
  std::atomic<int> job_done{0};
  std::atomic<bool> stop{false};
  std::mutex job_mu;
  std::condition_variable job_cv;
  std::atomic<int> sleepers{0};
  struct Job { int B = 0, k = 0, mask_slot = -1; uint32_t epoch = 0; } job;
  std::vector<int> job_rc;
  std::vector<std::string> job_err;
  // introspection
  float last_kernel_ms = 0.0f;
  bool timing_pending = false;
  int last_regime = 0, last_launches = 0, last_path = 0;
};

namespace {

inline int shard_of(const rag_sharded* s, int64_t grow) { return (int)((grow >> kStripeShift) % s->G); }
inline int64_t local_of(const rag_sharded* s, int64_t grow) {
  return (((grow >> kStripeShift) / s->G) << kStripeShift) | (grow & (kStripeRows - 1));
}
inline RowMap map_of(const rag_sharded* s, int g) { return RowMap{(uint32_t)g << kStripeShift, kStripeShift, (uint32_t)s->G}; }
// local rows shard g holds when the global high-water mark is n
inline int64_t local_rows(const rag_sharded* s, int g, int64_t n) {
  const int64_t chunks = n >> kStripeShift, rem = n & (kStripeRows - 1);
  int64_t r = (chunks / s->G) * kStripeRows;
  const int64_t c = chunks % s->G;
  if (g < c) r += kStripeRows;
  else if (g == c) r += rem;
  return r;
}
inline bool g_is_live(const rag_sharded* s, int64_t r) {
  return r >= 0 && r < s->rows && ((s->h_live[(size_t)(r >> 5)] >> (r & 31)) & 1u);
}

// H2D of the batch + the shard's fused launch (any thread; cudaSetDevice done by the caller)
int run_fused_shard(rag_sharded* s, int g, const SearchOut& so) {
  rag_store* st = s->shard[g];
  QueryCtx* c = &s->ctx[g];
  const rag_sharded::Job& j = s->job;
  int rc = flush_if_pending(st);
  if (rc != RAG_OK) return rc;
  RdLock rl(&st->lock);
  const size_t in_b = align_up((size_t)j.B * s->dim * sizeof(float), 256);
  const size_t out_b = align_up((size_t)j.B * j.k * sizeof(int64_t), 256) + align_up((size_t)j.B * j.k * sizeof(float), 256) +
                       align_up((size_t)j.B * sizeof(int32_t), 256);
  const int grid_x = scan_stream_grid_x(st->sm_count, st->rows);
  rc = c->ensure_dev(in_b + out_b + search_scratch_bytes(st, j.B, j.k, grid_x));
  if (rc != RAG_OK) return rc;
  unsigned char* d = c->d_buf;
  CUDA_TRY(cudaMemcpyAsync(d, s->h_q, (size_t)j.B * s->dim * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  SearchOut o = so;
  if (!o.rows) {              // shards other than 0 still need somewhere to put the (identical) result
    o.rows = reinterpret_cast<int64_t*>(d + in_b);
    o.dists = nullptr; o.counts = nullptr;
  }
  return search_device(st, c, d + in_b + out_b, j.B, reinterpret_cast<const float*>(d), j.k, j.mask_slot, 1, map_of(s, g), o,
                       g == 0, s->xchg[g], j.epoch, false);
}

void worker_main(rag_sharded* s, int g) {
  cudaSetDevice(s->device[g]);
  uint64_t seen = 0;
  for (;;) {
    // spin briefly (a query is ~0.3 ms; the next one usually follows at once), then sleep
    int spins = 0;
    while (s->job_seq.load(std::memory_order_acquire) == seen && !s->stop.load(std::memory_order_acquire)) {
      if (++spins < 20000) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        continue;
      }
      std::unique_lock<std::mutex> lk(s->job_mu);
      s->sleepers.fetch_add(1);
      s->job_cv.wait(lk, [&] { return s->job_seq.load(std::memory_order_acquire) != seen || s->stop.load(); });
      s->sleepers.fetch_sub(1);
    }
    if (s->stop.load(std::memory_order_acquire)) return;
    seen = s->job_seq.load(std::memory_order_acquire);
    const int rc = run_fused_shard(s, g, SearchOut{});
    s->job_rc[(size_t)g] = rc;
    if (rc != RAG_OK) s->job_err[(size_t)g] = rag_last_error();
    s->job_done.fetch_add(1, std::memory_order_release);
  }
}

// map every shard's exchange buffer into every other shard's device (same process: plain peer access)
int connect_local(rag_sharded* s) {
  const int G = s->G;
  for (int a = 0; a < G; ++a)
    for (int b = 0; b < G; ++b) {
      if (a == b) continue;
      if (s->device[a] == s->device[b]) return RAG_EINVAL;      // shards sharing a device never run the fused kernel
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, s->device[a], s->device[b]) != cudaSuccess || !can) { (void)cudaGetLastError(); return RAG_EINVAL; }
    }
  for (int a = 0; a < G; ++a) {
    CUDA_TRY(cudaSetDevice(s->device[a]));
    for (int b = 0; b < G; ++b) {
      if (a == b) continue;
      cudaError_t e = cudaDeviceEnablePeerAccess(s->device[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return RAG_EINVAL; }
      (void)cudaGetLastError();
    }
  }
  s->xchg.assign((size_t)G, nullptr);
  for (int g = 0; g < G; ++g) {
    int rc = rag_exchange_create(s->device[g], g, G, 8192, &s->xchg[(size_t)g]);
    if (rc != RAG_OK) return rc;
  }
  for (int g = 0; g < G; ++g) {
    rag_exchange* x = s->xchg[(size_t)g];
    for (int p = 0; p < G; ++p) x->peers[(size_t)p] = s->xchg[(size_t)p]->d_local;
    CUDA_TRY(cudaSetDevice(s->device[g]));
    CUDA_TRY(cudaMemcpy(x->d_peers, x->peers.data(), (size_t)G * sizeof(unsigned char*), cudaMemcpyHostToDevice));
    x->connected = true;
    x->ipc = false;
  }
  return RAG_OK;
}

}  // namespace

extern "C" {

int rag_sharded_create(int dim, int dtype, int space, int n_devices, const int* devices, int64_t capacity_hint, int flags,
                       rag_sharded** out) {
  if (!out) return fail(RAG_EINVAL, "out is NULL");
  *out = nullptr;
  if (n_devices < 1 || n_devices > kMaxShards || !devices) return fail(RAG_EINVAL, "a sharded store takes 1..%d devices", kMaxShards);
  rag_sharded* s = new (std::nothrow) rag_sharded();
  if (!s) return fail(RAG_ENOMEM, "out of host memory");
  s->G = n_devices; s->dim = dim; s->dtype = dtype; s->space = space;
  pthread_rwlock_init(&s->lock, nullptr);
  s->device.assign(devices, devices + n_devices);
  s->ctx.resize((size_t)n_devices);
  s->ev_keys.assign((size_t)n_devices, nullptr);
  s->job_rc.assign((size_t)n_devices, 0);
  s->job_err.resize((size_t)n_devices);
  // whole chunks per shard, so that a shard never has to grow in the middle of a bulk load
  const int64_t per = (std::max<int64_t>(capacity_hint, 0) + n_devices * kStripeRows - 1) / (n_devices * kStripeRows) * kStripeRows;
  for (int g = 0; g < n_devices; ++g) {
    rag_store* st = nullptr;
    int rc = rag_store_create_ex(dim, dtype, space, devices[g], per, flags, &st);
    if (rc != RAG_OK) { rag_sharded_destroy(s); return rc; }
    st->external_rows = true;
    s->shard.push_back(st);
    cudaError_t e = cudaStreamCreateWithFlags(&s->ctx[(size_t)g].stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_keys[(size_t)g], cudaEventDisableTiming);
    if (e != cudaSuccess) { (void)cudaGetLastError(); rag_sharded_destroy(s); return fail(RAG_ECUDA, "stream/event creation failed: %s", cudaGetErrorString(e)); }
    s->ctx[(size_t)g].own_stream = true;
  }
  s->fused_ok = false;
  if (n_devices > 1 && !(getenv("RAG_B200_FUSED_EXCHANGE") && atoi(getenv("RAG_B200_FUSED_EXCHANGE")) == 0)) {
    if (connect_local(s) == RAG_OK) {
      s->fused_ok = true;
      for (int g = 1; g < n_devices; ++g) s->workers.emplace_back(worker_main, s, g);
    } else {
      for (rag_exchange* x : s->xchg) if (x) rag_exchange_destroy(x);
      s->xchg.clear();
    }
  }
  *out = s;
  return RAG_OK;
}

int rag_sharded_destroy(rag_sharded* s) {
  if (!s) return RAG_OK;
  s->stop.store(true, std::memory_order_release);
  { std::lock_guard<std::mutex> lk(s->job_mu); s->job_cv.notify_all(); }
  for (std::thread& t : s->workers) if (t.joinable()) t.join();
  for (size_t g = 0; g < s->shard.size(); ++g) {
    cudaSetDevice(s->device[g]);
    cudaDeviceSynchronize();
  }
  for (rag_exchange* x : s->xchg) if (x) rag_exchange_destroy(x);
  for (size_t g = 0; g < s->ctx.size(); ++g) {
    if (g < s->device.size()) cudaSetDevice(s->device[g]);
    s->ctx[g].destroy();
    if (s->ev_keys[g]) cudaEventDestroy(s->ev_keys[g]);
  }
  if (!s->device.empty()) cudaSetDevice(s->device[0]);
  if (s->d_gather) cudaFree(s->d_gather);
  if (s->h_q) cudaFreeHost(s->h_q);
  for (rag_store* st : s->shard) rag_store_destroy(st);
  pthread_rwlock_destroy(&s->lock);
  delete s;
  return RAG_OK;
}

int rag_sharded_shards(const rag_sharded* s) { return s ? s->G : 0; }
rag_store* rag_sharded_shard(rag_sharded* s, int g) { return (s && g >= 0 && g < s->G) ? s->shard[(size_t)g] : nullptr; }
int64_t rag_sharded_count(const rag_sharded* s) { return s ? s->live : 0; }
int64_t rag_sharded_rows(const rag_sharded* s) { return s ? s->rows : 0; }
int rag_sharded_fused(const rag_sharded* s) { return (s && s->fused_ok) ? 1 : 0; }

int rag_sharded_is_live(const rag_sharded* s, int64_t row) {
  if (!s) return 0;
  RdLock g(const_cast<pthread_rwlock_t*>(&s->lock));
  return g_is_live(s, row) ? 1 : 0;
}

int rag_sharded_reserve(rag_sharded* s, int64_t rows) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  const int64_t per = (std::max<int64_t>(rows, 0) + s->G * kStripeRows - 1) / (s->G * kStripeRows) * kStripeRows;
  for (rag_store* st : s->shard) {
    int rc = rag_store_reserve(st, per);
    if (rc != RAG_OK) return rc;
  }
  return RAG_OK;
}

int rag_sharded_flush(rag_sharded* s) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  for (rag_store* st : s->shard) {
    int rc = rag_store_flush(st);
    if (rc != RAG_OK) return rc;
  }
  return RAG_OK;
}

int rag_sharded_upsert(rag_sharded* s, int64_t n, const float* vectors, const int64_t* rows, int64_t* out_rows) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (n < 0) return fail(RAG_EINVAL, "n < 0");
  if (n == 0) return RAG_OK;
  if (!vectors) return fail(RAG_EINVAL, "vectors is NULL");
  WrLock wl(&s->lock);
  // global rows: explicit, a free row, or the next one
  std::vector<int64_t> dst((size_t)n);
  std::vector<int64_t> taken;
  int64_t hwm = s->rows;
  for (int64_t i = 0; i < n; ++i) {
    int64_t r = rows ? rows[i] : -1;
    if (r >= 0) {
      if (r >= hwm) {
        for (int64_t t : taken) s->free_rows.push_back(t);
        return fail(RAG_EINVAL, "upsert row %lld is beyond the store's %lld rows", (long long)r, (long long)hwm);
      }
    } else {
      while (!s->free_rows.empty()) {
        const int64_t c = s->free_rows.back();
        s->free_rows.pop_back();
        if (c < s->rows && !g_is_live(s, c) && std::find(taken.begin(), taken.end(), c) == taken.end()) { r = c; break; }
      }
      if (r >= 0) taken.push_back(r);
      else r = hwm++;
    }
    dst[(size_t)i] = r;
  }
  if (hwm > 0xFFFFFFF0ll) return fail(RAG_EINVAL, "a sharded store holds at most 2^32-16 rows");
  // split per shard, keeping the caller's order inside a shard
  std::vector<std::vector<int64_t>> idx((size_t)s->G);
  for (int64_t i = 0; i < n; ++i) idx[(size_t)shard_of(s, dst[(size_t)i])].push_back(i);
  std::vector<float> stage;
  std::vector<int64_t> lrows;
  for (int g = 0; g < s->G; ++g) {
    const std::vector<int64_t>& ix = idx[(size_t)g];
    if (ix.empty()) continue;
    lrows.resize(ix.size());
    for (size_t j = 0; j < ix.size(); ++j) lrows[j] = local_of(s, dst[(size_t)ix[j]]);
    // runs of consecutive input rows need no staging copy
    const bool run = ix.back() - ix.front() + 1 == (int64_t)ix.size();
    const float* src = vectors + (size_t)ix.front() * s->dim;
    if (!run) {
      stage.resize(ix.size() * (size_t)s->dim);
      for (size_t j = 0; j < ix.size(); ++j)
        memcpy(stage.data() + j * (size_t)s->dim, vectors + (size_t)ix[j] * s->dim, (size_t)s->dim * sizeof(float));
      src = stage.data();
    }
    int rc = rag_store_upsert(s->shard[(size_t)g], (int64_t)ix.size(), src, lrows.data(), nullptr);
    if (rc != RAG_OK) {
      for (int64_t t : taken) s->free_rows.push_back(t);
      return rc;      // shards already written keep their rows; the global bitmap below was not touched
    }
  }
  s->h_live.resize((size_t)((std::max(hwm, s->rows) + 31) / 32 + 1), 0u);
  for (int64_t r : dst) {
    uint32_t& w = s->h_live[(size_t)(r >> 5)];
    const uint32_t bit = 1u << (r & 31);
    if (!(w & bit)) { w |= bit; s->live++; }
  }
  s->rows = std::max(s->rows, hwm);
  if (out_rows) memcpy(out_rows, dst.data(), (size_t)n * sizeof(int64_t));
  return RAG_OK;
}

int rag_sharded_delete(rag_sharded* s, int64_t n, const int64_t* rows) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (n <= 0) return RAG_OK;
  if (!rows) return fail(RAG_EINVAL, "rows is NULL");
  WrLock wl(&s->lock);
  std::vector<std::vector<int64_t>> per((size_t)s->G);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t r = rows[i];
    if (!g_is_live(s, r)) continue;
    s->h_live[(size_t)(r >> 5)] &= ~(1u << (r & 31));
    s->live--;
    s->free_rows.push_back(r);
    per[(size_t)shard_of(s, r)].push_back(local_of(s, r));
  }
  for (int g = 0; g < s->G; ++g) {
    if (per[(size_t)g].empty()) continue;
    int rc = rag_store_delete(s->shard[(size_t)g], (int64_t)per[(size_t)g].size(), per[(size_t)g].data());
    if (rc != RAG_OK) return rc;
  }
  return RAG_OK;
}

int rag_sharded_fetch(rag_sharded* s, int64_t n, const int64_t* rows, float* out, int exact) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (n <= 0) return RAG_OK;
  if (!rows || !out) return fail(RAG_EINVAL, "rows/out is NULL");
  RdLock rl(&s->lock);
  std::vector<std::vector<int64_t>> idx((size_t)s->G), lrows((size_t)s->G);
  for (int64_t i = 0; i < n; ++i) {
    if (rows[i] < 0 || rows[i] >= s->rows) return fail(RAG_EINVAL, "fetch row %lld out of range", (long long)rows[i]);
    const int g = shard_of(s, rows[i]);
    idx[(size_t)g].push_back(i);
    lrows[(size_t)g].push_back(local_of(s, rows[i]));
  }
  std::vector<float> tmp;
  for (int g = 0; g < s->G; ++g) {
    if (idx[(size_t)g].empty()) continue;
    tmp.resize(idx[(size_t)g].size() * (size_t)s->dim);
    int rc = exact ? rag_store_fetch_exact(s->shard[(size_t)g], (int64_t)idx[(size_t)g].size(), lrows[(size_t)g].data(), tmp.data())
                   : rag_store_fetch(s->shard[(size_t)g], (int64_t)idx[(size_t)g].size(), lrows[(size_t)g].data(), tmp.data());
    if (rc != RAG_OK) return rc;
    for (size_t j = 0; j < idx[(size_t)g].size(); ++j)
      memcpy(out + (size_t)idx[(size_t)g][j] * s->dim, tmp.data() + j * (size_t)s->dim, (size_t)s->dim * sizeof(float));
  }
  return RAG_OK;
}

int rag_sharded_set_mask(rag_sharded* s, int slot, const uint64_t* bits, int64_t nbits) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (nbits < 0 || (nbits > 0 && !bits)) return fail(RAG_EINVAL, "bad mask arguments");
  WrLock wl(&s->lock);
  // a 1024-row chunk is 32 whole 32-bit words: the shard's bitmap is every G-th 128-byte run of the global one
  const uint32_t* w = reinterpret_cast<const uint32_t*>(bits);
  const int64_t gwords = (nbits + 31) / 32;
  constexpr int64_t kWordsPerChunk = kStripeRows / 32;
  std::vector<uint32_t> local;
  for (int g = 0; g < s->G; ++g) {
    const int64_t lbits = local_rows(s, g, nbits);
    local.assign((size_t)((lbits + 63) / 64 * 2), 0u);
    const int64_t lwords = (lbits + 31) / 32;
    for (int64_t lw0 = 0; lw0 < lwords; lw0 += kWordsPerChunk) {
      const int64_t gw0 = ((lw0 / kWordsPerChunk) * s->G + g) * kWordsPerChunk;
      const int64_t cnt = std::min<int64_t>(kWordsPerChunk, std::min(lwords - lw0, gwords - gw0));
      if (cnt > 0) memcpy(local.data() + lw0, w + gw0, (size_t)cnt * 4);
    }
    // the last global word may carry bits past nbits; rag_store_set_mask clears bits past lbits
    int rc = rag_store_set_mask(s->shard[(size_t)g], slot, reinterpret_cast<const uint64_t*>(local.data()), lbits);
    if (rc != RAG_OK) return rc;
  }
  return RAG_OK;
}

int rag_sharded_patch_mask(rag_sharded* s, int slot, int64_t n, const int64_t* rows, const unsigned char* pass) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (n <= 0) return RAG_OK;
  if (!rows || !pass) return fail(RAG_EINVAL, "rows/pass is NULL");
  WrLock wl(&s->lock);
  std::vector<std::vector<int64_t>> lr((size_t)s->G);
  std::vector<std::vector<unsigned char>> lp((size_t)s->G);
  for (int64_t i = 0; i < n; ++i) {
    if (rows[i] < 0) return fail(RAG_EINVAL, "mask row %lld out of range", (long long)rows[i]);
    const int g = shard_of(s, rows[i]);
    lr[(size_t)g].push_back(local_of(s, rows[i]));
    lp[(size_t)g].push_back(pass[i]);
  }
  for (int g = 0; g < s->G; ++g) {
    if (lr[(size_t)g].empty()) continue;
    int rc = rag_store_patch_mask(s->shard[(size_t)g], slot, (int64_t)lr[(size_t)g].size(), lr[(size_t)g].data(), lp[(size_t)g].data());
    if (rc != RAG_OK) return rc;
  }
  return RAG_OK;
}

int rag_sharded_clear_mask(rag_sharded* s, int slot) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  WrLock wl(&s->lock);
  for (rag_store* st : s->shard) {
    int rc = rag_store_clear_mask(st, slot);
    if (rc != RAG_OK) return rc;
  }
  return RAG_OK;
}

int rag_sharded_query(rag_sharded* s, int B, const float* queries, int k, int mask_slot, int flags,
                      int64_t* out_rows, float* out_dists, int32_t* out_counts) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (!out_rows || !out_dists || !out_counts) return fail(RAG_EINVAL, "output pointer is NULL");
  NvtxRange nvtx("rag:sharded_query");
  RdLock rl(&s->lock);
  rag_store* s0 = s->shard[0];
  int rc = check_query_args(s0, B, queries, k, mask_slot);
  if (rc != RAG_OK) return rc;
  if (s->live == 0) {
    const float inf = __builtin_inff();
    for (int64_t i = 0; i < (int64_t)B * k; ++i) { out_rows[i] = -1; out_dists[i] = inf; }
    for (int b = 0; b < B; ++b) out_counts[b] = 0;
    return RAG_OK;
  }
  std::lock_guard<std::mutex> ql(s->query_mu);
  const int G = s->G;
  // batches the scratch of one launch cannot take are cut up (as rag_store_query does)
  int lim = 4096;
  for (rag_store* st : s->shard) lim = std::min(lim, batch_limit(st, k));
  float total_ms = 0.0f;
  int total_launches = 0;
  for (int b0 = 0; b0 < B; b0 += lim) {
    const int Bc = std::min(lim, B - b0);
    const size_t in_bytes = (size_t)Bc * s->dim * sizeof(float);
    if (in_bytes > s->h_q_bytes) {
      for (int g = 0; g < G; ++g) { cudaSetDevice(s->device[(size_t)g]); cudaStreamSynchronize(s->ctx[(size_t)g].stream); }
      if (s->h_q) cudaFreeHost(s->h_q);
      s->h_q = nullptr; s->h_q_bytes = 0;
      const size_t want = align_up(std::max(in_bytes, (size_t)1 << 16), 4096);
      CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&s->h_q), want, cudaHostAllocPortable));
      s->h_q_bytes = want;
    }
    memcpy(s->h_q, queries + (size_t)b0 * s->dim, in_bytes);
    const int regime = choose_regime(s0, Bc, k, flags);
    if (regime < 0) return fail(RAG_EINVAL, "tensor regime does not support this store/query (dtype %d, dim %d, k %d)", s->dtype, s->dim, k);
    const size_t in_b = align_up(in_bytes, 256);
    const size_t rows_b = align_up((size_t)Bc * k * sizeof(int64_t), 256);
    const size_t dist_b = align_up((size_t)Bc * k * sizeof(float), 256);
    const size_t cnt_b = align_up((size_t)Bc * sizeof(int32_t), 256);
    QueryCtx* c0 = &s->ctx[0];
    CUDA_TRY(cudaSetDevice(s->device[0]));
    rc = c0->ensure_host(rows_b + dist_b + cnt_b + 256);      // + the completion flag
    if (rc != RAG_OK) return rc;
    bool armed = false;
    uint32_t* done_flag = reinterpret_cast<uint32_t*>(c0->h_pin + rows_b + dist_b + cnt_b);
    uint32_t done_seq = 0;

    bool fused = s->fused_ok && regime == 1;
    if (fused)
      for (int g = 0; g < G && fused; ++g) fused = rag_store_fused_ok(s->shard[(size_t)g], s->xchg[(size_t)g], Bc, k, flags) != 0;
    unsigned char* d_out = nullptr;          // rows | dists | counts on device 0
    if (fused) {
      // ---- one fused launch per device, issued concurrently by the resident worker threads ----
      s->last_path = 1;
      s->job.B = Bc; s->job.k = k; s->job.mask_slot = mask_slot;
      uint32_t epoch = 0;
      for (int g = 0; g < G; ++g) epoch = ++s->xchg[(size_t)g]->epoch;       // all exchanges advance in step
      s->job.epoch = epoch;
      s->job_done.store(0, std::memory_order_relaxed);
      s->job_seq.fetch_add(1, std::memory_order_release);
      if (s->sleepers.load(std::memory_order_acquire) > 0) { std::lock_guard<std::mutex> lk(s->job_mu); s->job_cv.notify_all(); }
      // shard 0 runs on this thread
      const int grid_x = scan_stream_grid_x(s0->sm_count, s0->rows);
      rc = c0->ensure_dev(in_b + rows_b + dist_b + cnt_b + search_scratch_bytes(s0, Bc, k, grid_x));
      SearchOut so{};
      if (rc == RAG_OK) {
        d_out = c0->d_buf + in_b;
        if (direct_host_ok(s0, Bc, k, 1)) {
          // device 0 writes the merged result straight into the pinned block and raises a flag there
          d_out = c0->h_pin;
          done_seq = ++c0->signal_seq ? c0->signal_seq : ++c0->signal_seq;
          so.done_flag = done_flag; so.done_seq = done_seq; so.armed = &armed;
        }
        so.rows = reinterpret_cast<int64_t*>(d_out);
        so.dists = reinterpret_cast<float*>(d_out + rows_b);
        so.counts = reinterpret_cast<int32_t*>(d_out + rows_b + dist_b);
        rc = run_fused_shard(s, 0, so);
      }
      while (s->job_done.load(std::memory_order_acquire) < G - 1) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      for (int g = 1; g < G && rc == RAG_OK; ++g)
        if (s->job_rc[(size_t)g] != RAG_OK) rc = fail(s->job_rc[(size_t)g], "shard %d: %s", g, s->job_err[(size_t)g].c_str());
      if (rc != RAG_OK) return rc;
      total_launches += G;
    } else {
      // ---- every shard searches on its own stream; keys are gathered on device 0 and merged there ----
      s->last_path = 2;
      const size_t keys_b = (size_t)Bc * k * sizeof(uint64_t);
      if ((size_t)G * keys_b > s->d_gather_bytes) {
        CUDA_TRY(cudaStreamSynchronize(c0->stream));
        if (s->d_gather) cudaFree(s->d_gather);
        s->d_gather = nullptr; s->d_gather_bytes = 0;
        const size_t want = align_up((size_t)G * keys_b, 1 << 16);
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_gather), want));
        s->d_gather_bytes = want;
      }
      for (int g = 0; g < G; ++g) {
        rag_store* st = s->shard[(size_t)g];
        QueryCtx* c = &s->ctx[(size_t)g];
        rc = flush_if_pending(st);
        if (rc != RAG_OK) return rc;
        CUDA_TRY(cudaSetDevice(s->device[(size_t)g]));
        RdLock sl(&st->lock);
        const int grid_x = scan_stream_grid_x(st->sm_count, st->rows);
        const size_t own_keys = align_up(keys_b, 256);
        rc = c->ensure_dev(in_b + own_keys + rows_b + dist_b + cnt_b + search_scratch_bytes(st, Bc, k, grid_x));
        if (rc != RAG_OK) return rc;
        unsigned char* d = c->d_buf;
        CUDA_TRY(cudaMemcpyAsync(d, s->h_q, in_bytes, cudaMemcpyHostToDevice, c->stream));
        uint64_t* d_keys = reinterpret_cast<uint64_t*>(d + in_b);
        if (st->live == 0) {
          CUDA_TRY(cudaMemsetAsync(d_keys, 0xFF, keys_b, c->stream));
        } else {
          SearchOut so{};
          so.keys = d_keys;
          rc = search_device(st, c, d + in_b + own_keys + rows_b + dist_b + cnt_b, Bc, reinterpret_cast<const float*>(d), k, mask_slot,
                             regime, map_of(s, g), so, g == 0, nullptr, 0u, flags == RAG_QUERY_FORCE_TENSOR);
          if (rc != RAG_OK) return rc;
          total_launches += st->last_launches.load();
        }
        // the copy runs on the SOURCE shard's stream (ordered behind its search); device 0's stream waits for it
        unsigned char* dst = s->d_gather + (size_t)g * keys_b;
        if (s->device[(size_t)g] == s->device[0]) CUDA_TRY(cudaMemcpyAsync(dst, d_keys, keys_b, cudaMemcpyDeviceToDevice, c->stream));
        else CUDA_TRY(cudaMemcpyPeerAsync(dst, s->device[0], d_keys, s->device[(size_t)g], keys_b, c->stream));
        CUDA_TRY(cudaEventRecord(s->ev_keys[(size_t)g], c->stream));
      }
      CUDA_TRY(cudaSetDevice(s->device[0]));
      for (int g = 0; g < G; ++g) CUDA_TRY(cudaStreamWaitEvent(c0->stream, s->ev_keys[(size_t)g], 0));
      d_out = c0->d_buf + in_b + align_up(keys_b, 256);
      MergeArgs ma{};
      ma.keys = reinterpret_cast<const uint64_t*>(s->d_gather); ma.S = G; ma.B = Bc; ma.k = k;
      ma.out_rows = reinterpret_cast<int64_t*>(d_out);
      ma.out_dists = reinterpret_cast<float*>(d_out + rows_b);
      ma.out_counts = reinterpret_cast<int32_t*>(d_out + rows_b + dist_b);
      CUDA_TRY(launch_merge(ma, c0->stream));
      total_launches += 1;
    }
    CUDA_TRY(cudaSetDevice(s->device[0]));
    if (armed) {
      rc = wait_host_flag(done_flag, done_seq, c0->stream);
      if (rc != RAG_OK) return rc;
    } else {
      CUDA_TRY(cudaMemcpyAsync(c0->h_pin, d_out, rows_b + dist_b + cnt_b, cudaMemcpyDeviceToHost, c0->stream));
      CUDA_TRY(cudaStreamSynchronize(c0->stream));
    }
    memcpy(out_rows + (size_t)b0 * k, c0->h_pin, (size_t)Bc * k * sizeof(int64_t));
    memcpy(out_dists + (size_t)b0 * k, c0->h_pin + rows_b, (size_t)Bc * k * sizeof(float));
    memcpy(out_counts + b0, c0->h_pin + rows_b + dist_b, (size_t)Bc * sizeof(int32_t));
    for (int b = 0; b < Bc; ++b)
      if (out_counts[b0 + b] < 0) return fail(RAG_ECUDA, "fused exchange: a shard did not deliver its candidates within 20 s");
    float ms = 0.0f;
    if (armed) s->timing_pending = true;              // read lazily by rag_sharded_last_query_info
    else if (c0->ev0 && c0->ev1 && cudaEventElapsedTime(&ms, c0->ev0, c0->ev1) == cudaSuccess) total_ms += ms;
    else (void)cudaGetLastError();
    s->last_regime = s0->last_regime.load();
  }
  s->last_kernel_ms = total_ms;
  s->last_launches = total_launches;
  return RAG_OK;
}

int rag_sharded_last_query_info(const rag_sharded* s, float* kernel_ms, int* regime, int* launches, int* path) {
  if (!s) return fail(RAG_EINVAL, "store is NULL");
  if (s->timing_pending) {
    rag_sharded* m = const_cast<rag_sharded*>(s);
    QueryCtx* c0 = &m->ctx[0];
    float ms = 0.0f;
    m->timing_pending = false;
    if (cudaSetDevice(s->device[0]) == cudaSuccess && c0->ev1 && cudaEventSynchronize(c0->ev1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, c0->ev0, c0->ev1) == cudaSuccess) m->last_kernel_ms = ms;
    else (void)cudaGetLastError();
  }
  if (kernel_ms) *kernel_ms = s->last_kernel_ms;
  if (regime) *regime = s->last_regime;
  if (launches) *launches = s->last_launches;
  if (path) *path = s->last_path;
  return RAG_OK;
}

}  // extern "C"
