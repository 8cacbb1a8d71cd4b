// Internal (not exported) view of one device store, shared by api.cu (the C ABI of a single store) and
// sharded.cu (the single-process multi-device store built from several of them).
#pragma once
#include <pthread.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/rag_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace rag {

int fail(int code, const char* fmt, ...);      // sets the thread-local rag_last_error() text, returns code

#define CUDA_TRY(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      (void)cudaGetLastError();                                                                   \
      return ::rag::fail(_e == cudaErrorMemoryAllocation ? RAG_ENOMEM : RAG_ECUDA, "%s: %s (%s:%d)", \
                         #expr, cudaGetErrorString(_e), __FILE__, __LINE__);                      \
    }                                                                                             \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// NVTX range around a host-side phase (header-only NVTX3: a no-op unless a tool is attached).  Ranges:
// rag:upsert, rag:flush_writes, rag:delete, rag:mask, rag:search:stream, rag:search:stream+exchange,
// rag:search:tensor, rag:merge, rag:rerank, rag:sharded_query
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// per-caller scratch: one stream + pinned staging + device scratch
struct QueryCtx {
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  unsigned char* h_pin = nullptr;
  size_t h_bytes = 0;
  unsigned char* d_buf = nullptr;
  size_t d_bytes = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;     // timed launches (synchronous API)
  cudaEvent_t ev_done = nullptr;                 // marker a writer records on this context's stream (asynchronous readers)
  bool launched = false;                         // ... since the last write waited for it
  uint64_t seen_write = 0;                       // last write sequence this context's stream has been ordered behind
  uint32_t signal_seq = 0;                       // last value handed to a kernel as its completion signal
  unsigned int* d_tickets = nullptr;             // kMaxTickets zeroed counters, self-resetting (scan kernel)
  static constexpr int kMaxTickets = 4096;
  // fp32 tensor regime: how many queries of this context's last such batch the guard sent to the exact re-run
  // (copied back behind the batch, read lazily: statistics for rag_store_f32_tensor_info and the shadow policy)
  int* h_redo = nullptr;                         // pinned [1]
  int redo_batch = 0;                            // queries of the batch h_redo belongs to (0 = nothing outstanding)
  int redo_kind = 0;                             // shadow kind that batch ran on

  int ensure_tickets();
  int ensure_host(size_t bytes);
  int ensure_dev(size_t bytes);
  int ensure_events();
  void destroy();
};

// small writes parked on the host until the next read (or until the buffer is full): the reference
// adds one document per call (api/app.py:209-225) and upserts 1-5 chunks per call
// (scripts/build_index.py:89-96); each would otherwise cost a host->device copy and a launch.
struct PendingWrites {
  static constexpr int64_t kMaxRows = 256;       // rows parked at most
  static constexpr int64_t kSmallCall = 64;      // calls with more rows than this go straight to the device
  unsigned char* h = nullptr;                    // pinned: [kMaxRows][dim] fp32 | [kMaxRows] int64 rows
  int64_t n = 0;
  std::unordered_map<int64_t, int64_t> slot_of;  // destination row -> slot (a second write to a row replaces the first)
  cudaEvent_t ev_h2d = nullptr;                  // the previous flush has left the pinned block
  bool in_flight = false;
};

// Host-buffer queries in flight (rag_store_query_submit / _wait): a server keeps a few requests in the air,
// so the copy + launch of request i+1 overlaps the scan of request i and consecutive scans overlap by
// programmatic dependent launch -- what the device-resident back-to-back loop measures, through host buffers.
struct PipeSlot {
  unsigned char* h = nullptr;      // pinned: queries | rows | dists | counts | flag
  unsigned char* d = nullptr;      // device mirror
  size_t bytes = 0;
  cudaEvent_t ev = nullptr;        // behind the slot's last device-to-host copy (when the flag path is not used)
  uint32_t seq = 0;
  uint32_t qseq = 0;               // arrival flag value of the slot's queries (flag word: last 4 bytes of the 256-byte tail of `d`)
  bool busy = false, armed = false;
  int B = 0, k = 0;
  size_t in_b = 0, rows_b = 0, dist_b = 0, cnt_b = 0;
};
struct Pipeline {
  static constexpr int kSlots = 4;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // carries the queries (and their arrival flags) of in-flight requests
  QueryCtx* ctx = nullptr;         // registered with the store like any asynchronous reader (scratch, write ordering)
  PipeSlot slot[kSlots];
  uint32_t next = 0;
  std::mutex mu;
};

struct SearchOut {
  uint64_t* keys = nullptr;
  int64_t* rows = nullptr;
  float* dists = nullptr;
  int32_t* counts = nullptr;
  // optional arrival flag of queries copied on another stream (see ScanArgs::query_flag)
  const uint32_t* query_flag = nullptr;
  uint32_t query_seq = 0;
  // optional completion flag in mapped pinned host memory (rows / dists / counts then point there too): if the
  // launch can raise it (fused stream regime, one query group) search_device sets *armed and the caller polls
  // the flag with wait_host_flag() instead of copying the result back and synchronising the stream
  uint32_t* done_flag = nullptr;
  uint32_t done_seq = 0;
  bool* armed = nullptr;
};

}  // namespace rag

// peer-mapped buffers for the fused scan + all-gather + merge launch (multi-GPU)
struct rag_exchange {
  int device = 0, rank = 0, world = 1;
  int64_t slot_keys = 0;
  size_t bytes = 0;
  unsigned char* d_local = nullptr;
  std::vector<unsigned char*> peers;     // [world] base of every rank's buffer as mapped into this process
  unsigned char** d_peers = nullptr;     // the same table on the device
  uint32_t epoch = 0;
  bool connected = false;
  bool ipc = true;                       // peers were opened with cudaIpcOpenMemHandle (one process per GPU)
  rag::QueryCtx host;                    // stream + pinned staging of the host-buffer call (rag_store_query_fused)
};

struct rag_store {
  int dim = 0, dtype = 0, space = 0, device = 0;
  int row_elems = 0;        // dim padded so that a row is a whole number of 16-byte chunks
  size_t row_bytes = 0;
  int exact_elems = 0;      // row pitch of the fp32 re-ranking plane (dim padded to 4), 0 = no plane
  int sm_count = 0;
  int64_t capacity = 0;     // rows allocated (multiple of 128)
  int64_t rows = 0;         // high-water mark
  int64_t live = 0;
  void* d_vectors = nullptr;
  float* d_exact = nullptr;               // bf16 stores: un-rounded fp32 rows (normalised for cosine) for the exact re-ranking
  float* d_norms2 = nullptr;
  float* d_max_norm2 = nullptr;           // [4]: [0] largest, [1] smallest |stored row|^2 ever written; [2] largest |x - bf16(x)|^2
                                          // over the rows written while a shadow existed (reset when one is built)
  uint32_t* d_live = nullptr;
  // fp32 stores, tensor regime: bf16 shadow of the rows -- [capacity][hi(row_elems)] (kShadowHi: bf16 filter + exact
  // re-ranking + guard) or [capacity][hi | lo] (kShadowHiLo: split precision) -- built on the first large-batch
  // query, kept in step by upsert, dropped (and rebuilt lazily) when the store grows or the kind changes.
  // A store starts with tensor::default_shadow_kind(); when the guard of the hi-only filter sends more than 1/8 of
  // the queries to the exact re-run (rows packed closer than bf16 can tell apart) it moves to the hi/lo split for good.
  __nv_bfloat16* d_shadow = nullptr;
  int shadow_kind = 0;                    // kind of d_shadow / of the shadow to build (write lock or shadow_mu)
  std::atomic<int> want_shadow_kind{0};   // != shadow_kind: switch before the next search (flush_if_pending)
  std::atomic<bool> shadow_pinned{false}; // kind fixed by rag_store_set_f32_shadow / RAG_B200_F32_SHADOW: no policy
  std::atomic<int64_t> f32_tensor_queries{0}, f32_tensor_reruns{0};      // lifetime totals (all kinds)
  std::atomic<int64_t> hi_window_queries{0}, hi_window_reruns{0};        // current policy window (hi-only batches)
  std::atomic<int> last_batch_reruns{1};  // re-runs of the most recent harvested batch (sizes the next re-run launch)
  std::mutex shadow_mu;
  uint32_t* d_masks[RAG_MAX_MASK_SLOTS] = {};     // each capacity / 32 words, zero beyond mask_words
  int64_t mask_words[RAG_MAX_MASK_SLOTS] = {};
  bool mask_set[RAG_MAX_MASK_SLOTS] = {};
  std::vector<uint32_t> h_live;
  std::vector<int64_t> free_rows;
  // rows are placed by an outer layer (the single-process multi-device store): explicit rows may lie at or
  // beyond the high-water mark and deleted rows are not remembered here (the outer layer re-uses them)
  bool external_rows = false;
  pthread_rwlock_t lock;
  // scratch pool for the synchronous API
  std::mutex pool_mu;
  std::condition_variable pool_cv;
  std::vector<rag::QueryCtx*> pool_free;
  int pool_created = 0;
  static constexpr int kMaxPool = 8;
  // scratch for the asynchronous API, one per caller stream
  std::mutex dev_mu;
  std::unordered_map<void*, rag::QueryCtx*> dev_ctx;
  rag::QueryCtx admin;      // upsert / delete / fetch / masks (used under the write lock)
  cudaEvent_t ev_write = nullptr;         // recorded behind the last device-side write
  std::atomic<uint64_t> write_seq{0};
  rag::PendingWrites pending;
  rag::Pipeline* pipe = nullptr;
  std::atomic<int64_t> pending_n{0};
  std::atomic<int64_t> launches{0};
  std::atomic<int> last_regime{0};
  std::atomic<int> last_launches{0};
  float last_kernel_ms = 0.0f;
  std::atomic<rag::QueryCtx*> timing_ctx{nullptr};   // events of the last flag-signalled query, read lazily
  float last_upsert_ms = 0.0f;            // device time of the last rag_store_upsert_dev kernel (CUDA events)
};

namespace rag {

// read lock held by the caller.  Orders the context's stream behind the last write, runs the regime's
// kernels asynchronously on c->stream.  `scratch` must hold search_scratch_bytes().
size_t search_scratch_bytes(const rag_store* s, int B, int k, int grid_x);
int search_device(rag_store* s, QueryCtx* c, unsigned char* scratch, int B, const float* d_queries_raw, int k,
                  int mask_slot, int regime, RowMap rows_map, const SearchOut& out, bool timed,
                  rag_exchange* xchg, uint32_t xchg_epoch, bool forced_tensor);
int choose_regime(const rag_store* s, int B, int k, int flags);
int batch_limit(const rag_store* s, int k);
int check_query_args(const rag_store* s, int B, const void* q, int k, int mask_slot);
// can a (B, k) search of this store write its result straight to host memory and raise a flag?
bool direct_host_ok(const rag_store* s, int B, int k, int regime);
// spin on a completion flag in mapped pinned memory; falls back to the stream's status to surface errors
int wait_host_flag(const volatile uint32_t* flag, uint32_t seq, cudaStream_t st);
// pending small writes -> device, pending change of the fp32 shadow kind (takes the write lock itself when there
// is something to do)
int flush_if_pending(rag_store* s);
int dev_ctx_for(rag_store* s, void* stream, QueryCtx** out);
// list length the scan keeps for a request of k hits (k + slack with the exact re-ranking)
int scan_k(const rag_store* s, int k);

}  // namespace rag
