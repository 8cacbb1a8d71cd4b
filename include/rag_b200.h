/*
 * rag_b200.h -- C ABI of the B200-native exact dense-retrieval engine.
 *
 * This is the drop-in boundary for the vector-search hot path of
 * akak0487521/Local-RAG-System.  The reference has no FFI of its own: its hot
 * path is the `chromadb` Python API (chromadb==0.5.3, requirements.txt:6).
 * Each entry point below names the reference call it serves; the Python
 * binding a maintainer adds is in INTEGRATION.md (ctypes, as shipped in
 * local-rag-system_b200/_native.py).
 *
 * Conventions
 *   - every function returns 0 on success, a negative RAG_E* code on failure;
 *     rag_last_error() returns a thread-local message for the last failure.
 *   - no exceptions, no torch types, no C++ types cross this boundary.
 *   - the caller owns every host buffer; the engine owns all device memory.
 *   - rows are dense int64 indices local to one store (= one GPU shard);
 *     string ids, metadata and documents stay on the host side of the ABI.
 *   - a "key" is the 64-bit sortable pair  (ordered(fp32 distance) << 32 | row)
 *     used for candidate exchange between shards; smaller key = better hit,
 *     RAG_EMPTY_KEY marks an unused slot.
 *   - functions taking `stream` (a cudaStream_t passed as void*) are
 *     asynchronous on that stream; `*_dev` pointers are device pointers on the
 *     store's device.  All other functions take HOST pointers and are
 *     synchronous.
 *   - thread safety: any number of concurrent queries; upsert/delete/set_mask
 *     are exclusive (internal reader/writer lock), matching how the reference
 *     calls Chroma from FastAPI worker threads + BackgroundTasks
 *     (api/routes/kb.py:102-103,115,149).
 */
#ifndef RAG_B200_H
#define RAG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAG_B200_ABI_VERSION 3

#if defined(__GNUC__)
#define RAG_API __attribute__((visibility("default")))
#else
#define RAG_API
#endif

/* element type the corpus is stored in */
#define RAG_DTYPE_F32 0
#define RAG_DTYPE_BF16 1

/* distance space == Chroma collection metadata "hnsw:space"
 *   l2     d = sum((a-b)^2)                (reference default, api/app.py:91)
 *   cosine d = 1 - a.b/(|a||b|)            (rows L2-normalised on upsert)
 *   ip     d = 1 - a.b                                                       */
#define RAG_SPACE_L2 0
#define RAG_SPACE_COSINE 1
#define RAG_SPACE_IP 2

#define RAG_OK 0
#define RAG_EINVAL (-1)   /* bad argument                                   */
#define RAG_ECUDA (-2)    /* CUDA runtime / driver error (message has it)   */
#define RAG_ENOMEM (-3)   /* device or host allocation failed               */
#define RAG_ENODEV (-4)   /* no usable sm_100 device                        */

#define RAG_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define RAG_MAX_K 1024
#define RAG_MAX_MASK_SLOTS 16

typedef struct rag_store rag_store;
typedef struct rag_exchange rag_exchange;

RAG_API const char* rag_last_error(void);
RAG_API int rag_abi_version(void);

/* number of visible CUDA devices (0 if none / no driver); never fails */
RAG_API int rag_device_count(void);

/* -- lifetime ------------------------------------------------------------
 * chromadb.PersistentClient(path).get_or_create_collection(name, metadata)
 *   api/app.py:89-91, scripts/build_index.py:15-17                          */
RAG_API int rag_store_create(int dim, int dtype, int space, int device,
                     int64_t capacity_hint, rag_store** out);
/* same with flags.  By default a bf16 store also keeps the un-rounded fp32 rows (normalised for
 * cosine): searches rank the bf16 rows (the HBM-bound scan reads only those), keep k + 6..16
 * candidates and re-rank them exactly against the fp32 plane with the un-rounded query, so a bf16
 * store returns the hits and the fp32 distances of the reference's fp32 index (recall@k >= 0.999;
 * ranking on bf16 alone gives ~0.993 on 1M unit-norm rows).  RAG_STORE_NO_RERANK drops the plane
 * (2 bytes per element instead of 6) and with it that guarantee.                              */
#define RAG_STORE_NO_RERANK 1
RAG_API int rag_store_create_ex(int dim, int dtype, int space, int device,
                        int64_t capacity_hint, int flags, rag_store** out);
RAG_API int rag_store_destroy(rag_store* s);
/* make room for at least `rows` rows without further reallocation */
RAG_API int rag_store_reserve(rag_store* s, int64_t rows);

/* -- writes ----------------------------------------------------------------
 * Collection.add / Collection.upsert
 *   api/app.py:221, scripts/build_index.py:92-96, scripts/bulk_import.py:66-70,
 *   scripts/ingest_docs_to_chroma.py:31
 * vectors: n x dim fp32, row-major.  rows[i] >= 0 overwrites that row in
 * place (upsert of an existing id); rows == NULL or rows[i] == -1 appends
 * (a free row left by a delete is reused first).  out_rows (may be NULL)
 * receives the row each vector landed in.  Cosine stores: the row is
 * L2-normalised by the upsert kernel.                                        */
RAG_API int rag_store_upsert(rag_store* s, int64_t n, const float* vectors,
                     const int64_t* rows, int64_t* out_rows);
/* Calls of <= 64 rows -- the reference's pattern: one col.add per document (api/app.py:209-225), 1-5
 * chunks per col.upsert (scripts/build_index.py:89-96) -- are parked in pinned host memory and reach the
 * device as one copy + one launch when the next read arrives or 256 rows have piled up; counts, rows and
 * liveness reflect them at once.  rag_store_flush() forces them down and waits.                      */
RAG_API int rag_store_flush(rag_store* s);
/* same, vectors already on the store's device (bulk load / synthetic data) */
RAG_API int rag_store_upsert_dev(rag_store* s, int64_t n, const float* vectors_dev,
                         const int64_t* rows, int64_t* out_rows);

/* Collection.delete(ids=...) / delete(where=...) after the host resolved ids
 * to rows -- api/app.py:269, 306, 311.  Deleting a dead row is a no-op.      */
RAG_API int rag_store_delete(rag_store* s, int64_t n, const int64_t* rows);

/* -- state -------------------------------------------------------------------
 * Collection.count() -- api/routes/system.py:33.  O(1).                      */
RAG_API int64_t rag_store_count(const rag_store* s);      /* live rows               */
RAG_API int64_t rag_store_rows(const rag_store* s);       /* high-water mark         */
RAG_API int64_t rag_store_capacity(const rag_store* s);
RAG_API int rag_store_dim(const rag_store* s);
RAG_API int rag_store_dtype(const rag_store* s);
RAG_API int rag_store_space(const rag_store* s);
RAG_API int rag_store_device(const rag_store* s);
RAG_API int rag_store_has_rerank(const rag_store* s);   /* 1: bf16 store with the fp32 re-ranking plane */
/* 1 if row is live, 0 if dead / out of range */
RAG_API int rag_store_is_live(const rag_store* s, int64_t row);
/* number of engine kernels launched by this store since creation */
RAG_API int64_t rag_store_kernel_launches(const rag_store* s);

/* read vectors back as fp32 (the stored values: normalised / bf16-rounded) --
 * Collection.get(include=["embeddings"])                                     */
RAG_API int rag_store_fetch(rag_store* s, int64_t n, const int64_t* rows, float* out);
/* the un-rounded fp32 rows where the store keeps them (RAG_STORE_NO_RERANK not set), else as above */
RAG_API int rag_store_fetch_exact(rag_store* s, int64_t n, const int64_t* rows, float* out);

/* -- `where` filters -----------------------------------------------------------
 * Collection.query(where=...) -- api/app.py:540-548.  The host compiles the
 * predicate to a bitmap (bit r of word r/64, LSB first, 1 = row r passes) and
 * parks it in one of RAG_MAX_MASK_SLOTS slots; queries name the slot.
 * nbits may be smaller than the row count: missing rows do not pass.         */
RAG_API int rag_store_set_mask(rag_store* s, int slot, const uint64_t* bits, int64_t nbits);
RAG_API int rag_store_clear_mask(rag_store* s, int slot);
/* keep a parked mask valid across writes: bit rows[i] := pass[i] (0 / 1) for the n rows a write touched
 * (rows distinct; rows past the mask's end extend it, bits in between stay 0).  The reference interleaves
 * col.add with filtered /search (api/app.py:209-225, 540-548): re-uploading N/8 bytes per query after
 * every write is what this avoids.                                                                    */
RAG_API int rag_store_patch_mask(rag_store* s, int slot, int64_t n, const int64_t* rows,
                         const unsigned char* pass);

/* -- search --------------------------------------------------------------------
 * Collection.query(query_embeddings, n_results, where)
 *   api/app.py:544-549, scripts/query_local.py:29-34
 * queries: B x dim fp32.  mask_slot = -1 for no filter.  Output arrays are
 * B x k; out_counts[b] = number of valid hits (min(k, live & passing rows)),
 * hits ascending by (distance, row); unused tail = row -1, distance +inf.
 * `flags`: 0 = choose the kernel regime automatically,
 *          RAG_QUERY_FORCE_STREAM / RAG_QUERY_FORCE_TENSOR pin it (tests).   */
#define RAG_QUERY_AUTO 0
#define RAG_QUERY_FORCE_STREAM 1
#define RAG_QUERY_FORCE_TENSOR 2
RAG_API int rag_store_query(rag_store* s, int B, const float* queries, int k, int mask_slot,
                    int flags, int64_t* out_rows, float* out_dists, int32_t* out_counts);

/* Asynchronous shard-local search for the multi-GPU path: queries_dev and
 * out_keys_dev (B x k keys, ascending, RAG_EMPTY_KEY padded) are device
 * pointers; nothing is copied to the host.  `row_base` (this shard's first
 * global row) is added to the row field of every emitted key, so keys from
 * different shards merge by plain 64-bit comparison in exactly the order a
 * single store would produce.  Global rows must stay below 2^32.
 * out_rows_dev / out_dists_dev / out_counts_dev (B x k, B x k, B; each may be
 * NULL) additionally receive the decoded result -- a single-shard caller needs
 * no merge step.  At least one of out_keys_dev / out_rows_dev must be given.   */
RAG_API int rag_store_query_dev(rag_store* s, int B, const float* queries_dev, int k, int mask_slot,
                        int flags, uint32_t row_base, uint64_t* out_keys_dev,
                        int64_t* out_rows_dev, float* out_dists_dev, int32_t* out_counts_dev,
                        void* stream);

/* Host-buffer queries IN FLIGHT: what a server with concurrent requests does.  submit() stages the queries in
 * pinned memory, copies them down and launches the search on the store's pipeline stream, and returns a ticket
 * at once; wait() blocks until that query's result is on the host.  Up to 4 tickets may be outstanding; the
 * copy + launch of request i+1 then overlaps the scan of request i, and consecutive scans overlap by programmatic
 * dependent launch.  With an exchange (x != NULL; every rank submits the same sequence) this is the multi-GPU
 * fused search of rag_store_query_fused; with x == NULL the single-store search of rag_store_query.       */
RAG_API int rag_store_query_submit(rag_store* s, rag_exchange* x, int B, const float* queries, int k,
                           int mask_slot, int flags, uint32_t row_base, int* ticket);
RAG_API int rag_store_query_wait(rag_store* s, int ticket, int64_t* out_rows, float* out_dists,
                         int32_t* out_counts);

/* Cross-shard merge (after the NCCL all-gather of per-shard candidates):
 * keys_dev is G x B x k (each list ascending).  Writes B x k global rows /
 * distances and B counts on the device; any output pointer may be NULL.      */
RAG_API int rag_merge_keys_dev(int device, int G, int B, int k, const uint64_t* keys_dev,
                       uint64_t* out_keys_dev, int64_t* out_rows_dev, float* out_dists_dev,
                       int32_t* out_counts_dev, void* stream);

/* -- fused cross-shard exchange (multi-GPU, small batches) ------------------------
 * One process per GPU.  Instead of  scan -> NCCL all-gather -> merge kernel, the scan
 * kernel's last CTA stores the shard's B x k keys straight into every rank's exchange
 * buffer over NVLink (peer-mapped memory), raises a flag, waits for the other ranks'
 * flags and merges: ONE launch per query batch per GPU, no collective call on the data
 * path.  Serves the same reference call as rag_store_query (api/app.py:544-549).
 *   rag_exchange_create   allocates this rank's buffer (slot_keys >= B * k of any call)
 *   rag_exchange_handle   64-byte CUDA IPC handle of it; the caller all-gathers the
 *                         handles of all ranks (rank order) by any means
 *   rag_exchange_connect  maps every peer's buffer
 * All ranks must then issue the same sequence of rag_store_query_fused_dev calls
 * (same B, k), each on one stream per exchange; results (GLOBAL rows) land on every
 * rank.  A rank whose peers never show up gives up after 20 s and sets the status
 * word (rag_exchange_status) instead of hanging the GPU.
 * rag_store_fused_ok() tells whether a (B, k, flags) batch is served by this path
 * (stream regime, k <= 128, B * k <= slot_keys); otherwise use rag_store_query_dev +
 * all-gather + rag_merge_keys_dev.                                                 */
#define RAG_EXCHANGE_HANDLE_BYTES 64
RAG_API int rag_exchange_create(int device, int rank, int world, int64_t slot_keys, rag_exchange** out);
RAG_API int rag_exchange_handle(rag_exchange* x, void* out_handle);
RAG_API int rag_exchange_connect(rag_exchange* x, const void* handles);
RAG_API int rag_exchange_status(rag_exchange* x, int* timed_out);
RAG_API int rag_exchange_destroy(rag_exchange* x);
RAG_API int rag_store_fused_ok(const rag_store* s, const rag_exchange* x, int B, int k, int flags);
RAG_API int rag_store_query_fused_dev(rag_store* s, rag_exchange* x, int B, const float* queries_dev, int k,
                              int mask_slot, int flags, uint32_t row_base, int64_t* out_rows_dev,
                              float* out_dists_dev, int32_t* out_counts_dev, void* stream);
/* the same with HOST buffers, synchronous: pinned staging, H2D of the queries, the one fused launch,
 * one D2H of rows | dists | counts, all on the exchange's own stream -- the whole multi-GPU query is one
 * C call per rank (what Collection.query costs on N GPUs).  Fails with RAG_ECUDA if a peer timed out. */
RAG_API int rag_store_query_fused(rag_store* s, rag_exchange* x, int B, const float* queries, int k,
                          int mask_slot, int flags, uint32_t row_base, int64_t* out_rows,
                          float* out_dists, int32_t* out_counts);

/* -- one collection over several devices of ONE process ----------------------------------------
 * The reference builds one in-process collection and queries it from FastAPI worker threads
 * (api/app.py:87-91, api/routes/kb.py:173-206); this is the multi-GPU form of exactly that object
 * (Collection metadata "b200:devices": "0-7").  Chunks of 1024 rows are dealt round-robin to one store
 * per device, global rows stay dense, and every call below means what its rag_store_* namesake means,
 * with GLOBAL rows.  Small batches on distinct, peer-connected devices are answered by one fused
 * scan + exchange + merge launch per device, issued concurrently by resident worker threads; other
 * batches (tensor regime, shards sharing a device, no peer access) by per-shard searches whose keys
 * are gathered on the first device and merged there.  Either way the answer is bit-identical to one
 * store holding all rows.  `devices` may repeat a device (logical shards; used by the tests).        */
typedef struct rag_sharded rag_sharded;
RAG_API int rag_sharded_create(int dim, int dtype, int space, int n_devices, const int* devices,
                       int64_t capacity_hint, int flags, rag_sharded** out);
RAG_API int rag_sharded_destroy(rag_sharded* s);
RAG_API int rag_sharded_shards(const rag_sharded* s);
RAG_API rag_store* rag_sharded_shard(rag_sharded* s, int g);       /* borrowed; for introspection */
RAG_API int rag_sharded_fused(const rag_sharded* s);               /* 1: the one-launch path is available */
RAG_API int64_t rag_sharded_count(const rag_sharded* s);
RAG_API int64_t rag_sharded_rows(const rag_sharded* s);
RAG_API int rag_sharded_is_live(const rag_sharded* s, int64_t row);
RAG_API int rag_sharded_reserve(rag_sharded* s, int64_t rows);
RAG_API int rag_sharded_flush(rag_sharded* s);
RAG_API int rag_sharded_upsert(rag_sharded* s, int64_t n, const float* vectors, const int64_t* rows,
                       int64_t* out_rows);
RAG_API int rag_sharded_delete(rag_sharded* s, int64_t n, const int64_t* rows);
RAG_API int rag_sharded_fetch(rag_sharded* s, int64_t n, const int64_t* rows, float* out, int exact);
RAG_API int rag_sharded_set_mask(rag_sharded* s, int slot, const uint64_t* bits, int64_t nbits);
RAG_API int rag_sharded_patch_mask(rag_sharded* s, int slot, int64_t n, const int64_t* rows,
                           const unsigned char* pass);
RAG_API int rag_sharded_clear_mask(rag_sharded* s, int slot);
RAG_API int rag_sharded_query(rag_sharded* s, int B, const float* queries, int k, int mask_slot, int flags,
                      int64_t* out_rows, float* out_dists, int32_t* out_counts);
/* path: 1 = fused one-launch-per-device, 2 = per-shard search + gather + merge kernel */
RAG_API int rag_sharded_last_query_info(const rag_sharded* s, float* kernel_ms, int* regime, int* launches,
                                int* path);

/* helpers to (de)compose keys on the host */
RAG_API uint64_t rag_key_pack(float dist, uint32_t row);
RAG_API float rag_key_dist(uint64_t key);
RAG_API uint32_t rag_key_row(uint64_t key);

/* -- introspection for bench.py / profiles -------------------------------------
 * Device-side duration (ms, CUDA events on the engine's stream) of the scan /
 * contraction kernel of the most recent rag_store_query on this thread's
 * store, and which regime it ran (1 = stream, 2 = tensor).                   */
RAG_API int rag_store_last_query_info(const rag_store* s, float* kernel_ms, int* regime, int* launches);
/* debugging aid: with RAG_B200_TENSOR_STATS=1 in the environment the tensor-regime epilogue counts, on the
 * CURRENT device, [0] tiles drained per warp, [1] / [2] tiles passing the first / second reject test,
 * [3] candidate scores examined, [4] list insertions, [5] quantile-list updates; out8 receives 8 counters */
RAG_API int rag_debug_tensor_stats(uint64_t* out8, int reset);
/* debugging aid, pure host logic (no device needed): the shared-memory ring the tensor-regime kernel runs with when
 * `stages_available` stages fit, a corpus tile takes `stages_per_tile` of them and `accumulators` (2 or 4) buffers sit
 * in tensor memory -- how many stages it uses and whether one warp issues every tile (tests/test_ring_protocol.py
 * holds the answer to tools/ring_protocol_model.py) */
RAG_API int rag_debug_ring_plan(int stages_available, int stages_per_tile, int accumulators, int* stages_used,
                                int* one_issuer);
/* device time (ms, CUDA events on the admin stream) of the upsert kernel of the last rag_store_upsert_dev */
RAG_API float rag_store_last_upsert_ms(const rag_store* s);

/* -- fp32 stores in the tensor regime (large query batches; replaces the same collection.query call,
 * api/app.py:544-549, on a store created with dtype f32) -------------------------------------------
 * The tensor cores contract a bf16 SHADOW of the fp32 rows; the survivors are re-ranked exactly from the
 * fp32 rows and a guard certifies, per query, that no row outside the survivors can belong to the top k
 * (queries it cannot certify are re-run on the exact stream kernel inside the same call).  Two shadows:
 *   RAG_F32_SHADOW_HI    bf16(x) only: half the bytes and a third of the tensor work; rounding error ~1e-3 |q||x|,
 *                        covered by keeping 64-128 survivors.  Any dim % 8 == 0 up to 768.  Default.
 *   RAG_F32_SHADOW_HILO  [hi | lo] split precision, error ~1e-6 |q||x|; dim % 16 == 0 up to 384.  A store moves
 *                        here by itself when the hi-only guard keeps failing (> 1/8 of the queries).
 * rag_store_set_f32_shadow pins a kind (RAG_F32_SHADOW_AUTO = back to the policy); _info reports the kind in
 * force and how many tensor-regime queries were served / had to be re-run exactly since the store was created. */
#define RAG_F32_SHADOW_AUTO 0
#define RAG_F32_SHADOW_HI 1
#define RAG_F32_SHADOW_HILO 2
RAG_API int rag_store_set_f32_shadow(rag_store* s, int kind);
RAG_API int rag_store_f32_tensor_info(const rag_store* s, int* shadow_kind, int64_t* queries, int64_t* reruns);

#ifdef __cplusplus
}
#endif
#endif /* RAG_B200_H */
