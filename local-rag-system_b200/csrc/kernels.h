// Internal launch interface between the C-ABI host code (api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"

namespace rag {

constexpr int kScanThreads = 256;             // 8 warps per CTA
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kRowsPerBlock = 32;             // one live/filter bitmap word

// ---- K2+K4: HBM-streaming scan with fused top-k ------------------------------
struct ScanArgs {
  const void* vectors;      // [rows][row_elems] f32 or bf16, row-major, 16B-aligned rows
  int dtype;                // RAG_DTYPE_*
  int row_elems;            // padded dim
  int cpr;                  // 16-byte chunks per row
  int64_t n_rows;           // high-water mark
  const uint32_t* live;     // bitmap, ceil(n_rows/32) words valid, bit=1 live
  const uint32_t* filter;   // optional bitmap (nullptr = all pass)
  int64_t filter_words;     // words available in `filter`
  const float* queries;     // [B][row_elems] prepared fp32, or nullptr when queries_raw is given
  const float* queries_raw; // [B][dim] raw fp32: the kernel normalises / rounds them itself (saves a launch)
  int dim, normalise, round_bf16;
  int B;
  int k;                    // list length kept per query during the scan
  int k_out;                // hits emitted per query (<= k; < k only with the exact re-ranking below)
  int l2;                   // 1: d = sum((q-x)^2), 0: d = 1 - q.x
  int grid_x;
  // Exact fp32 re-ranking (bf16 stores with an fp32 plane; nullptr = off): the k best rows
  // by stored-precision distance are re-scored against the un-rounded fp32 rows with the un-rounded
  // query, re-sorted, and the best k_out emitted.  Needs queries_raw (the fused single-launch mode).
  const float* exact;       // [rows][exact_elems] fp32 (normalised for cosine stores)
  int exact_elems;          // row pitch of `exact` in floats (dim padded to 4)
  // fused cross-CTA merge: every CTA publishes its sorted top-k, the LAST CTA of a query
  // group to finish (atomic ticket) merges all of them and emits the final result.
  uint64_t* partial;        // [grid_x][B][k] keys (ascending per list)
  unsigned int* done;       // [ceil(B / QB)] tickets, zero on entry; the last CTA re-zeroes its ticket
  int merge_keys_cap;       // keys of dynamic shared memory usable by the final in-smem sort (power of two)
  RowMap rows_map;          // local -> global row of emitted hits (multi-GPU: this shard's placement)
  uint64_t* out_keys;       // [B][k] or nullptr
  int64_t* out_rows;        // [B][k] or nullptr
  float* out_dists;         // [B][k] or nullptr
  int32_t* out_counts;      // [B]    or nullptr
  // Fused cross-shard exchange over NVLink peer memory (multi-GPU, optional; nullptr = off).
  // The last CTA of a query group stores the shard's k best keys into EVERY rank's exchange
  // buffer (peer-mapped stores), raises its flag there, waits for all ranks' flags in its own
  // buffer, merges the `xchg_world` lists and emits the GLOBAL result: scan + all-gather +
  // merge are one launch, no NCCL call on the data path.
  unsigned char* const* xchg_peers;   // device array [xchg_world]: base of every rank's exchange buffer
  int xchg_rank, xchg_world;
  uint32_t xchg_epoch;      // >= 1, identical on every rank, +1 per call (parity picks the buffer half)
  int64_t xchg_slot_keys;   // keys per (parity, rank) slot; B * k must fit
  // Queries that arrive on ANOTHER stream (queries in flight, rag_store_query_submit): the host copies them to
  // `queries_raw` on a copy stream and then copies `query_seq` to *query_flag; the kernel -- launched with no stream
  // dependency on those copies, so that no operation sits between two searches and their programmatic overlap
  // survives -- waits for the flag before it reads the queries (it is normally there long before).  nullptr = off.
  const uint32_t* query_flag;
  uint32_t query_seq;
  // Completion signal for host-resident outputs (nullptr = off; only with one query group per launch):
  // out_rows / out_dists / out_counts may point into MAPPED PINNED host memory -- the last CTA then writes the
  // B x k result straight over PCIe and finally stores done_seq to *done_flag (release, system scope), which
  // the host polls.  Replaces a device-to-host copy + a stream synchronisation (~15 us of a 0.3 ms query).
  uint32_t* done_flag;
  uint32_t done_seq;
  // optional device-side query list: only the first *q_count queries are searched and query i is
  // queries_raw[q_index[i]] / results go to slot q_index[i] (nullptr = all B queries, identity)
  const int* q_count;
  const int* q_index;
};
// ---- layout of one rank's exchange buffer (identical on every rank) ----------------------
//   u32 flags[2][kXchgMaxGroups][kXchgMaxWorld]   flags[p][y][g] = last epoch of parity p for which rank g
//                                                 has delivered the lists of query group y
//   u32 status[64]                                status[0] != 0: a wait timed out (peer missing)
//   u64 keys[2][world][slot_keys]                 keys[p][g][b * k + j]
constexpr int kXchgMaxGroups = 64;
constexpr int kXchgMaxWorld = 32;
constexpr size_t kXchgFlagBytes = 2ull * kXchgMaxGroups * kXchgMaxWorld * sizeof(uint32_t);
constexpr size_t kXchgStatusOff = kXchgFlagBytes;
constexpr size_t kXchgKeysOff = kXchgFlagBytes + 64 * sizeof(uint32_t);
inline size_t xchg_buffer_bytes(int world, int64_t slot_keys) {
  return kXchgKeysOff + 2ull * world * static_cast<size_t>(slot_keys) * sizeof(uint64_t);
}
// picks QB (queries per pass) and the kernel instantiation; returns cudaError_t
cudaError_t launch_scan_stream(const ScanArgs& a, int sm_count, cudaStream_t st, int* launches);
// grid_x the launcher will use for this problem (so the caller can size `partial`)
int scan_stream_grid_x(int sm_count, int64_t n_rows);
// max queries per launch the stream kernel handles in one corpus pass
int scan_stream_max_qb(int dtype, int row_elems, int k);
// number of query groups (grid.y) the launcher will use; `done` needs that many tickets
int scan_stream_groups(int B, int dtype, int row_elems, int k);

// ---- K4b/K6: merge S sorted candidate lists per query ------------------------
struct MergeArgs {
  const uint64_t* keys;     // [S][B][k]
  int S, B, k;
  RowMap rows_map;          // local -> global row of emitted keys
  uint64_t* out_keys;       // [B][k] or nullptr
  int64_t* out_rows;        // [B][k] or nullptr
  float* out_dists;         // [B][k] or nullptr
  int32_t* out_counts;      // [B]    or nullptr
};
cudaError_t launch_merge(const MergeArgs& a, cudaStream_t st);

// exact re-scoring + re-sort of the B x k_in approximate winners (tensor regime); emits the best k
struct RefineArgs {
  const uint64_t* keys;     // [B][k_in] merged winners (rows local to the store)
  const void* vectors;
  const float* queries;     // [B][row_elems] prepared fp32 (same pitch as the rows re-scored)
  int dtype, row_elems, B, k;   // dtype / row_elems describe `vectors` (the fp32 plane of a bf16 store: dtype 0)
  int k_in;                 // candidates per query (>= k)
  int l2;                   // 1: sum((q-x)^2), 0: 1 - q.x
  // exactness guard of the split-precision regime (nullptr = off): the candidates were ranked by
  // approximate distances with |approx - exact| <= guard_eps.  If the list is full and its worst
  // approximate distance is not at least 2*guard_eps beyond the exact k-th distance, a row outside
  // the list could belong to the top k: the query index is appended to redo_list (count in
  // redo_count) and re-run on the exact fp32 stream kernel.
  float guard_rel;          // |approx - exact| <= guard_rel * |q| * max|x| (x 2 for l2) ...
  const float* q_norm2;     // [B]
  const float* x_max_norm2; // [1]
  // ... plus, when the rows were contracted as bf16(x) with bf16(q) only (hi-only shadow, nullptr otherwise), the
  // rounding that contraction ignores:  |q.x - qh.xh| <= |q - qh| |x| + |qh| |x - xh|   (Cauchy-Schwarz), with
  // |q - qh| known per query and |x - xh| bounded by its maximum over the store
  const float* q_lo_norm2;  // [B] |q - bf16(q)|^2
  const float* x_lo_max2;   // [1] max over the store of |x - bf16(x)|^2
  int* redo_count;
  int* redo_list;
  RowMap rows_map;
  uint64_t* out_keys;
  int64_t* out_rows;
  float* out_dists;
  int32_t* out_counts;
};
cudaError_t launch_refine(const RefineArgs& a, cudaStream_t st);
constexpr float kGuardRel = 1.2e-4f;      // fp32 accumulation + distance rounding, relative to |q| max|x|

// Hi-only FILTER mode (fp32 store contracted as bf16(q).bf16(x), k <= 16; tensor_regime.h, Result::filt).
// `keys` = the merged approximate top-k (k_in entries per query); A_k = its k-th distance.  With
// eps = bf16_contraction_eps() >= |approx - exact|: the approximate top-k rows all have exact <= A_k + eps, so the
// exact k-th best is <= A_k + eps, so every row of the exact top-k has approx <= A_k + 2 eps.  The contraction left,
// per (CTA, query), every row within 2 eps of a bound >= A_k (`extra`): all rows within 2 eps of A_k are re-scored
// from the fp32 rows, sorted, and the best k emitted -- exact, no a-posteriori guard.  A query whose buffers
// overflowed (more candidates within 2 eps than they hold) goes to redo_list instead.  Uses RefineArgs' fields plus:
struct RefineFilterArgs {
  RefineArgs r;             // keys / vectors / queries / k / k_in / l2 / guard_rel / norms / redo_* / outputs as above
  const uint64_t* extra;    // [S][B][cap]
  const int* extra_cnt;     // [S][B]
  int S, cap;
};
cudaError_t launch_refine_filter(const RefineFilterArgs& a, cudaStream_t st);

// ---- K1: normalise / convert on upsert ------------------------------------------
struct UpsertArgs {
  const float* src;         // [n][dim] fp32 (device)
  const int64_t* rows;      // [n] destination rows (device) or nullptr => row0 + i
  int64_t row0;
  int64_t n;
  int dim, row_elems, dtype;
  int normalise;            // cosine
  void* vectors;
  float* norms2;            // [capacity] sum of squares of the stored row
  float* max_norm2;         // [2] running maximum [0] and minimum [1] of norms2 over everything ever stored (error / rejection bounds)
  uint32_t* live;
  // optional un-rounded fp32 plane of a bf16 store (exact re-ranking): [capacity][exact_elems]; nullptr when absent
  float* exact;
  int exact_elems;
  // optional bf16 shadow of an fp32 store for the tensor regime (nullptr when absent), by shadow_kind:
  //   kShadowHiLo  row r = [hi(row_elems) | lo(row_elems)] with x = hi + lo + O(2^-17 |x|)   (split precision)
  //   kShadowHi    row r = hi(row_elems) = bf16(x) only                                      (bf16 filter + exact re-rank)
  __nv_bfloat16* shadow;
  int shadow_kind;
  float* lo_max2;           // [1] running maximum of |x - bf16(x)|^2 over the rows written while a shadow exists, or nullptr
};
constexpr int kShadowNone = 0, kShadowHi = 1, kShadowHiLo = 2;
// (re)build the shadow of rows [row0, row0 + n) from the stored fp32 rows; lo_max2 as in UpsertArgs
cudaError_t launch_split_rows(const float* vectors, int row_elems, int64_t row0, int64_t n,
                              __nv_bfloat16* shadow, int shadow_kind, float* lo_max2, cudaStream_t st);
cudaError_t launch_upsert(const UpsertArgs& a, cudaStream_t st);

// K7: clear live bits
cudaError_t launch_clear_live(uint32_t* live, const int64_t* rows_dev, int64_t n, cudaStream_t st);

// query preparation: normalise (cosine), round (bf16 stores), zero-pad
struct PrepArgs {
  const float* src;         // [B][dim]
  int B, dim, row_elems;
  int normalise, round_bf16;
  int split;                // 1: q_bf16 rows are [hi(row_elems) | lo(row_elems)] (fp32 stores, tensor regime)
  float* q_f32;             // [B][row_elems]
  float* q_exact;           // [B][exact_elems] normalised but NOT rounded (exact re-ranking of bf16 stores) or nullptr
  int exact_elems;
  __nv_bfloat16* q_bf16;    // [Bpad][row_elems * (split ? 2 : 1)] or nullptr (rows >= B zero-filled by caller)
  float* q_norm2;           // [B] or nullptr
  float* q_lo_norm2;        // [B] |x - bf16(x)|^2 of the prepared (un-rounded) query, or nullptr
  // optional initialisation of the scan kernel's merge state (done here to save launches)
  uint64_t* init_keys;      // filled with kEmptyKey (init_keys_n entries) or nullptr
  int64_t init_keys_n;
  int* init_zero;           // zero-filled (init_zero_n ints) or nullptr
  int64_t init_zero_n;
};
cudaError_t launch_prep_queries(const PrepArgs& a, cudaStream_t st);

// set / clear single bits of a `where` bitmap: bit rows[i] := pass[i] (rows are distinct within a call)
cudaError_t launch_patch_mask(uint32_t* mask, const int64_t* rows_dev, const unsigned char* pass_dev, int64_t n,
                              cudaStream_t st);

// fetch rows back as fp32
cudaError_t launch_fetch(const void* vectors, int dtype, int dim, int row_elems,
                         const int64_t* rows_dev, int64_t n, float* out, cudaStream_t st);

}  // namespace rag
