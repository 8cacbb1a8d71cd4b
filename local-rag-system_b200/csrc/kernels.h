// Internal launch interface between the C-ABI host code (api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

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
  int k;
  int l2;                   // 1: d = sum((q-x)^2), 0: d = 1 - q.x
  int grid_x;
  // fused cross-CTA merge: every CTA publishes its sorted top-k, the LAST CTA of a query
  // group to finish (atomic ticket) merges all of them and emits the final result.
  uint64_t* partial;        // [grid_x][B][k] keys (ascending per list)
  unsigned int* done;       // [ceil(B / QB)] tickets, zero on entry; the last CTA re-zeroes its ticket
  int merge_keys_cap;       // keys of dynamic shared memory usable by the final in-smem sort (power of two)
  uint32_t row_base;        // added to emitted rows (multi-GPU: this shard's first global row)
  uint64_t* out_keys;       // [B][k] or nullptr
  int64_t* out_rows;        // [B][k] or nullptr
  float* out_dists;         // [B][k] or nullptr
  int32_t* out_counts;      // [B]    or nullptr
};
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
  uint32_t row_base;        // added to the row field of emitted keys
  uint64_t* out_keys;       // [B][k] or nullptr
  int64_t* out_rows;        // [B][k] or nullptr
  float* out_dists;         // [B][k] or nullptr
  int32_t* out_counts;      // [B]    or nullptr
};
cudaError_t launch_merge(const MergeArgs& a, cudaStream_t st);

// exact l2 re-scoring + re-sort of the B x k winners (tensor regime)
struct RefineArgs {
  const uint64_t* keys;     // [B][k] merged winners (rows local to the store)
  const void* vectors;
  const float* queries;     // [B][row_elems] prepared fp32
  int dtype, row_elems, B, k;
  uint32_t row_base;
  uint64_t* out_keys;
  int64_t* out_rows;
  float* out_dists;
  int32_t* out_counts;
};
cudaError_t launch_refine_l2(const RefineArgs& a, cudaStream_t st);

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
  uint32_t* live;
};
cudaError_t launch_upsert(const UpsertArgs& a, cudaStream_t st);

// K7: clear live bits
cudaError_t launch_clear_live(uint32_t* live, const int64_t* rows_dev, int64_t n, cudaStream_t st);

// query preparation: normalise (cosine), round (bf16 stores), zero-pad
struct PrepArgs {
  const float* src;         // [B][dim]
  int B, dim, row_elems;
  int normalise, round_bf16;
  float* q_f32;             // [B][row_elems]
  __nv_bfloat16* q_bf16;    // [Bpad][row_elems] or nullptr (rows >= B zero-filled by caller)
  float* q_norm2;           // [B] or nullptr
  // optional initialisation of the scan kernel's merge state (done here to save launches)
  uint64_t* init_keys;      // filled with kEmptyKey (init_keys_n entries) or nullptr
  int64_t init_keys_n;
  int* init_zero;           // zero-filled (init_zero_n ints) or nullptr
  int64_t init_zero_n;
};
cudaError_t launch_prep_queries(const PrepArgs& a, cudaStream_t st);

// fetch rows back as fp32
cudaError_t launch_fetch(const void* vectors, int dtype, int dim, int row_elems,
                         const int64_t* rows_dev, int64_t n, float* out, cudaStream_t st);

}  // namespace rag
