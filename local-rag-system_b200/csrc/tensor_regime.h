// Interface of the tensor-core (tcgen05) search regime used for large query batches.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace rag {
namespace tensor {

// batches larger than this go to the tensor regime when it supports the store.  Measured on B200 with
// tools/crossover.py (final round-1 kernels), 10M x 768 bf16: stream 2.07 / 2.24 / 3.09 ms at B = 1 / 2 / 3,
// tensor 2.36-2.47 ms for any B <= 16
constexpr int kStreamMaxBatch = 2;
// fp32 stores (1M x 384): the exact stream kernel 0.21 / 0.27 / 0.38 / 0.41 ms at B = 1 / 4 / 6 / 8; the bf16-shadow
// contraction + exact re-ranking 0.26-0.28 ms (hi-only filter) / 0.35 ms (hi/lo split) for any B <= 32
constexpr int kStreamMaxBatchF32 = 4;
// ... and ANY batch, down to a single query, once an fp32 store with the hi-only shadow (top-k <= 16: the filter path)
// holds this many bytes of rows: the shadow is half the bytes of the fp32 rows the stream kernel reads, which beats
// the fixed cost of the tensor regime's five launches (tools/crossover.py, B = 1: 1M x 384 0.198 vs 0.215 ms,
// 1M x 768 0.323 vs 0.423 ms; the reference's own 25-vector collection stays on the one-launch stream kernel)
constexpr size_t kF32TensorAlwaysBytes = (size_t)1 << 30;

struct Problem {
  const void* vectors;      // [n_rows][row_elems] bf16 (or fp32 when `shadow` is streamed instead), row-major
  const void* shadow;       // fp32 stores: bf16 shadow of the rows, [n_rows][hi | lo] (kShadowHiLo) or [n_rows][hi] (kShadowHi)
  int shadow_kind;          // kernels.h: kShadowNone (bf16 stores), kShadowHi, kShadowHiLo
  const float* norms2;      // [n_rows] |x|^2 of the stored rows (l2 space)
  const float* min_norm2;   // [1] lower bound of norms2 over everything the store ever held
  const float* max_norm2;   // [1] upper bound of the same   } error bound of the hi-only contraction
  const float* lo_max2;     // [1] max |x - bf16(x)|^2        } (fp32 stores; kernels.h, UpsertArgs::lo_max2)
  int64_t n_rows;
  int row_elems, dim, dtype, space;
  const uint32_t* live;
  const uint32_t* filter;
  int64_t filter_words;
  int dense;                // no filter and every row below n_rows is live
  int rerank;               // bf16 store with an fp32 plane: keep k + slack candidates, also emit un-rounded queries
  int exact_elems;          // row pitch of that plane (floats)
  const float* queries_raw; // [B][dim] fp32, unprepared
  int B, k;
  unsigned char* scratch;   // scratch_bytes() bytes
  int sm_count;
};

// shadow_kind: which bf16 shadow an fp32 store is contracted through (ignored for bf16 stores)
bool supported(int dtype, int row_elems, int k, int space, int rerank, int shadow_kind);
// the shadow a NEW fp32 store of this row length starts with (RAG_B200_F32_SHADOW=hi|hilo overrides):
// kShadowHi where the kernel can take it, else kShadowHiLo, else kShadowNone (no tensor regime for the store)
int default_shadow_kind(int row_elems);
// candidates the kernel keeps per query: k for bf16 stores (exact ranking of the stored values);
// more for fp32 stores, whose rows are ranked through a bf16 shadow and re-ranked exactly -- k + 6..16 behind the
// hi/lo split (error ~1e-6), 64..128 behind the hi-only filter (error ~1e-3) --
// (and for bf16 stores that keep an un-rounded fp32 plane: ranked in bf16, re-ranked against the plane)
int candidates_kept(int dtype, int k, int rerank, int shadow_kind);
size_t scratch_bytes(int dtype, int row_elems, int B, int k, int sm_count, int rerank, int shadow_kind);
struct Result {
  const uint64_t* partial;  // [S][B][k_kept] ascending candidate lists per query (inside `scratch`)
  int S;
  int k_kept;               // list length: k, or k + slack when ranking was approximate (fp32 stores)
  const float* q_norm2;     // [B] |prepared query|^2
  const float* q_lo_norm2;  // [B] |q - bf16(q)|^2 (kShadowHi only, else nullptr)
  const float* q_f32;       // [B][row_elems] prepared queries (for the l2 refinement)
  const float* q_exact;     // [B][exact_elems] normalised, un-rounded queries (Problem::rerank)
  uint64_t* merged;         // [B][k_kept] scratch for the merged keys before refinement
  // hi-only FILTER mode (fp32 store, kShadowHi, k <= 16): next to the approximate top-k lists every (CTA, query)
  // left the rows within 2 eps of its bound; launch_refine_filter re-scores those within 2 eps of the merged k-th best
  int filt;
  const uint64_t* extra;    // [S][B][extra_cap]
  const int* extra_cnt;     // [S][B], -1 = overflow
  int extra_cap;
};
// shared-memory ring the kernel runs with (host logic; tensor_regime.cu has the reasoning)
void plan_ring(int avail, int per_tile, int nbuf, int* stages, int* one_issuer);
// epilogue selection counters (RAG_B200_TENSOR_STATS=1), see tensor_regime.cu
int read_stats(unsigned long long* out8, int reset);
// Runs prep + contraction + fused select.
cudaError_t launch(const Problem& p, cudaStream_t st, Result* out, int* launches);

}  // namespace tensor
}  // namespace rag
