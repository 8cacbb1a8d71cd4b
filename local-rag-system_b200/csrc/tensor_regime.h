// Interface of the tensor-core (tcgen05) search regime used for large query batches.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace rag {
namespace tensor {

// batches larger than this go to the tensor regime when it supports the store (measured on
// B200, 10M x 768 bf16: stream 2.24 ms @ B=2, 2.68 ms @ B=4; tensor 2.5 ms for any B <= 128)
constexpr int kStreamMaxBatch = 3;

struct Plan;   // cached TMA descriptors etc. for one store

struct Problem {
  const void* vectors;      // [n_rows][row_elems] bf16, row-major
  const float* norms2;      // [n_rows] |x|^2 of the stored rows (l2 space)
  int64_t n_rows;
  int row_elems, dim, dtype, space;
  const uint32_t* live;
  const uint32_t* filter;
  int64_t filter_words;
  int dense;                // no filter and every row below n_rows is live
  const float* queries_raw; // [B][dim] fp32, unprepared
  int B, k;
  unsigned char* scratch;   // scratch_bytes() bytes
  int sm_count;
};

bool supported(int dtype, int row_elems, int k, int space);
size_t scratch_bytes(int dtype, int row_elems, int B, int k, int sm_count);
Plan* create_plan();
void destroy_plan(Plan* p);
void invalidate(Plan* p);     // corpus pointer / capacity changed
struct Result {
  const uint64_t* partial;  // [S][B][k] ascending candidate lists per query (inside `scratch`)
  int S;
  const float* q_f32;       // [B][row_elems] prepared queries (for the l2 refinement)
  uint64_t* merged;         // [B][k] scratch for the merged keys before refinement
};
// Runs prep + contraction + fused select.
cudaError_t launch(Plan* plan, const Problem& p, cudaStream_t st, Result* out, int* launches);

}  // namespace tensor
}  // namespace rag
