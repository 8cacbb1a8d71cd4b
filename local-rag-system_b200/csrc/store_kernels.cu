// K1 / K7: corpus-store maintenance kernels.
//   upsert  : what hnswlib does on add_items for the cosine space (normalise the
//             row) plus the store's element conversion; serves Collection.add /
//             upsert (api/app.py:221, scripts/build_index.py:92-96).
//   clear   : tombstones for Collection.delete (api/app.py:269,306,311).
//   prep    : the same normalise / round treatment for query vectors.
//   fetch   : stored rows back to fp32 (Collection.get(include=["embeddings"])).
// One warp per row, 16-byte vector accesses when the row pitch allows it.
#include "common.cuh"
#include "kernels.h"

namespace rag {
namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerCta = kThreads / 32;

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}

__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// scale = 1/|x| (cosine) or 1; the value written is what every later distance uses
__global__ void __launch_bounds__(kThreads) upsert_kernel(const UpsertArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (i >= a.n) return;
  const int64_t row = a.rows ? a.rows[i] : (a.row0 + i);
  const float* src = a.src + i * a.dim;
  float ss = 0.0f;
  for (int e = lane; e < a.dim; e += 32) { float x = src[e]; ss = fmaf(x, x, ss); }
  ss = warp_sum(ss);
  const float scale = a.normalise ? (ss > 0.0f ? rsqrtf(ss) : 0.0f) : 1.0f;
  float stored_ss = 0.0f;
  if (a.dtype == 1) {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.vectors) + row * a.row_elems;
    for (int e = lane; e < a.row_elems; e += 32) {
      float x = (e < a.dim) ? src[e] * scale : 0.0f;
      __nv_bfloat16 h = __float2bfloat16_rn(x);
      float y = __bfloat162float(h);
      stored_ss = fmaf(y, y, stored_ss);
      dst[e] = h;
    }
  } else {
    float* dst = reinterpret_cast<float*>(a.vectors) + row * a.row_elems;
    for (int e = lane; e < a.row_elems; e += 32) {
      float x = (e < a.dim) ? src[e] * scale : 0.0f;
      stored_ss = fmaf(x, x, stored_ss);
      dst[e] = x;
      if (a.shadow) {                  // keep the split-precision shadow in step with the row
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        a.shadow[row * 2 * a.row_elems + e] = h;
        a.shadow[row * 2 * a.row_elems + a.row_elems + e] = __float2bfloat16_rn(x - __bfloat162float(h));
      }
    }
  }
  stored_ss = warp_sum(stored_ss);
  if (lane == 0 && a.max_norm2 != nullptr) {    // non-negative floats order like their bit patterns
    const unsigned int bits = __float_as_uint(stored_ss);
    if (bits > *reinterpret_cast<volatile unsigned int*>(a.max_norm2))
      atomicMax(reinterpret_cast<unsigned int*>(a.max_norm2), bits);
    if (bits < *reinterpret_cast<volatile unsigned int*>(a.max_norm2 + 1))
      atomicMin(reinterpret_cast<unsigned int*>(a.max_norm2 + 1), bits);
  }
  if (lane == 0) {
    a.norms2[row] = stored_ss;
    atomicOr(a.live + (row >> 5), 1u << (row & 31));
  }
}

// x = hi + lo + O(2^-18 |x|): two bf16 planes that let the tensor cores contract fp32 rows
__global__ void __launch_bounds__(kThreads) split_rows_kernel(const float* vectors, int row_elems, int64_t row0,
                                                             int64_t n, __nv_bfloat16* shadow) {
  const int64_t total = n * row_elems;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / row_elems;
    const int e = static_cast<int>(i - r * row_elems);
    const float x = vectors[(row0 + r) * row_elems + e];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    __nv_bfloat16* dst = shadow + (row0 + r) * 2 * row_elems;
    dst[e] = h;
    dst[row_elems + e] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

__global__ void clear_live_kernel(uint32_t* live, const int64_t* rows, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = rows[i];
  atomicAnd(live + (row >> 5), ~(1u << (row & 31)));
}

__global__ void __launch_bounds__(kThreads) prep_queries_kernel(const PrepArgs a) {
  const int lane = threadIdx.x & 31;
  {   // merge-state initialisation for the scan kernel that follows on the stream
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
    if (a.init_keys) for (int64_t i = tid; i < a.init_keys_n; i += nth) a.init_keys[i] = kEmptyKey;
    if (a.init_zero) for (int64_t i = tid; i < a.init_zero_n; i += nth) a.init_zero[i] = 0;
  }
  const int b = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (b >= a.B) return;
  const float* src = a.src + static_cast<size_t>(b) * a.dim;
  float ss = 0.0f;
  for (int e = lane; e < a.dim; e += 32) { float x = src[e]; ss = fmaf(x, x, ss); }
  ss = warp_sum(ss);
  const float scale = a.normalise ? (ss > 0.0f ? rsqrtf(ss) : 0.0f) : 1.0f;
  float n2 = 0.0f;
  for (int e = lane; e < a.row_elems; e += 32) {
    float x = (e < a.dim) ? src[e] * scale : 0.0f;
    if (a.round_bf16) x = round_bf16(x);
    n2 = fmaf(x, x, n2);
    a.q_f32[static_cast<size_t>(b) * a.row_elems + e] = x;
    if (a.q_bf16) {
      if (a.split) {
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        __nv_bfloat16* dst = a.q_bf16 + static_cast<size_t>(b) * 2 * a.row_elems;
        dst[e] = h;
        dst[a.row_elems + e] = __float2bfloat16_rn(x - __bfloat162float(h));
      } else {
        a.q_bf16[static_cast<size_t>(b) * a.row_elems + e] = __float2bfloat16_rn(x);
      }
    }
  }
  n2 = warp_sum(n2);
  if (lane == 0 && a.q_norm2) a.q_norm2[b] = n2;
}

__global__ void __launch_bounds__(kThreads) fetch_kernel(const void* vectors, int dtype, int dim, int row_elems,
                                                         const int64_t* rows, int64_t n, float* out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (i >= n) return;
  const int64_t row = rows[i];
  for (int e = lane; e < dim; e += 32) {
    float x = (dtype == 1)
                  ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(vectors)[row * row_elems + e])
                  : reinterpret_cast<const float*>(vectors)[row * row_elems + e];
    out[i * dim + e] = x;
  }
}

}  // namespace

cudaError_t launch_upsert(const UpsertArgs& a, cudaStream_t st) {
  if (a.n <= 0) return cudaSuccess;
  const int64_t ctas = (a.n + kWarpsPerCta - 1) / kWarpsPerCta;
  upsert_kernel<<<static_cast<unsigned>(ctas), kThreads, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_split_rows(const float* vectors, int row_elems, int64_t row0, int64_t n, __nv_bfloat16* shadow,
                              cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t total = n * row_elems;
  int64_t ctas = (total + kThreads - 1) / kThreads;
  if (ctas > 148 * 16) ctas = 148 * 16;
  split_rows_kernel<<<static_cast<unsigned>(ctas), kThreads, 0, st>>>(vectors, row_elems, row0, n, shadow);
  return cudaGetLastError();
}

cudaError_t launch_clear_live(uint32_t* live, const int64_t* rows_dev, int64_t n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  clear_live_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(live, rows_dev, n);
  return cudaGetLastError();
}

cudaError_t launch_prep_queries(const PrepArgs& a, cudaStream_t st) {
  if (a.B <= 0) return cudaSuccess;
  prep_queries_kernel<<<(a.B + kWarpsPerCta - 1) / kWarpsPerCta, kThreads, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_fetch(const void* vectors, int dtype, int dim, int row_elems, const int64_t* rows_dev,
                         int64_t n, float* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  fetch_kernel<<<static_cast<unsigned>((n + kWarpsPerCta - 1) / kWarpsPerCta), kThreads, 0, st>>>(
      vectors, dtype, dim, row_elems, rows_dev, n, out);
  return cudaGetLastError();
}

}  // namespace rag
