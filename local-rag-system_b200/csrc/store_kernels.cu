// K1 / K7: corpus-store maintenance kernels.
//   upsert  : what hnswlib does on add_items for the cosine space (normalise the
//             row) plus the store's element conversion; serves Collection.add /
//             upsert (api/app.py:221, scripts/build_index.py:92-96).
//   clear   : tombstones for Collection.delete (api/app.py:269,306,311).
//   prep    : the same normalise / round treatment for query vectors.
//   fetch   : stored rows back to fp32 (Collection.get(include=["embeddings"])).
// One warp per row; the upsert kernel is single-pass with 16-byte accesses whenever dim % 4 == 0.
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

// ---- K1: normalise / convert on upsert -------------------------------------------------------------
// Roofline: HBM.  Algorithmic bytes per row = dim * 4 read + row_bytes written (+ dim * 4 for the fp32
// re-ranking plane of a bf16 store, + row_bytes / 2 or row_bytes for the hi / hi-lo shadow of an fp32 store once it exists).
//
// Vector path (dim % 4 == 0, dim <= 2048): ONE pass, one warp per row.  Lane l holds the 8 consecutive
// elements 8 * (l + 32 j) .. + 7 of the row for j < NJ in registers (two 16-byte streaming loads each, a warp
// reads 1 KB contiguous per j), the sum of squares is reduced by shuffles, and the scaled / rounded row is
// written with 16-byte stores.  The scale and the stored values are what every later distance uses.
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  const uint4 u = ldg_stream(reinterpret_cast<const uint4*>(p));
  return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(lo))) |
         (static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(hi))) << 16);
}
__device__ __forceinline__ float bf16_lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }

__device__ __forceinline__ void track_norm_bounds(float* max_norm2, float stored_ss) {
  if (max_norm2 == nullptr) return;            // non-negative floats order like their bit patterns
  const unsigned int bits = __float_as_uint(stored_ss);
  if (bits > *reinterpret_cast<volatile unsigned int*>(max_norm2))
    atomicMax(reinterpret_cast<unsigned int*>(max_norm2), bits);
  if (bits < *reinterpret_cast<volatile unsigned int*>(max_norm2 + 1))
    atomicMin(reinterpret_cast<unsigned int*>(max_norm2 + 1), bits);
}

__device__ __forceinline__ void track_max(float* slot, float v) {      // non-negative floats order like their bit patterns
  if (slot == nullptr) return;
  const unsigned int bits = __float_as_uint(v);
  if (bits > *reinterpret_cast<volatile unsigned int*>(slot)) atomicMax(reinterpret_cast<unsigned int*>(slot), bits);
}

template <int NJ>
__global__ void __launch_bounds__(kThreads) upsert_kernel(const UpsertArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (i >= a.n) return;
  const int64_t row = a.rows ? a.rows[i] : (a.row0 + i);
  const float4* src = reinterpret_cast<const float4*>(a.src + i * a.dim);
  const int dim4 = a.dim >> 2;
  float4 v[NJ][2];
  float ss = 0.0f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = 2 * (lane + 32 * j);                 // float4 index of this lane's group
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      v[j][h] = (c + h < dim4) ? ldg_stream_f4(src + c + h) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      ss = fmaf(v[j][h].x, v[j][h].x, ss); ss = fmaf(v[j][h].y, v[j][h].y, ss);
      ss = fmaf(v[j][h].z, v[j][h].z, ss); ss = fmaf(v[j][h].w, v[j][h].w, ss);
    }
  }
  ss = warp_sum(ss);
  const float scale = a.normalise ? (ss > 0.0f ? rsqrtf(ss) : 0.0f) : 1.0f;
  float stored_ss = 0.0f;
  float lo_ss = 0.0f;                                  // |x - bf16(x)|^2 (fp32 stores with a shadow)
  const int re4 = a.row_elems >> 2;                    // destination row in float4 (fp32) units
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int g = lane + 32 * j;                       // group of 8 elements
    const int c = 2 * g;
    float4 lo = v[j][0], hi = v[j][1];
    lo.x *= scale; lo.y *= scale; lo.z *= scale; lo.w *= scale;
    hi.x *= scale; hi.y *= scale; hi.z *= scale; hi.w *= scale;
    if (a.dtype == 1) {
      if (a.exact != nullptr) {                        // un-rounded fp32 plane for the exact re-ranking
        float4* ex = reinterpret_cast<float4*>(a.exact + row * a.exact_elems);
        if (c < dim4) ex[c] = lo;
        if (c + 1 < dim4) ex[c + 1] = hi;
      }
      uint4 w;
      w.x = pack_bf16x2(lo.x, lo.y); w.y = pack_bf16x2(lo.z, lo.w);
      w.z = pack_bf16x2(hi.x, hi.y); w.w = pack_bf16x2(hi.z, hi.w);
      float t;
      t = bf16_lo(w.x); stored_ss = fmaf(t, t, stored_ss); t = bf16_hi(w.x); stored_ss = fmaf(t, t, stored_ss);
      t = bf16_lo(w.y); stored_ss = fmaf(t, t, stored_ss); t = bf16_hi(w.y); stored_ss = fmaf(t, t, stored_ss);
      t = bf16_lo(w.z); stored_ss = fmaf(t, t, stored_ss); t = bf16_hi(w.z); stored_ss = fmaf(t, t, stored_ss);
      t = bf16_lo(w.w); stored_ss = fmaf(t, t, stored_ss); t = bf16_hi(w.w); stored_ss = fmaf(t, t, stored_ss);
      if (8 * g < a.row_elems)                         // bf16 rows are whole 16-byte chunks (row_elems % 8 == 0)
        reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.vectors) + row * a.row_elems)[g] = w;
    } else {
      stored_ss = fmaf(lo.x, lo.x, stored_ss); stored_ss = fmaf(lo.y, lo.y, stored_ss);
      stored_ss = fmaf(lo.z, lo.z, stored_ss); stored_ss = fmaf(lo.w, lo.w, stored_ss);
      stored_ss = fmaf(hi.x, hi.x, stored_ss); stored_ss = fmaf(hi.y, hi.y, stored_ss);
      stored_ss = fmaf(hi.z, hi.z, stored_ss); stored_ss = fmaf(hi.w, hi.w, stored_ss);
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.vectors) + row * a.row_elems);
      if (c < re4) dst[c] = lo;
      if (c + 1 < re4) dst[c + 1] = hi;
      if (a.shadow != nullptr) {                       // keep the bf16 shadow in step with the row
        const bool hilo = (a.shadow_kind == kShadowHiLo);
        __nv_bfloat16* sh = a.shadow + row * (hilo ? 2 : 1) * a.row_elems;
        const float4 q[2] = {lo, hi};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (c + h >= re4) continue;
          uint2 wh;
          wh.x = pack_bf16x2(q[h].x, q[h].y); wh.y = pack_bf16x2(q[h].z, q[h].w);
          const float r0 = q[h].x - bf16_lo(wh.x), r1 = q[h].y - bf16_hi(wh.x);
          const float r2 = q[h].z - bf16_lo(wh.y), r3 = q[h].w - bf16_hi(wh.y);
          lo_ss = fmaf(r0, r0, lo_ss); lo_ss = fmaf(r1, r1, lo_ss); lo_ss = fmaf(r2, r2, lo_ss); lo_ss = fmaf(r3, r3, lo_ss);
          *reinterpret_cast<uint2*>(sh + 4 * (c + h)) = wh;
          if (hilo) {
            uint2 wl;
            wl.x = pack_bf16x2(r0, r1); wl.y = pack_bf16x2(r2, r3);
            *reinterpret_cast<uint2*>(sh + a.row_elems + 4 * (c + h)) = wl;
          }
        }
      }
    }
  }
  stored_ss = warp_sum(stored_ss);
  if (a.shadow != nullptr && a.lo_max2 != nullptr) lo_ss = warp_sum(lo_ss);
  if (lane == 0) {
    track_norm_bounds(a.max_norm2, stored_ss);
    if (a.shadow != nullptr) track_max(a.lo_max2, lo_ss);
    a.norms2[row] = stored_ss;
    atomicOr(a.live + (row >> 5), 1u << (row & 31));
  }
}

// any dim (dim % 4 != 0 or dim > 2048): two passes over the source row, scalar accesses
__global__ void __launch_bounds__(kThreads) upsert_generic_kernel(const UpsertArgs a) {
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
  float lo_ss = 0.0f;
  if (a.dtype == 1) {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.vectors) + row * a.row_elems;
    for (int e = lane; e < a.row_elems; e += 32) {
      float x = (e < a.dim) ? src[e] * scale : 0.0f;
      __nv_bfloat16 h = __float2bfloat16_rn(x);
      float y = __bfloat162float(h);
      stored_ss = fmaf(y, y, stored_ss);
      dst[e] = h;
      if (a.exact != nullptr && e < a.exact_elems) a.exact[row * a.exact_elems + e] = x;
    }
  } else {
    float* dst = reinterpret_cast<float*>(a.vectors) + row * a.row_elems;
    for (int e = lane; e < a.row_elems; e += 32) {
      float x = (e < a.dim) ? src[e] * scale : 0.0f;
      stored_ss = fmaf(x, x, stored_ss);
      dst[e] = x;
      if (a.shadow) {                  // keep the bf16 shadow in step with the row
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        const float r = x - __bfloat162float(h);
        lo_ss = fmaf(r, r, lo_ss);
        if (a.shadow_kind == kShadowHiLo) {
          a.shadow[row * 2 * a.row_elems + e] = h;
          a.shadow[row * 2 * a.row_elems + a.row_elems + e] = __float2bfloat16_rn(r);
        } else {
          a.shadow[row * a.row_elems + e] = h;
        }
      }
    }
  }
  stored_ss = warp_sum(stored_ss);
  lo_ss = warp_sum(lo_ss);
  if (lane == 0) {
    track_norm_bounds(a.max_norm2, stored_ss);
    if (a.shadow != nullptr) track_max(a.lo_max2, lo_ss);
    a.norms2[row] = stored_ss;
    atomicOr(a.live + (row >> 5), 1u << (row & 31));
  }
}

// bf16 shadow of fp32 rows for the tensor regime, one warp per row.  kShadowHiLo: x = hi + lo + O(2^-18 |x|), two
// bf16 planes side by side that let the tensor cores contract fp32 rows; kShadowHi: bf16(x) only (half the bytes; the
// contraction is then a FILTER whose survivors are re-ranked exactly).  Either way the largest |x - bf16(x)|^2 of a
// row is tracked: with the norms it bounds what the hi-only contraction can miss (refine_kernel's guard).
__global__ void __launch_bounds__(kThreads) split_rows_kernel(const float* vectors, int row_elems, int64_t row0,
                                                             int64_t n, __nv_bfloat16* shadow, int kind, float* lo_max2) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * kWarpsPerCta;
  float worst = 0.0f;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5); r < n; r += warps) {
    const float* src = vectors + (row0 + r) * row_elems;
    __nv_bfloat16* dst = shadow + (row0 + r) * (kind == kShadowHiLo ? 2 : 1) * row_elems;
    float lo_ss = 0.0f;
    for (int e = lane; e < row_elems; e += 32) {
      const float x = src[e];
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      const float rem = x - __bfloat162float(h);
      lo_ss = fmaf(rem, rem, lo_ss);
      dst[e] = h;
      if (kind == kShadowHiLo) dst[row_elems + e] = __float2bfloat16_rn(rem);
    }
    lo_ss = warp_sum(lo_ss);
    worst = fmaxf(worst, lo_ss);
  }
  if (lane == 0 && worst > 0.0f) track_max(lo_max2, worst);
}

__global__ void clear_live_kernel(uint32_t* live, const int64_t* rows, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = rows[i];
  atomicAnd(live + (row >> 5), ~(1u << (row & 31)));
}

__global__ void patch_mask_kernel(uint32_t* mask, const int64_t* rows, const unsigned char* pass, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t row = rows[i];
  if (pass[i]) atomicOr(mask + (row >> 5), 1u << (row & 31));
  else atomicAnd(mask + (row >> 5), ~(1u << (row & 31)));
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
  float n2 = 0.0f, lo2 = 0.0f;
  for (int e = lane; e < a.row_elems; e += 32) {
    float x = (e < a.dim) ? src[e] * scale : 0.0f;
    if (a.q_exact != nullptr && e < a.exact_elems) a.q_exact[static_cast<size_t>(b) * a.exact_elems + e] = x;
    if (a.round_bf16) x = round_bf16(x);
    n2 = fmaf(x, x, n2);
    { const float rem = x - round_bf16(x); lo2 = fmaf(rem, rem, lo2); }
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
  if (a.q_lo_norm2) {
    lo2 = warp_sum(lo2);
    if (lane == 0) a.q_lo_norm2[b] = lo2;
  }
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
  const unsigned ctas = static_cast<unsigned>((a.n + kWarpsPerCta - 1) / kWarpsPerCta);
  const int width = a.dim > a.row_elems ? a.dim : a.row_elems;
  const int groups = (width + 255) / 256;                // groups of 8 elements per lane
  if (a.dim % 4 != 0 || groups > 8) upsert_generic_kernel<<<ctas, kThreads, 0, st>>>(a);
  else if (groups <= 1) upsert_kernel<1><<<ctas, kThreads, 0, st>>>(a);
  else if (groups <= 2) upsert_kernel<2><<<ctas, kThreads, 0, st>>>(a);
  else if (groups <= 3) upsert_kernel<3><<<ctas, kThreads, 0, st>>>(a);
  else if (groups <= 4) upsert_kernel<4><<<ctas, kThreads, 0, st>>>(a);
  else if (groups <= 6) upsert_kernel<6><<<ctas, kThreads, 0, st>>>(a);
  else upsert_kernel<8><<<ctas, kThreads, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_patch_mask(uint32_t* mask, const int64_t* rows_dev, const unsigned char* pass_dev, int64_t n,
                              cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  patch_mask_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(mask, rows_dev, pass_dev, n);
  return cudaGetLastError();
}

cudaError_t launch_split_rows(const float* vectors, int row_elems, int64_t row0, int64_t n, __nv_bfloat16* shadow,
                              int shadow_kind, float* lo_max2, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  int64_t ctas = (n + kWarpsPerCta - 1) / kWarpsPerCta;
  if (ctas > 148 * 16) ctas = 148 * 16;
  split_rows_kernel<<<static_cast<unsigned>(ctas), kThreads, 0, st>>>(vectors, row_elems, row0, n, shadow, shadow_kind, lo_max2);
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
