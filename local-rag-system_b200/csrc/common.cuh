// Shared device/host helpers for the B200 dense-retrieval engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace rag {

constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr int kWarp = 32;

// ---- sortable keys --------------------------------------------------------
// key = (ordered(fp32 distance) << 32) | row.  `ordered` maps IEEE-754 floats
// to unsigned ints with the same ordering, so "k smallest keys" is "k smallest
// distances, ties broken by lower row" -- the oracle's order
// (oracle/exact_search.py: topk_stable).
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float dist, uint32_t row) {
  return (static_cast<uint64_t>(float_to_ordered(dist)) << 32) | row;
}
__host__ __device__ __forceinline__ float key_dist(uint64_t k) {
  return ordered_to_float(static_cast<uint32_t>(k >> 32));
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) {
  return static_cast<uint32_t>(k & 0xFFFFFFFFull);
}

// ---- local row -> global row --------------------------------------------------------------
// A shard emits keys whose row field is the GLOBAL row, so a cross-shard merge is a plain 64-bit
// compare.  Two placements:
//   contiguous (shift == 0): global = local + base                    one process per GPU, rank g owns
//                                                                     [g * stride, g * stride + rows_g)
//   striped    (shift  > 0): chunks of 2^shift rows dealt round-robin to `world` shards; shard g holds
//                            chunks g, g + world, ...; base = g << shift:
//                            global = ((local >> shift) * world << shift) + base + (local & (2^shift - 1))
//                            (the single-process multi-device store: global rows stay dense while every
//                            shard grows at the same rate)
struct RowMap {
  uint32_t base;
  uint32_t shift;
  uint32_t world;
};
__host__ __device__ __forceinline__ uint32_t to_global_row(const RowMap m, uint32_t local) {
  if (m.shift == 0u) return local + m.base;
  return (((local >> m.shift) * m.world) << m.shift) + m.base + (local & ((1u << m.shift) - 1u));
}
__host__ __device__ __forceinline__ uint64_t key_to_global(const RowMap m, uint64_t key) {
  return (key & 0xFFFFFFFF00000000ull) | static_cast<uint64_t>(to_global_row(m, static_cast<uint32_t>(key)));
}

__host__ __device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

#ifdef __CUDACC__
// ---- error bound of a bf16 contraction of fp32 vectors ----------------------------------------------
// |q.x - bf16(q).bf16(x)| <= |q - qh| |x| + |qh| |x - xh| (Cauchy-Schwarz) with |q - qh| known per query and |x|,
// |x - xh| bounded by their maxima over the store, plus guard_rel |q||x| for the fp32 accumulation and the
// rounding of the distances themselves.  l2 distances |q|^2 + |x|^2 - 2 q.x move by twice that.  Used by the
// hi-only filter of fp32 stores (tensor_regime.cu) and by the kernels that certify its result (merge.cu): both
// sides must compute the SAME number, hence one function.  Arguments are squared norms.
__device__ __forceinline__ float bf16_contraction_eps(float q_norm2, float q_lo_norm2, float x_max_norm2, float x_lo_max2,
                                                      float guard_rel, bool l2) {
  const float qn = sqrtf(q_norm2), xn = sqrtf(x_max_norm2), ql = sqrtf(q_lo_norm2), xl = sqrtf(x_lo_max2);
  const float eps = guard_rel * qn * xn + 1.01f * (ql * xn + (qn + ql) * xl);
  return l2 ? 2.0f * eps : eps;
}

// ---- streaming 16-byte load: read-only path, do not allocate in L1 ---------
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// ---- system-scope accesses for the cross-GPU exchange over peer-mapped memory ----
__device__ __forceinline__ void st_relaxed_sys_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), src);
  uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// ---- warp-cooperative insertion into an ascending list in shared memory ----
// L[0..k) sorted ascending (kEmptyKey padded).  Precondition: key < L[k-1].
// All 32 lanes call with the same arguments.
__device__ __forceinline__ void warp_list_insert(uint64_t* L, int k, uint64_t key, int lane) {
  int pos = 0;
  for (int base = 0; base < k; base += kWarp) {
    int i = base + lane;
    uint64_t e = (i < k) ? L[i] : kEmptyKey;
    pos += __popc(__ballot_sync(0xffffffffu, e < key));
  }
  for (int base = ((k - 1) / kWarp) * kWarp; base >= 0; base -= kWarp) {
    int i = base + lane;
    bool mv = (i < k) && (i > pos);
    uint64_t prev = mv ? L[i - 1] : 0;
    __syncwarp();
    if (mv) L[i] = prev;
    __syncwarp();
  }
  if (lane == 0) L[pos] = key;
  __syncwarp();
}

// ---- CTA-wide bitonic sort of n (power of two) keys in shared memory ---------
__device__ __forceinline__ void block_bitonic_sort(uint64_t* a, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        uint64_t x = a[lo], y = a[hi];
        if ((x > y) == up) { a[lo] = y; a[hi] = x; }
      }
    }
  }
  __syncthreads();
}
#endif  // __CUDACC__

}  // namespace rag
