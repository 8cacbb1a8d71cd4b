// K4b / K6: merge S ascending candidate lists per query into the
// final top-k.  Used twice on the path behind collection.query
// (api/app.py:544-549):
//   * after the scan, S = number of persistent CTAs of the scan grid;
//   * after the NCCL all-gather of per-shard candidates, S = number of shards.
// One CTA per query; every warp walks a strided subset of the source lists with
// an early exit (lists are sorted, so the first key that fails the running
// threshold ends that list), then the 8 warp lists are sorted together.
#include "common.cuh"
#include "kernels.h"

namespace rag {
namespace {

constexpr int kMergeThreads = 256;
constexpr int kMergeWarps = kMergeThreads / 32;

__global__ void __launch_bounds__(kMergeThreads) merge_kernel(const MergeArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const int k = a.k;
  const int kpad = next_pow2(k);

  for (int i = threadIdx.x; i < kMergeWarps * kpad; i += blockDim.x) lists[i] = kEmptyKey;
  __syncthreads();

  uint64_t* L = lists + static_cast<size_t>(warp) * kpad;
  for (int s = warp; s < a.S; s += kMergeWarps) {
    const uint64_t* src = a.keys + (static_cast<size_t>(s) * a.B + b) * k;
    bool done = false;
    for (int j0 = 0; j0 < k && !done; j0 += 32) {
      const int j = j0 + lane;
      const uint64_t cand = (j < k) ? src[j] : kEmptyKey;
      uint32_t bal = __ballot_sync(0xffffffffu, cand < L[k - 1]);
      if (bal != 0xffffffffu) done = true;   // some entry failed: the rest of this list fails too
      while (bal) {
        const int sl = __ffs(bal) - 1;
        bal &= (bal - 1);
        const uint64_t ck = shfl_u64(cand, sl);
        if (ck < L[k - 1]) warp_list_insert(L, k, ck, lane);
        else bal = 0;                        // ascending source: nothing later can pass
      }
    }
  }
  __syncthreads();
  block_bitonic_sort(lists, kMergeWarps * kpad);

  int cnt = 0;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    uint64_t key = lists[j];
    const bool valid = key != kEmptyKey;
    if (valid) {
      key = (key & 0xFFFFFFFF00000000ull) | static_cast<uint64_t>(key_row(key) + a.row_base);
      cnt++;
    }
    const size_t o = static_cast<size_t>(b) * k + j;
    if (a.out_keys) a.out_keys[o] = key;
    if (a.out_rows) a.out_rows[o] = valid ? static_cast<int64_t>(key_row(key)) : -1;
    if (a.out_dists) a.out_dists[o] = valid ? key_dist(key) : __int_as_float(0x7f800000);
  }
  if (a.out_counts) {
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    if (cnt) atomicAdd(&total, cnt);
    __syncthreads();
    if (threadIdx.x == 0) a.out_counts[b] = total;
  }
}

}  // namespace

cudaError_t launch_merge(const MergeArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.k <= 0 || a.S <= 0) return cudaErrorInvalidValue;
  const size_t smem = static_cast<size_t>(kMergeWarps) * next_pow2(a.k) * sizeof(uint64_t);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  merge_kernel<<<a.B, kMergeThreads, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace rag
