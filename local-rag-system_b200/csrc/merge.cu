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
      key = key_to_global(a.rows_map, key);
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

// Small fan-in fast path (cross-shard merge: S = number of GPUs <= 32, k <= 64): one warp per
// query, lane s walks shard s's sorted list, k rounds of warp-wide min.  No shared memory, no
// block barriers: ~1 us instead of the ~4 us of the general kernel -- it sits on the critical
// path of every multi-GPU query.
__global__ void __launch_bounds__(kMergeThreads) merge_small_kernel(const MergeArgs a) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * kMergeWarps + (threadIdx.x >> 5);
  if (b >= a.B) return;
  const int k = a.k;
  const bool own = lane < a.S;
  const uint64_t* lst = a.keys + (static_cast<size_t>(own ? lane : 0) * a.B + b) * k;
  int pos = 0;
  uint64_t cur = own ? __ldcg(lst) : kEmptyKey;
  uint64_t nxt = (own && k > 1) ? __ldcg(lst + 1) : kEmptyKey;
  int cnt = 0;
  for (int r = 0; r < k; ++r) {
    uint64_t m = cur;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const uint64_t o = shfl_u64(m, lane ^ off);
      m = o < m ? o : m;
    }
    const bool valid = m != kEmptyKey;
    if (valid && cur == m) {          // keys are unique across shards: exactly one lane advances
      cur = nxt;
      pos++;
      nxt = (pos + 1 < k) ? __ldcg(lst + pos + 1) : kEmptyKey;
    }
    if (lane == 0) {
      uint64_t key = m;
      if (valid) key = key_to_global(a.rows_map, key);
      const size_t o = static_cast<size_t>(b) * k + r;
      if (a.out_keys) a.out_keys[o] = key;
      if (a.out_rows) a.out_rows[o] = valid ? static_cast<int64_t>(key_row(key)) : -1;
      if (a.out_dists) a.out_dists[o] = valid ? key_dist(key) : __int_as_float(0x7f800000);
    }
    cnt += valid ? 1 : 0;
  }
  if (lane == 0 && a.out_counts) a.out_counts[b] = cnt;
}

// Medium fan-in (33 <= S <= 512 lists, k <= 128: the tensor regime's per-CTA lists when the batch has few query
// tiles, e.g. S = 148 at B <= 128): one CTA per query, thread t owns lists t and t + 256 and keeps each list's head
// and next key in registers; k rounds of block-wide minimum (two barriers each, no memory latency unless the same
// list wins twice in a row).  The general kernel above walks the lists warp by warp with cooperative insertions:
// 34.8 us for S = 148, k = 16, B = 32 (ncu, profiles/r02_launches_f32_b32.csv) -- 8 % of that step.
__global__ void __launch_bounds__(kMergeThreads) merge_kway_kernel(const MergeArgs a) {
  __shared__ uint64_t wmin[kMergeWarps];
  __shared__ int total;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const int k = a.k;
  uint64_t cur[2], nxt[2];
  int pos[2];
  const uint64_t* lst[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int s = threadIdx.x + u * kMergeThreads;
    lst[u] = a.keys + (static_cast<size_t>(s < a.S ? s : 0) * a.B + b) * k;
    pos[u] = 0;
    cur[u] = (s < a.S) ? __ldcg(lst[u]) : kEmptyKey;
    nxt[u] = (s < a.S && k > 1) ? __ldcg(lst[u] + 1) : kEmptyKey;
  }
  if (threadIdx.x == 0) total = 0;
  for (int r = 0; r < k; ++r) {
    uint64_t m = cur[0] < cur[1] ? cur[0] : cur[1];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const uint64_t o = shfl_u64(m, lane ^ off);
      m = o < m ? o : m;
    }
    if (lane == 0) wmin[warp] = m;
    __syncthreads();
    uint64_t w = wmin[0];
#pragma unroll
    for (int i = 1; i < kMergeWarps; ++i) w = wmin[i] < w ? wmin[i] : w;
    if (threadIdx.x == 0) {
      uint64_t key = w;
      const bool valid = key != kEmptyKey;
      if (valid) { key = key_to_global(a.rows_map, key); total++; }
      const size_t o = static_cast<size_t>(b) * k + r;
      if (a.out_keys) a.out_keys[o] = key;
      if (a.out_rows) a.out_rows[o] = valid ? static_cast<int64_t>(key_row(key)) : -1;
      if (a.out_dists) a.out_dists[o] = valid ? key_dist(key) : __int_as_float(0x7f800000);
    }
    if (w != kEmptyKey) {                // keys are unique: exactly one list head equals w
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (cur[u] == w) {
          cur[u] = nxt[u];
          pos[u]++;
          nxt[u] = (pos[u] + 1 < k) ? __ldcg(lst[u] + pos[u] + 1) : kEmptyKey;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && a.out_counts) a.out_counts[b] = total;
}

// ---- exact re-scoring of the approximate winners (tensor regime) --------------------------------
// The tensor kernel ranks l2 by |q|^2 + |x|^2 - 2 q.x (cancellation error grows with the norms) and,
// for fp32 stores, contracts bf16 hi/lo splits (~1e-6 absolute error on a unit-norm dot product).
// The k_in >= k survivors per query are re-scored here from the stored rows with the same direct
// fp32 arithmetic the stream kernel uses, re-sorted, and the best k emitted -- so distances meet
// the same tolerance in both regimes and near-ties at the k-th place are decided exactly.
// One CTA per query, one warp per row.
__global__ void __launch_bounds__(kMergeThreads) refine_kernel(const RefineArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const int k = a.k;
  const int kin = a.k_in;
  const int kpad = next_pow2(kin);
  for (int i = threadIdx.x; i < kpad; i += blockDim.x) keys[i] = kEmptyKey;
  __syncthreads();
  const float* q = a.queries + static_cast<size_t>(b) * a.row_elems;
  for (int j = warp; j < kin; j += kMergeWarps) {
    const uint64_t key = a.keys[static_cast<size_t>(b) * kin + j];
    if (key == kEmptyKey) continue;
    const uint32_t row = key_row(key);
    float acc = 0.0f;
    if (a.dtype == 1) {
      const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(a.vectors) + static_cast<size_t>(row) * a.row_elems;
      for (int e = lane; e < a.row_elems; e += 32) {
        const float xv = __bfloat162float(x[e]);
        if (a.l2) { const float d = xv - q[e]; acc = fmaf(d, d, acc); } else { acc = fmaf(xv, q[e], acc); }
      }
    } else {
      const float* x = reinterpret_cast<const float*>(a.vectors) + static_cast<size_t>(row) * a.row_elems;
      for (int e = lane; e < a.row_elems; e += 32) {
        if (a.l2) { const float d = x[e] - q[e]; acc = fmaf(d, d, acc); } else { acc = fmaf(x[e], q[e], acc); }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) keys[j] = make_key(a.l2 ? acc : 1.0f - acc, row);
  }
  __syncthreads();
  block_bitonic_sort(keys, kpad);
  if (a.redo_count != nullptr && threadIdx.x == 0) {
    // exactness guard: candidates were the k_in best by APPROXIMATE distance.  A row outside the
    // list has approx >= a_max, hence exact >= a_max - eps; it cannot displace the exact k-th
    // best d_k (which may itself be over-estimated by eps) when a_max - d_k > 2 eps.
    const uint64_t worst_in = a.keys[static_cast<size_t>(b) * kin + kin - 1];
    if (worst_in != kEmptyKey && keys[k - 1] != kEmptyKey) {
      const float a_max = key_dist(worst_in);
      const float d_k = key_dist(keys[k - 1]);
      // hi-only contraction: plus what rounding q and x to bf16 can have moved (common.cuh)
      const float eps = bf16_contraction_eps(a.q_norm2[b], a.q_lo_norm2 ? a.q_lo_norm2[b] : 0.0f, a.x_max_norm2[0],
                                             a.q_lo_norm2 ? a.x_lo_max2[0] : 0.0f, a.guard_rel, a.l2 != 0);
      if (!(a_max - d_k > 2.0f * eps)) a.redo_list[atomicAdd(a.redo_count, 1)] = b;
    }
  }
  int cnt = 0;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    uint64_t key = keys[j];
    const bool valid = key != kEmptyKey;
    if (valid) {
      key = key_to_global(a.rows_map, key);
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

// ---- hi-only filter: exact re-scoring of every row within 2 eps of the approximate k-th best (kernels.h) ------
constexpr int kFiltRefineCap = 512;       // candidates per query the re-scoring takes; more -> exact re-run
__global__ void __launch_bounds__(kMergeThreads) refine_filter_kernel(const RefineFilterArgs fa) {
  const RefineArgs& a = fa.r;
  __shared__ uint64_t cand[kFiltRefineCap];
  __shared__ int s_n, s_over;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  const int k = a.k;
  if (threadIdx.x == 0) { s_n = 0; s_over = 0; }
  for (int i = threadIdx.x; i < kFiltRefineCap; i += blockDim.x) cand[i] = kEmptyKey;
  __syncthreads();
  // threshold: approximate k-th best + 2 eps (+inf while fewer than k rows exist: everything buffered is a candidate)
  float thr = __int_as_float(0x7f800000);
  const uint64_t kth = a.keys[static_cast<size_t>(b) * a.k_in + (k - 1)];
  if (kth != kEmptyKey)
    thr = key_dist(kth) + 2.0f * bf16_contraction_eps(a.q_norm2[b], a.q_lo_norm2[b], a.x_max_norm2[0], a.x_lo_max2[0],
                                                       a.guard_rel, a.l2 != 0);
  for (int c = threadIdx.x; c < fa.S; c += blockDim.x) {      // one thread per CTA of the contraction: its few buffered rows
    const size_t slot = static_cast<size_t>(c) * a.B + b;
    const int cnt = fa.extra_cnt[slot];
    if (cnt < 0 || cnt > fa.cap) { atomicExch(&s_over, 1); continue; }
    for (int i0 = 0; i0 < cnt; i0 += 8) {           // 8 independent loads per round (a CTA leaves ~15-20 rows, few pass)
      uint64_t kk[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) kk[u] = (i0 + u < cnt) ? __ldcg(fa.extra + slot * fa.cap + i0 + u) : kEmptyKey;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (kk[u] != kEmptyKey && key_dist(kk[u]) <= thr) {
          const int p = atomicAdd(&s_n, 1);
          if (p < kFiltRefineCap) cand[p] = kk[u]; else atomicExch(&s_over, 1);
        }
      }
    }
  }
  __syncthreads();
  const int n = min(s_n, kFiltRefineCap);
  const float* q = a.queries + static_cast<size_t>(b) * a.row_elems;
  for (int j = warp; j < n; j += kMergeWarps) {       // one warp per candidate, the stream kernel's direct fp32 arithmetic
    const uint32_t row = key_row(cand[j]);
    const float* x = reinterpret_cast<const float*>(a.vectors) + static_cast<size_t>(row) * a.row_elems;
    float acc = 0.0f;
    for (int e = lane; e < a.row_elems; e += 32) {
      if (a.l2) { const float d = x[e] - q[e]; acc = fmaf(d, d, acc); } else { acc = fmaf(x[e], q[e], acc); }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    __syncwarp();
    if (lane == 0) cand[j] = make_key(a.l2 ? acc : 1.0f - acc, row);
  }
  __syncthreads();
  int npad = next_pow2(n > k ? n : k);
  if (npad > kFiltRefineCap) npad = kFiltRefineCap;
  block_bitonic_sort(cand, npad);
  if (threadIdx.x == 0 && s_over && a.redo_count != nullptr) a.redo_list[atomicAdd(a.redo_count, 1)] = b;
  int cnt = 0;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    uint64_t key = cand[j];
    const bool valid = key != kEmptyKey;
    if (valid) { key = key_to_global(a.rows_map, key); cnt++; }
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

cudaError_t launch_refine_filter(const RefineFilterArgs& a, cudaStream_t st) {
  if (a.r.B <= 0 || a.r.k <= 0 || a.r.k_in < a.r.k || a.r.k > kFiltRefineCap || a.S <= 0 || a.cap <= 0) return cudaErrorInvalidValue;
  if (a.r.dtype == 1 || a.r.q_lo_norm2 == nullptr || a.r.x_lo_max2 == nullptr) return cudaErrorInvalidValue;
  refine_filter_kernel<<<a.r.B, kMergeThreads, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_refine(const RefineArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.k <= 0 || a.k_in < a.k) return cudaErrorInvalidValue;
  const size_t smem = static_cast<size_t>(next_pow2(a.k_in)) * sizeof(uint64_t);
  refine_kernel<<<a.B, kMergeThreads, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_merge(const MergeArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.k <= 0 || a.S <= 0) return cudaErrorInvalidValue;
  if (a.S <= 32 && a.k <= 64) {
    merge_small_kernel<<<(a.B + kMergeWarps - 1) / kMergeWarps, kMergeThreads, 0, st>>>(a);
    return cudaGetLastError();
  }
  if (a.S <= 2 * kMergeThreads && a.k <= 128) {
    merge_kway_kernel<<<a.B, kMergeThreads, 0, st>>>(a);
    return cudaGetLastError();
  }
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
