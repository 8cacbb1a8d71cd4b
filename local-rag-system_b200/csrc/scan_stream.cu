// K2 + K4: HBM-streaming exact scan with fused top-k.
//
// Replaces, for small query batches, what Chroma's BruteForceIndex.query /
// hnswlib knn_query compute behind collection.query (reference call site
// api/app.py:544-549): distances from every live, filter-passing corpus row to
// each query, then the k smallest.  The B x N distance matrix never exists:
// every warp keeps a running sorted top-k list per query in shared memory and
// only rows beating the list's current k-th key are inserted.
//
// Roofline: HBM.  Algorithmic bytes per launch = live rows x row_bytes (+ one
// bitmap word per 32 rows); every corpus byte is read exactly once per launch
// (up to QB queries share the pass).
//
// Work decomposition
//   grid.x = persistent CTAs (2 per SM), grid.y = query groups of QB
//   one warp <- one block of 32 consecutive rows (= one word of the live and
//   filter bitmaps), blocks interleaved over all warps of the grid.  Dead or
//   filtered rows are skipped without touching their bytes (rows are whole
//   128-byte lines for the dims that matter), so a selective `where` reads
//   ~selectivity x corpus bytes.
//   Within a block the warp takes R live rows at a time; every lane issues
//   R x NJ independent 16-byte streaming loads (no L1 allocation) before the
//   first FMA, so 16 resident warps/SM keep ~96 KB/SM in flight.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace rag {
namespace {

__device__ __forceinline__ float bf_lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf_hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }

// ---- multi-value warp reduction ("transposed butterfly") -----------------------
// Reduces V (power of two <= 32) per-lane partial sums across the warp with
// V-1 + (5 - log2 V) shuffles instead of 5 V.  On return lane l holds the
// warp total of value index  l >> (5 - log2 V).
template <int C, int OFF>
struct Fold {
  template <int V>
  static __device__ __forceinline__ float run(float (&v)[V], int lane) {
    const bool upper = (lane & OFF) != 0;
#pragma unroll
    for (int i = 0; i < C / 2; ++i) {
      float send = upper ? v[i] : v[i + C / 2];
      float keep = upper ? v[i + C / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
    return Fold<C / 2, OFF / 2>::run(v, lane);
  }
};
template <int OFF>
struct Fold<1, OFF> {
  template <int V>
  static __device__ __forceinline__ float run(float (&v)[V], int) {
    float x = v[0];
#pragma unroll
    for (int off = OFF; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
  }
};

template <int V> struct Log2 { static constexpr int value = 1 + Log2<V / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

// one 16-byte chunk of a row against the matching query elements
template <bool BF16, bool L2>
__device__ __forceinline__ float chunk_acc(float acc, const uint4& v, const float4& qa, const float4& qb) {
  if constexpr (BF16) {
    float x0 = bf_lo(v.x), x1 = bf_hi(v.x), x2 = bf_lo(v.y), x3 = bf_hi(v.y);
    float x4 = bf_lo(v.z), x5 = bf_hi(v.z), x6 = bf_lo(v.w), x7 = bf_hi(v.w);
    if constexpr (L2) {
      x0 -= qa.x; x1 -= qa.y; x2 -= qa.z; x3 -= qa.w;
      x4 -= qb.x; x5 -= qb.y; x6 -= qb.z; x7 -= qb.w;
      acc = fmaf(x0, x0, acc); acc = fmaf(x1, x1, acc); acc = fmaf(x2, x2, acc); acc = fmaf(x3, x3, acc);
      acc = fmaf(x4, x4, acc); acc = fmaf(x5, x5, acc); acc = fmaf(x6, x6, acc); acc = fmaf(x7, x7, acc);
    } else {
      acc = fmaf(x0, qa.x, acc); acc = fmaf(x1, qa.y, acc); acc = fmaf(x2, qa.z, acc); acc = fmaf(x3, qa.w, acc);
      acc = fmaf(x4, qb.x, acc); acc = fmaf(x5, qb.y, acc); acc = fmaf(x6, qb.z, acc); acc = fmaf(x7, qb.w, acc);
    }
  } else {
    float x0 = __uint_as_float(v.x), x1 = __uint_as_float(v.y), x2 = __uint_as_float(v.z), x3 = __uint_as_float(v.w);
    if constexpr (L2) {
      x0 -= qa.x; x1 -= qa.y; x2 -= qa.z; x3 -= qa.w;
      acc = fmaf(x0, x0, acc); acc = fmaf(x1, x1, acc); acc = fmaf(x2, x2, acc); acc = fmaf(x3, x3, acc);
    } else {
      acc = fmaf(x0, qa.x, acc); acc = fmaf(x1, qa.y, acc); acc = fmaf(x2, qa.z, acc); acc = fmaf(x3, qa.w, acc);
    }
  }
  return acc;
}

// BF16: element type.  QB queries share one pass.  LPR lanes cooperate on one row (32, or a
// sub-warp of 16 / 8 for narrow rows, so that e.g. a 768-byte row = 3 x 16 lanes x 16 B keeps
// every lane busy; the 32/LPR sub-warps of a warp work on different rows).  NJ > 0: a row is
// exactly NJ x LPR chunks (fully unrolled, no guards); NJ == 0: any row length (LPR = 32).
// R row slots in flight per warp, i.e. R x 32/LPR rows.
template <bool BF16, int QB, int NJ, int R, bool L2, int LPR>
__global__ void __launch_bounds__(kScanThreads, 2) scan_stream_kernel(const ScanArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int V = R * QB;
  constexpr int G = 32 / LPR;                      // rows per slot
  static_assert(V <= LPR, "R*QB must fit the lanes of one row");
  static_assert(NJ > 0 || LPR == 32, "generic row length needs the full warp");
  constexpr int SHIFT = Log2<LPR>::value - Log2<V>::value;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int cpr = a.cpr;
  const int row_elems = a.row_elems;
  const int k = a.k;
  const int kpad = next_pow2(k);
  const int b0 = blockIdx.y * QB;
  // ---- programmatic dependent launch, part 1 -----------------------------------------------------
  // Allow the NEXT launch on the stream to be scheduled as soon as every CTA of this one has
  // started: its CTAs then take over each SM slot the moment one of ours exits, so back-to-back
  // queries overlap the stragglers, the cross-CTA merge and the cross-GPU exchange of one query
  // with the scan of the next (worth ~7 % at a 1.9 GB shard).  See part 2 for the other half.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // optional indirection (exact re-run of the few queries the split-precision tensor regime could
  // not certify): the launch covers the worst case, the device-side count says how many are real
  const int nB = a.q_count ? min(*a.q_count, a.B) : a.B;
  if (b0 >= nB) return;

  float* q_s = reinterpret_cast<float*>(smem_raw);
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(QB) * row_elems * sizeof(float));

  if (a.query_flag != nullptr) {       // the queries come down on a copy stream: wait until they have landed
    if (threadIdx.x == 0) {
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys_u32(a.query_flag) != a.query_seq) {
        if (global_timer_ns() - t0 > 5000000000ull) __trap();      // 5 s: the copy failed; surface it as a launch error
      }
    }
    __syncthreads();
  }

  // ---- stage queries (fp32, chunk-major so that lanes read consecutive float4) ----
  // With queries_raw the CTA prepares them itself exactly as prep_queries_kernel would (same
  // lane-strided summation order -> bit-identical values in both regimes): cosine stores scale
  // by 1/|q|, bf16 stores round to bf16.
  __shared__ float s_scale[QB];
  auto raw_q = [&](int bs, int e) -> float { return a.queries_raw[static_cast<size_t>(bs) * a.dim + e]; };
  const bool have_raw = a.queries_raw != nullptr;
  if (have_raw) {
    if (warp < QB) {
      const int b = b0 + warp;
      float ss = 0.0f;
      if (a.normalise && b < nB) {
        const int bs = a.q_index ? a.q_index[b] : b;
        for (int e = lane; e < a.dim; e += 32) { const float x = raw_q(bs, e); ss = fmaf(x, x, ss); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
      }
      if (lane == 0) s_scale[warp] = a.normalise ? (ss > 0.0f ? rsqrtf(ss) : 0.0f) : 1.0f;
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < QB * row_elems; idx += blockDim.x) {
    int qb = idx / row_elems, e = idx - qb * row_elems;
    int b = b0 + qb;
    float val = 0.0f;
    if (b < nB) {
      if (have_raw) {
        const int bs = a.q_index ? a.q_index[b] : b;
        val = (e < a.dim) ? raw_q(bs, e) * s_scale[qb] : 0.0f;
        if (a.round_bf16) val = __bfloat162float(__float2bfloat16_rn(val));
      } else {
        val = a.queries[static_cast<size_t>(b) * row_elems + e];
      }
    }
    int dst;
    if constexpr (BF16) {
      int c = e >> 3, w = e & 7;
      dst = qb * row_elems + (w >> 2) * (cpr * 4) + c * 4 + (w & 3);
    } else {
      dst = qb * row_elems + e;
    }
    q_s[dst] = val;
  }
  for (int idx = threadIdx.x; idx < QB * kScanWarps * kpad; idx += blockDim.x) lists[idx] = kEmptyKey;
  __syncthreads();

  // lane's value index after the multi-reduce, fixed for the whole kernel
  const int sl = lane & (LPR - 1);                 // lane within its row group
  const int grp = lane / LPR;                      // which of the slot's G rows this lane works on
  const int my_idx = sl >> SHIFT;
  const int my_i = my_idx / QB;
  const int my_qb = my_idx - my_i * QB;
  const bool rep = (sl & ((1 << SHIFT) - 1)) == 0;
  uint64_t my_tau = kEmptyKey;   // current k-th best key of (this warp, my_qb)

  const uint4* vec = reinterpret_cast<const uint4*>(a.vectors);
  const int64_t nblk = (a.n_rows + kRowsPerBlock - 1) / kRowsPerBlock;
  const int64_t wstride = static_cast<int64_t>(gridDim.x) * kScanWarps;

  // One step of the scan: the R row slots `r[]` (row index relative to `base` as an UNSIGNED 32-bit value, -1 =
  // 0xFFFFFFFF = empty slot -- no store has that many rows; with
  // sub-warp rows each lane group carries its own row per slot) are loaded, contracted with the
  // queries and offered to the top-k lists.  `row0` = store row of `base`.
  auto process = [&](const uint4* base, int64_t row0, const int (&r)[R]) {
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.0f;

    if constexpr (NJ > 0) {
      uint4 v[R][NJ];
#pragma unroll
      for (int i = 0; i < R; ++i) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          v[i][j] = (r[i] != -1) ? ldg_stream(base + static_cast<size_t>(static_cast<uint32_t>(r[i])) * cpr + j * LPR + sl)
                                 : make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = j * LPR + sl;
#pragma unroll
        for (int qb = 0; qb < QB; ++qb) {
          const float4* qrow = reinterpret_cast<const float4*>(q_s + qb * row_elems);
          float4 qa = qrow[c];
          float4 qh = BF16 ? qrow[cpr + c] : qa;
#pragma unroll
          for (int i = 0; i < R; ++i) acc[i * QB + qb] = chunk_acc<BF16, L2>(acc[i * QB + qb], v[i][j], qa, qh);
        }
      }
    } else {
      for (int c = lane; c < cpr; c += 32) {
        uint4 v[R];
#pragma unroll
        for (int i = 0; i < R; ++i)
          v[i] = (r[i] != -1) ? ldg_stream(base + static_cast<size_t>(static_cast<uint32_t>(r[i])) * cpr + c) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int qb = 0; qb < QB; ++qb) {
          const float4* qrow = reinterpret_cast<const float4*>(q_s + qb * row_elems);
          float4 qa = qrow[c];
          float4 qh = BF16 ? qrow[cpr + c] : qa;
#pragma unroll
          for (int i = 0; i < R; ++i) acc[i * QB + qb] = chunk_acc<BF16, L2>(acc[i * QB + qb], v[i], qa, qh);
        }
      }
    }

    const float total = Fold<V, LPR / 2>::run(acc, lane);
    int myr = r[0];
#pragma unroll
    for (int i = 1; i < R; ++i) myr = (my_i == i) ? r[i] : myr;
    const float dist = L2 ? total : (1.0f - total);
    const uint64_t key = make_key(dist, static_cast<uint32_t>(row0) + static_cast<uint32_t>(myr));
    uint32_t bal = __ballot_sync(0xffffffffu, rep && (myr != -1) && (key < my_tau));
    while (bal) {
      const int src = __ffs(bal) - 1;
      bal &= (bal - 1);
      const uint64_t ck = shfl_u64(key, src);
      const int sidx = (src & (LPR - 1)) >> SHIFT;
      const int cqb = sidx - (sidx / QB) * QB;
      uint64_t* L = lists + (static_cast<size_t>(cqb) * kScanWarps + warp) * kpad;
      if (ck < L[k - 1]) {   // tau may have tightened since the ballot
        warp_list_insert(L, k, ck, lane);
        if (my_qb == cqb) my_tau = L[k - 1];
      }
    }
  };

  if (a.filter == nullptr) {
    // ---- dense walk: one 32-row block (one bitmap word) per warp step ---------------------------
    for (int64_t blk = static_cast<int64_t>(blockIdx.x) * kScanWarps + warp; blk < nblk; blk += wstride) {
      uint32_t m = __ldg(a.live + blk);
      const uint4* bbase = vec + static_cast<size_t>(blk) * kRowsPerBlock * cpr;
      while (m) {
        int r[R];                                      // this lane's row in each slot (-1: none)
#pragma unroll
        for (int i = 0; i < R; ++i) {
          r[i] = -1;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int bit = m ? (__ffs(m) - 1) : -1;
            m &= (m - 1);
            if (G == 1 || g == grp) r[i] = bit;
          }
        }
        process(bbase, blk * kRowsPerBlock, r);
      }
    }
  } else {
    // ---- filtered walk (`where` bitmap): gather the passing rows of a 256-row chunk ------------
    // With a selective filter most 32-row blocks hold 0-3 passing rows; walking them block by
    // block leaves the row slots (= the loads in flight) mostly empty and pays one dependent
    // bitmap load per block.  Here lanes 0..7 fetch the 8 bitmap words of a chunk at once (the
    // next chunk's words are requested before this one is processed), a warp scan ranks the set
    // bits, and every step fills all R x G slots with the next passing rows of the chunk.
    constexpr int CW = 16;                              // bitmap words per chunk (512 rows)
    constexpr int kGatherMaxRows = CW * 32 / 4;         // gather when <= 25 % of the chunk's rows pass
    // per warp: store rows that pass and have not been processed yet -- this chunk's and up to R*G - 1 carried
    // over from earlier chunks: a step is only issued with all its slots filled (at 1 % selectivity a chunk holds
    // ~5 passing rows for 8 slots; processing chunk by chunk left 40 % of the loads in flight empty and paid 8
    // dependent steps per warp where 5 do)
    constexpr int RG = R * G;
    __shared__ uint32_t s_pass[kScanWarps][kGatherMaxRows + 32];
    int fill = 0;                                       // rows waiting in the list (uniform across the warp)
    const int64_t nchunk = (nblk + CW - 1) / CW;
    auto fetch_words = [&](int64_t chunk) -> uint32_t {
      const int64_t w = chunk * CW + lane;
      if (lane >= CW || chunk >= nchunk || w >= nblk) return 0u;
      const uint32_t f = (w < a.filter_words) ? __ldg(a.filter + w) : 0u;
      return f ? (f & __ldg(a.live + w)) : 0u;
    };
    int64_t chunk = static_cast<int64_t>(blockIdx.x) * kScanWarps + warp;
    uint32_t next_word = fetch_words(chunk);
    for (; chunk < nchunk; chunk += wstride) {
      const uint32_t word = next_word;
      next_word = fetch_words(chunk + wstride);
      const int pop = __popc(word);
      int incl = pop;
#pragma unroll
      for (int off = 1; off < CW; off <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
      }
      const int T = __shfl_sync(0xffffffffu, incl, CW - 1);   // lanes >= CW hold word = 0
      if (T == 0) continue;
      const uint4* cbase = vec + static_cast<size_t>(chunk) * (CW * kRowsPerBlock) * cpr;
      if (T > kGatherMaxRows) {
        // well-filled chunk: block by block as in the dense walk (cheaper row selection, rows
        // adjacent in memory), with the bitmap words already in registers
#pragma unroll 1
        for (int w = 0; w < CW; ++w) {
          uint32_t m = __shfl_sync(0xffffffffu, word, w);
          const uint4* bbase = cbase + static_cast<size_t>(w) * kRowsPerBlock * cpr;
          while (m) {
            int r[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
              r[i] = -1;
#pragma unroll
              for (int g = 0; g < G; ++g) {
                const int bit = m ? (__ffs(m) - 1) : -1;
                m &= (m - 1);
                if (G == 1 || g == grp) r[i] = bit;
              }
            }
            process(bbase, (chunk * CW + w) * kRowsPerBlock, r);
          }
        }
        continue;
      }
      // sparse chunk: every lane appends the offsets of its word's set bits to the warp's list (ranks
      // from the prefix scan), then each step fills all R x G row slots from the list with one
      // shared-memory read per slot.  (An earlier version located the j-th set bit per slot with
      // ballot + shuffle + __fns: ~70 instructions per slot, 2/3 of the kernel's issue slots at 10 %.)
      {
        uint32_t m = word;
        int pos = fill + incl - pop;
        const uint32_t row_of_word = static_cast<uint32_t>((chunk * CW + lane) * kRowsPerBlock);
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= (m - 1);
          s_pass[warp][pos++] = row_of_word + static_cast<uint32_t>(bit);
        }
      }
      fill += T;
      __syncwarp();
      int j0 = 0;
      for (; j0 + RG <= fill; j0 += RG) {              // full steps only
        int r[R];
#pragma unroll
        for (int i = 0; i < R; ++i) r[i] = static_cast<int>(s_pass[warp][j0 + i * G + grp]);   // rank -> this lane's slot i
        process(vec, 0, r);
      }
      const int left = fill - j0;                       // < RG <= 16 rows stay for the next chunk: move them to the front
      const uint32_t keep = (j0 > 0 && lane < left) ? s_pass[warp][j0 + lane] : 0u;
      __syncwarp();
      if (j0 > 0 && lane < left) s_pass[warp][lane] = keep;
      fill = left;
      __syncwarp();                                     // the list is appended to by the next chunk
    }
    if (fill > 0) {                                     // the last, partial step
      int r[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int j = i * G + grp;
        r[i] = (j < fill) ? static_cast<int>(s_pass[warp][j]) : -1;
      }
      process(vec, 0, r);
    }
  }

  // ---- programmatic dependent launch, part 2 -----------------------------------------------------
  // Everything above only READ (corpus, bitmaps, queries).  From here on the kernel writes global
  // scratch that is reused between queries (lists, tickets, outputs): wait until the previous
  // launch on the stream has completed.  A no-op for ordinary launches.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // ---- CTA merge: sort each query's 8 warp lists together, publish the first k -----------------
  __syncthreads();
  __shared__ int s_flag;
  __shared__ int s_total;
#pragma unroll 1
  for (int qb = 0; qb < QB; ++qb) {
    const int b = b0 + qb;
    if (b >= nB) break;
    uint64_t* region = lists + static_cast<size_t>(qb) * kScanWarps * kpad;
    block_bitonic_sort(region, kScanWarps * kpad);
    uint64_t* out = a.partial + (static_cast<size_t>(blockIdx.x) * a.B + b) * k;
    for (int j = threadIdx.x; j < k; j += blockDim.x) out[j] = region[j];
  }
  // ---- ticket: the last CTA of this query group merges every CTA's list and emits -------------
  if (a.done == nullptr) return;      // unfused mode: a separate merge kernel follows
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(a.done + blockIdx.y, 1u);
    s_flag = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  if (threadIdx.x == 0) a.done[blockIdx.y] = 0u;    // ticket is clean again for the next launch

  const int S = gridDim.x;
  const int total = S * k;
  const int ko = a.k_out;                                      // hits emitted per query (k_out <= k)
  uint64_t* pool = reinterpret_cast<uint64_t*>(smem_raw);      // queries / lists are dead now
#pragma unroll 1
  for (int qb = 0; qb < QB; ++qb) {
    const int b = b0 + qb;
    if (b >= nB) break;
    uint64_t* sorted;
    __syncthreads();
    if (k <= 64 && S <= 2 * static_cast<int>(blockDim.x)) {
      // k-way merge by k rounds of block-wide min over the heads of the S sorted lists.
      // Thread t owns lists t and t + 256; each keeps its current head and the next key in
      // registers, so a round costs two barriers (~0.1 us) and no memory latency unless the
      // same list wins twice in a row.  ~1-2 us for 296 lists x k = 10 (a bitonic sort of the
      // same candidates needs 55 barrier-separated stages per 1024 keys).
      uint64_t cur[2], nxt[2];
      int pos[2];
      const uint64_t* lst[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int s = threadIdx.x + u * static_cast<int>(blockDim.x);
        lst[u] = a.partial + (static_cast<size_t>(s < S ? s : 0) * a.B + b) * k;
        pos[u] = 0;
        cur[u] = (s < S) ? __ldcg(lst[u]) : kEmptyKey;
        nxt[u] = (s < S && k > 1) ? __ldcg(lst[u] + 1) : kEmptyKey;
      }
      uint64_t* wmin = pool;                 // [8] per-warp minima
      uint64_t* res = pool + 8;              // [k] result
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
        for (int i = 1; i < kScanWarps; ++i) w = wmin[i] < w ? wmin[i] : w;
        if (threadIdx.x == 0) res[r] = w;
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
      sorted = res;
    } else if (2 * kpad <= a.merge_keys_cap) {
      // rounds of: keep the best k so far in pool[0..k), append the next cap-k candidates,
      // sort the pool.  The pool is kept small (8 KB) on purpose: a larger dynamic
      // shared-memory request shrinks L1 for every CTA and costs the scan ~4% of HBM bandwidth.
      const int n = a.merge_keys_cap;
      const int take = n - k;
      for (int i = threadIdx.x; i < k; i += blockDim.x) pool[i] = kEmptyKey;
      for (int pos = 0; pos < total; pos += take) {
        uint64_t tmp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {        // cap <= 1024 keys: at most 4 candidates per thread per round
          const int i = threadIdx.x + u * static_cast<int>(blockDim.x);
          const int c = pos + i;
          tmp[u] = kEmptyKey;
          if (i < take && c < total) {
            const int s = c / k, j = c - s * k;
            tmp[u] = __ldcg(a.partial + (static_cast<size_t>(s) * a.B + b) * k + j);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = threadIdx.x + u * static_cast<int>(blockDim.x);
          if (i < take) pool[k + i] = tmp[u];
        }
        __syncthreads();
        block_bitonic_sort(pool, n);
      }
      sorted = pool;
    } else {
      // large k: per-warp running lists with early exit over the (sorted) source lists
      for (int i = threadIdx.x; i < kScanWarps * kpad; i += blockDim.x) pool[i] = kEmptyKey;
      __syncthreads();
      uint64_t* L = pool + static_cast<size_t>(warp) * kpad;
      for (int s = warp; s < S; s += kScanWarps) {
        const uint64_t* src = a.partial + (static_cast<size_t>(s) * a.B + b) * k;
        bool fin = false;
        for (int j0 = 0; j0 < k && !fin; j0 += 32) {
          const int j = j0 + lane;
          const uint64_t cand = (j < k) ? __ldcg(src + j) : kEmptyKey;
          uint32_t bal = __ballot_sync(0xffffffffu, cand < L[k - 1]);
          if (bal != 0xffffffffu) fin = true;
          while (bal) {
            const int sl = __ffs(bal) - 1;
            bal &= (bal - 1);
            const uint64_t ck = shfl_u64(cand, sl);
            if (ck < L[k - 1]) warp_list_insert(L, k, ck, lane);
            else bal = 0;
          }
        }
      }
      __syncthreads();
      block_bitonic_sort(pool, kScanWarps * kpad);
      sorted = pool;
    }
    if (a.exact != nullptr) {
      // ---- exact fp32 re-ranking (bf16 store with an fp32 plane) ---------------------------------
      // `sorted` holds the k best rows by STORED-precision distance.  Re-score them against the
      // un-rounded fp32 rows with the un-rounded (normalised) query -- one warp per candidate; slot j
      // is read and rewritten only by warp j % 8 -- then sort again; the best k_out are emitted.
      __syncthreads();
      const int bs_q = a.q_index ? a.q_index[b] : b;
      const float qscale = s_scale[qb];
      for (int j = warp; j < k; j += kScanWarps) {
        const uint64_t key = sorted[j];
        if (key == kEmptyKey) continue;
        const uint32_t row = key_row(key);
        const float* xr = a.exact + static_cast<size_t>(row) * a.exact_elems;
        float acc = 0.0f;
        for (int e = lane; e < a.dim; e += 32) {
          const float qv = raw_q(bs_q, e) * qscale;
          const float xv = __ldg(xr + e);
          if constexpr (L2) { const float d = xv - qv; acc = fmaf(d, d, acc); }
          else acc = fmaf(xv, qv, acc);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) sorted[j] = make_key(L2 ? acc : 1.0f - acc, row);
      }
      __syncthreads();
      for (int j = k + threadIdx.x; j < kpad; j += blockDim.x) sorted[j] = kEmptyKey;
      block_bitonic_sort(sorted, kpad);
    }
    if (a.xchg_peers != nullptr) {
      // publish this shard's list of query b into the slot [parity][my rank] of EVERY rank's buffer
      const int par = static_cast<int>(a.xchg_epoch & 1u);
      for (int idx = threadIdx.x; idx < a.xchg_world * ko; idx += blockDim.x) {
        const int g = idx / ko, j = idx - g * ko;
        uint64_t key = sorted[j];
        if (key != kEmptyKey) key = key_to_global(a.rows_map, key);
        uint64_t* dst = reinterpret_cast<uint64_t*>(a.xchg_peers[g] + kXchgKeysOff) +
                        (static_cast<size_t>(par) * a.xchg_world + a.xchg_rank) * a.xchg_slot_keys +
                        static_cast<size_t>(b) * ko + j;
        st_relaxed_sys_u64(dst, key);
      }
      continue;
    }
    int cnt = 0;
    const int bs = a.q_index ? a.q_index[b] : b;      // where this query's result belongs
    for (int j = threadIdx.x; j < ko; j += blockDim.x) {
      uint64_t key = sorted[j];
      const bool valid = key != kEmptyKey;
      if (valid) {
        key = key_to_global(a.rows_map, key);
        cnt++;
      }
      const size_t o = static_cast<size_t>(bs) * ko + j;
      if (a.out_keys) a.out_keys[o] = key;
      if (a.out_rows) a.out_rows[o] = valid ? static_cast<int64_t>(key_row(key)) : -1;
      if (a.out_dists) a.out_dists[o] = valid ? key_dist(key) : __int_as_float(0x7f800000);
    }
    if (a.out_counts) {           // block-wide sum of the valid entries
      if (threadIdx.x == 0) s_total = 0;
      __syncthreads();
      if (cnt) atomicAdd(&s_total, cnt);
      __syncthreads();
      if (threadIdx.x == 0) a.out_counts[bs] = s_total;
    }
  }
  if (a.xchg_peers == nullptr) {
    if (a.done_flag != nullptr) {         // results went to host memory: make them visible, then raise the flag
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) st_release_sys_u32(a.done_flag, a.done_seq);
    }
    return;
  }

  // ---- fused all-gather + cross-shard merge over peer memory ----------------------------------
  // Every thread's stores above are made visible system-wide before ONE flag per peer is raised;
  // a peer that sees flag >= epoch (acquire) therefore sees the keys.  Two buffer halves (epoch
  // parity) are enough: a rank can only start epoch e+2 after every rank delivered e+1, i.e.
  // after every rank's kernel of epoch e (which read half e&1) has finished.
  {
    const int par = static_cast<int>(a.xchg_epoch & 1u);
    const int G = a.xchg_world;
    const size_t flag_idx = (static_cast<size_t>(par) * kXchgMaxGroups + blockIdx.y) * kXchgMaxWorld;
    __threadfence_system();
    if (threadIdx.x == 0) s_flag = 0;                 // reused: 1 = a peer never delivered
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < G) {
      uint32_t* theirs = reinterpret_cast<uint32_t*>(a.xchg_peers[threadIdx.x]) + flag_idx + a.xchg_rank;
      st_release_sys_u32(theirs, a.xchg_epoch);
      const uint32_t* mine = reinterpret_cast<const uint32_t*>(a.xchg_peers[a.xchg_rank]) + flag_idx + threadIdx.x;
      const unsigned long long t0 = global_timer_ns();
      while (static_cast<int32_t>(ld_acquire_sys_u32(mine) - a.xchg_epoch) < 0) {
        if (global_timer_ns() - t0 > 20000000000ull) {    // 20 s: a peer never launched; flag it, do not hang the GPU
          atomicOr(reinterpret_cast<unsigned int*>(a.xchg_peers[a.xchg_rank] + kXchgStatusOff), 1u + threadIdx.x);
          s_flag = 1;
          break;
        }
      }
    }
    __syncthreads();
    const uint64_t* slots = reinterpret_cast<const uint64_t*>(a.xchg_peers[a.xchg_rank] + kXchgKeysOff) +
                            static_cast<size_t>(par) * G * a.xchg_slot_keys;
    // one warp per query: lane g walks shard g's ascending list; k rounds of warp-min
    for (int qb = warp; qb < QB; qb += kScanWarps) {
      const int b = b0 + qb;
      if (b >= nB) break;
      const uint64_t* lst = slots + static_cast<size_t>(lane < G ? lane : 0) * a.xchg_slot_keys + static_cast<size_t>(b) * ko;
      int pos = 0;
      uint64_t cur = (lane < G) ? ld_relaxed_sys_u64(lst) : kEmptyKey;
      int cnt = 0;
      for (int r = 0; r < ko; ++r) {
        uint64_t m = cur;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const uint64_t o = shfl_u64(m, lane ^ off);
          m = o < m ? o : m;
        }
        const bool valid = m != kEmptyKey;
        if (lane == 0) {
          const size_t o = static_cast<size_t>(b) * ko + r;
          if (a.out_keys) a.out_keys[o] = m;
          if (a.out_rows) a.out_rows[o] = valid ? static_cast<int64_t>(key_row(m)) : -1;
          if (a.out_dists) a.out_dists[o] = valid ? key_dist(m) : __int_as_float(0x7f800000);
        }
        if (valid) {
          cnt++;
          if (cur == m) {                        // global rows are unique: exactly one lane advances
            pos++;
            cur = (pos < ko) ? ld_relaxed_sys_u64(lst + pos) : kEmptyKey;
          }
        }
      }
      // a count of -1 tells the host that this result is built from a stale slot (a peer timed out)
      if (lane == 0 && a.out_counts) a.out_counts[b] = s_flag ? -1 : cnt;
    }
    if (a.done_flag != nullptr) {
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) st_release_sys_u32(a.done_flag, a.done_seq);
    }
  }
}

size_t scan_smem_bytes(int QB, int row_elems, int k) {
  return static_cast<size_t>(QB) * row_elems * sizeof(float) +
         static_cast<size_t>(QB) * kScanWarps * next_pow2(k) * sizeof(uint64_t);
}

// dynamic shared memory: the scan's own needs, or 8 KB if that is more (pool of the final merge)
size_t scan_total_smem(int QB, int row_elems, int k, int grid_x, int* merge_cap) {
  (void)grid_x;
  size_t need = scan_smem_bytes(QB, row_elems, k);
  if (need < 8 * 1024) need = 8 * 1024;
  int cap = 1;
  while (static_cast<size_t>(cap) * 2 * sizeof(uint64_t) <= need && cap < 1024) cap <<= 1;
  *merge_cap = cap;
  return need;
}

template <typename Kern>
cudaError_t launch_pdl(Kern kern, const ScanArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  static const bool pdl = !(getenv("RAG_B200_PDL") && atoi(getenv("RAG_B200_PDL")) == 0);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kScanThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

template <bool BF16, int QB, int NJ, int R, int LPR>
cudaError_t launch_one(const ScanArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  if (a.l2) return launch_pdl(scan_stream_kernel<BF16, QB, NJ, R, true, LPR>, a, grid, smem, st);
  return launch_pdl(scan_stream_kernel<BF16, QB, NJ, R, false, LPR>, a, grid, smem, st);
}

template <bool BF16, int QB>
cudaError_t launch_qb(const ScanArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  // row slots in flight per warp chosen so that every lane has ~12 independent 16-byte loads
  // outstanding and R*QB fits the lanes that share a row
  constexpr int R4 = (QB <= 8) ? 4 : 2;
  constexpr int R16 = (QB <= 4) ? 4 : 2;          // 16 lanes per row: R*QB <= 16
  constexpr int R8 = (QB <= 2) ? 4 : (QB <= 4 ? 2 : 1);
  if (a.cpr == 96) return launch_one<BF16, QB, 3, R4, 32>(a, grid, smem, st);    // 768-d bf16, 384-d fp32
  if (a.cpr == 192) return launch_one<BF16, QB, 6, 2, 32>(a, grid, smem, st);    // 1536-d bf16, 768-d fp32
  if (a.cpr == 64) return launch_one<BF16, QB, 2, R4, 32>(a, grid, smem, st);    // 512-d bf16, 256-d fp32
  if (a.cpr == 128) return launch_one<BF16, QB, 4, 2, 32>(a, grid, smem, st);    // 1024-d bf16, 512-d fp32
  if (a.cpr == 48) return launch_one<BF16, QB, 3, R16, 16>(a, grid, smem, st);   // 384-d bf16, 192-d fp32
  if (a.cpr == 24) return launch_one<BF16, QB, 3, R8, 8>(a, grid, smem, st);     // 192-d bf16, 96-d fp32
  return launch_one<BF16, QB, 0, R4, 32>(a, grid, smem, st);
}

}  // namespace

int scan_stream_grid_x(int sm_count, int64_t n_rows) {
  int64_t nblk = (n_rows + kRowsPerBlock - 1) / kRowsPerBlock;
  int64_t want = (nblk + kScanWarps - 1) / kScanWarps;   // CTAs that would each get >= 1 block per warp
  int64_t g = 2LL * sm_count;                              // 2 resident CTAs per SM, persistent
  if (want < g) g = want;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

int scan_stream_groups(int B, int dtype, int row_elems, int k) {
  const int max_qb = scan_stream_max_qb(dtype, row_elems, k);
  int QB = 1;
  while (QB < B && QB < max_qb) QB <<= 1;
  return (B + QB - 1) / QB;
}

int scan_stream_max_qb(int dtype, int row_elems, int k) {
  // shared memory per CTA must allow 2 CTAs/SM: keep it under ~100 KB
  const size_t budget = 100 * 1024;
  int qb = 8;
  while (qb > 1 && scan_smem_bytes(qb, row_elems, k) > budget) qb >>= 1;
  (void)dtype;
  return qb;
}

cudaError_t launch_scan_stream(const ScanArgs& a, int sm_count, cudaStream_t st, int* launches) {
  if (a.B <= 0 || a.k <= 0) return cudaErrorInvalidValue;
  const int max_qb = scan_stream_max_qb(a.dtype, a.row_elems, a.k);
  int QB = 1;
  while (QB < a.B && QB < max_qb) QB <<= 1;
  if (scan_smem_bytes(QB, a.row_elems, a.k) > 200 * 1024) return cudaErrorInvalidValue;
  dim3 grid(a.grid_x, (a.B + QB - 1) / QB, 1);
  ScanArgs aa = a;
  const size_t smem = scan_total_smem(QB, a.row_elems, a.k, a.grid_x, &aa.merge_keys_cap);
  const bool bf16 = (a.dtype == 1);
  cudaError_t e = cudaErrorInvalidValue;
  switch (QB) {
    case 1: e = bf16 ? launch_qb<true, 1>(aa, grid, smem, st) : launch_qb<false, 1>(aa, grid, smem, st); break;
    case 2: e = bf16 ? launch_qb<true, 2>(aa, grid, smem, st) : launch_qb<false, 2>(aa, grid, smem, st); break;
    case 4: e = bf16 ? launch_qb<true, 4>(aa, grid, smem, st) : launch_qb<false, 4>(aa, grid, smem, st); break;
    case 8: e = bf16 ? launch_qb<true, 8>(aa, grid, smem, st) : launch_qb<false, 8>(aa, grid, smem, st); break;
    default: break;
  }
  if (launches) *launches += 1;
  return e;
}

}  // namespace rag
