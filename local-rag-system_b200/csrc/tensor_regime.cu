// K3 + K4: tensor-core search regime for large query batches.
//
// Replaces, for B > 2 queries (B > 4 on fp32 stores), the hnswlib graph walk / numpy brute force
// behind collection.query (api/app.py:544-549) with an exact dense contraction
// Q[B x D] . X[N x D]^T on the 5th-generation tensor cores, with the top-k selection fused into
// the epilogue so the B x N score matrix never exists.
//
// Design (sm_100a only; tcgen05 + TMEM + TMA; measurements and the reasons in DESIGN.md 3.2 / 3.3):
//   * one persistent CTA per SM, dedicated for its whole life to ONE tile of 128 queries (the MMA M
//     dimension).  The 128 x D bf16 query tile is written ONCE into TENSOR MEMORY (D/2 columns) and
//     used as the A operand from TMEM (tcgen05.mma ... [d_tmem], [a_tmem], b_desc): the queries never
//     touch shared memory or L2 again, so on-chip traffic is the corpus stream only.
//   * the corpus streams HBM -> TMA (128B-swizzled boxes of 64 elements x 32/64 rows) -> a 224 KB
//     shared-memory ring -> tcgen05.mma as the B operand (64 rows per tile).
//   * with two query tiles or more, x-adjacent CTAs form cta_group::2 PAIRS: each loads only its 32
//     rows of a tile, the leader issues one M = 256 MMA per k-step for both tensor cores (MODE 2;
//     MODE 1 is the older TMA-multicast pair, MODE 0 a single CTA).
//   * accumulators: a ring of 64-column fp32 buffers at the top of TMEM (2 behind a 768-wide A
//     operand, 4 when D <= 512).  Four epilogue warps drain a tile with tcgen05.ld (thread = one
//     query, 64 scores), hand the buffer back, reject the whole tile with a 32-instruction
//     three-input-max test against a bound shared by all CTAs of the query tile, and only otherwise
//     walk the scores and insert survivors into a per-thread list (registers for k <= 16, a max-heap
//     in shared memory up to k = 128, in local memory beyond).  live/filter bitmaps are applied on the
//     survivor path; tiles whose 64 rows are all dead or filtered are skipped by every role.
//   * the bound: the thread's own k-th best, the best k-th best of the sibling CTAs of the query
//     tile, and a bound built from ALL of them (k <= 16: group bound; larger k: quantile bound).
//   * fp32 stores are contracted through a bf16 shadow of their rows -- bf16(x) alone (FILT: every row
//     within the rounding bound of the running k-th best is buffered and re-scored exactly afterwards)
//     or hi/lo pairs (3 MMAs per k-step, k + slack candidates re-ranked + an a-posteriori guard).
//   * warp roles: 0 = TMA producer, 1 and 6 = MMA issuers on alternate tiles (1 owns the TMEM
//     allocation; ONE issuer where plan_ring() says the ring is too short for two), 2-5 = epilogue.
//
// Roofline: HBM for B <= ~256 (corpus read once), tensor pipe beyond.  Algorithmic FLOPs = 2 B N D.
#include "tensor_regime.h"

#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace rag {
namespace tensor {

namespace {

constexpr int kM = 128;              // queries per CTA (MMA M)
constexpr int kNB = 64;              // corpus rows per tile (MMA N); kNBWide where tensor memory has room (see below)
constexpr int kNBWide = 128;
constexpr int kAtomK = 64;           // bf16 elements per 128-byte swizzle atom row
constexpr int kAtomsPerStage = 4;
constexpr int kStageK = kAtomK * kAtomsPerStage;      // 256 K elements per pipeline stage
// shared-memory ring geometry.  MODE 0 (one CTA) and 1 (multicast pair): a stage holds all 64 rows of
// a tile (4 atoms x 8 KB = 32 KB, 7 stages).  MODE 2 (cta_group::2 pair): each CTA of the pair holds
// only ITS 32 rows of every tile (4 atoms x 4 KB = 16 KB, 14 stages); the pair's tensor cores share
// the two halves, which halves the shared-memory traffic per SM.
// NB = 128 (pairs only, A operand <= 256 columns, i.e. D <= 512): every per-tile and per-instruction cost -- the
// issuing thread's ~13 scalar instructions per MMA, barrier hand-offs, the epilogue's fixed part -- is paid per
// 128 rows instead of 64.  At D = 384 a 64-row tile is only 768 cycles of tensor work, LESS than what the two
// issuing threads and the epilogue warps need per tile; the wide tile doubles the budget.
template <int MODE, int NB = kNB> struct Ring {
  static constexpr int kRows = (MODE == 2) ? NB / 2 : NB;
  static constexpr int kAtomBytes = kRows * kAtomK * 2;
  static constexpr int kStageBytes = kAtomsPerStage * kAtomBytes;
  static constexpr int kStages = (7 * kAtomsPerStage * kNB * kAtomK * 2) / kStageBytes;      // 224 KB in flight per SM either way
};
constexpr int kRingBytes = 7 * kAtomsPerStage * kNB * kAtomK * 2;
constexpr int kMmasPerStage = kStageK / 16;            // 16
constexpr int kThreads = 224;            // TMA producer, MMA issuer A, 4 epilogue warps, MMA issuer B
constexpr int kMmaWarpB = 6;
constexpr int kEpiWarp0 = 2;
constexpr int kTmemCols = 512;
constexpr int kMaxKCols = 384;       // A operand: up to 768 bf16 per query
constexpr int kMaxAccBufs = 4;
constexpr int kSmemBytes = kRingBytes + 1024 /*align*/ + 512 /*barriers*/;

// ---- PTX wrappers -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// exactly one lane of the (converged) warp gets true.  Using elect.sync -- rather than
// `lane == 0` -- lets ptxas keep the TMA / MMA operands in uniform registers; with a
// plain lane test it wraps every UTCHMMA / UTMALDG in a per-lane "waterfall" loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(dst), "l"(map), "r"(bar), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T ; A is 128 x 16 bf16 (8 columns), B is kNB x 16 bf16 K-major
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
// ---- cta_group::2 (CTA pair) variants ------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
// arrive on a barrier of another CTA of the cluster.  Default semantics (release at CTA scope) on purpose:
// what the waiter depends on are tensor-memory accesses already completed by tcgen05.wait + fenced by
// tcgen05.fence::before_thread_sync; a .release.cluster arrive costs MEMBAR.ALL.GPU + ERRBAR per tile
// (measured: 30 % of all stall samples of the pair kernel).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// load into THIS CTA's shared memory, complete_tx on a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem, both CTAs] (+)= A[tmem, 128 rows per CTA] . B[smem, kNB/2 rows per CTA]^T, issued by the leader only
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);        // start address
  d |= static_cast<uint64_t>(1) << 16;                           // leading byte offset (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // stride byte offset: next 8-row group
  d |= static_cast<uint64_t>(1) << 46;                           // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                           // SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M = 128, N = kNB
constexpr uint32_t make_idesc(int n, int m) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// the pair's instruction has M = 256 (128 rows in each CTA's tensor memory)

// selection statistics of the epilogue (RAG_B200_TENSOR_STATS=1; read with rag_debug_tensor_stats):
// [0] tiles drained per epilogue warp, [1] tiles that passed the first reject test, [2] ... the second (l2),
// [3] candidate scores examined, [4] list insertions, [5] quantile-list updates
__device__ unsigned long long g_tensor_stats[8];

struct Args {
  int lists_in_smem;             // 16 < k <= 128: the per-thread heaps live in shared memory behind the barriers
  uint32_t q_early, q_late_mask; // quantile bound: refreshed every 8th tile for the first q_early tiles, then when (it & mask) == 1
  int stats;                     // 1: count into g_tensor_stats (debug; costs a few atomics per slow-path tile)
  const __nv_bfloat16* q_bf16;   // [n_mtiles*128][row_elems] prepared queries, zero rows beyond B
  const float* q_norm2;          // [n_mtiles*128]
  const float* x_norm2;          // [n_rows] (l2)
  const float* x_min_norm2;      // [1] lower bound of x_norm2 over the store (l2)
  const uint32_t* live;
  const uint32_t* filter;
  int64_t filter_words;
  int64_t n_rows;
  int row_elems;
  int B, k;
  int cpm;                       // CTAs per query tile
  int dense;                     // 1: no filter and no tombstones -> tile masks are computed, not loaded
  int prefetch;                  // L2 prefetch distance in tiles (0 = off)
  int nbuf;                      // accumulator buffers in tensor memory (2 or 4)
  int stages_used;               // ring stages in use (<= Ring<MODE>::kStages)
  int one_issuer;                // 1: warp 1 issues every tile (the ring holds fewer than two whole tiles, see the MMA warps)
  int split_steps;               // > 0: split-precision rows [hi | lo], the first split_steps k-steps are hi
  uint64_t* partial;             // [cpm][B][k]
  // [n_mtiles*128] ordered(fp32): smallest k-th-best distance any CTA has reached for this query so far.
  // The cpm CTAs that share a query each see 1/cpm of the corpus; a row beyond ANY CTA's k-th best cannot
  // be in the global top-k, so publishing the bound lets every CTA reject with the tightest one
  // (an order of magnitude fewer list insertions at k = 100).  0xFFFFFFFF = nothing published yet.
  uint32_t* tau_shared;
  // Quantile bound for large k (nullptr = off): [n_mtiles*128][cpm] ordered(fp32).  CTA c publishes the
  // distance of its own q-th best row, q = ceil(k / cpm).  Every CTA then holds >= q rows at or below
  // T = max_c(slot c), i.e. >= k rows globally: nothing beyond T can be in the global top-k.  T estimates the
  // same quantile as the global k-th best, cpm times tighter than any single CTA's own k-th best above.
  uint32_t* tau_q;
  int tau_q_rank;                // q (1..8)
  // Group bound for k <= 16 (nullptr = off): [n_mtiles*128][grp_n] ordered(fp32), grp_n = min(cpm, k) groups of CTAs
  // (cj % grp_n).  Every CTA keeps atomicMin-ing its own r-th best (r = grp_rank = ceil(k / grp_n)) into its group's
  // slot: the CTA that set a slot holds r rows at or below it, the groups scan disjoint rows, so grp_n * r >= k rows
  // lie at or below T = max over the slots and nothing beyond T can be in the global top-k.  T tracks the k-th best of
  // everything ALL CTAs have seen -- `tau_shared` only the best single CTA's -- at grp_n <= 16 loads per refresh.
  // It is what lets small corpora (a few thousand rows per CTA) leave the warm-up phase at all: 1M rows over 148
  // CTAs had ~1 candidate per query and tile throughout, 7.6 k cycles of epilogue per 64-row tile.
  uint32_t* tau_grp;
  int grp_n, grp_rank;
  // [cpm][gridDim.y] inverted tile counters (0xFFFFFFFF - tiles issued) of the clusters that walk the same
  // corpus tiles for different query tiles.  A cluster may run at most kPaceWindow tiles ahead of its slowest
  // sibling, so a tile fetched from DRAM by the first cluster is still in L2 when the others ask for it
  // (without pacing the siblings drift apart and every tile is read from DRAM about twice).
  uint32_t* progress;
  uint32_t pace_window;          // tiles a cluster may run ahead of its slowest sibling
  // FILT (hi-only shadow of an fp32 store, k <= 16): the contraction ranks by bf16(q).bf16(x), which is within
  // eps = bf16_contraction_eps() of the exact distance.  Every row of the exact top-k then lies within 2 eps of the
  // approximate k-th best (see refine_filter_kernel), so next to its top-k list every epilogue thread keeps ALL rows
  // within `2 eps` of its running bound in a small buffer; the survivors are re-scored exactly afterwards.
  const float* q_lo_norm2;       // [B] |q - bf16(q)|^2
  const float* x_max_norm2;      // [1] max |x|^2 over the store
  const float* x_lo_max2;        // [1] max |x - bf16(x)|^2 over the store
  float guard_rel;
  uint64_t* extra;               // [cpm][B][kFiltCap] buffered keys (unordered)
  int* extra_cnt;                // [cpm][B] entries, -1 = the buffer overflowed (the query is re-run exactly)
};
// buffer entries per (CTA, query); 128 threads x 64 x 8 B = 64 KB of shared memory.  A thread ends with k rows + those
// in the 2 eps band above its bound (1M x 768 unit-norm rows over 148 CTAs: ~20 +- 4); 32 entries overflowed for
// 1.5 % of the queries (any of the 148 CTAs of a query), and an overflow costs an exact re-run
constexpr int kFiltCap = 64;
constexpr size_t kPaceBytes = 56u << 20;  // L2 budget (of 126 MB) for the tiles between the slowest and the fastest sibling

// per-thread running top-k.
//   KL <= 16        sorted list, fully unrolled, in registers;
//   16 < KL <= 128  a binary MAX-heap (root = current k-th best; an insertion is a root replacement + sift-down,
//                   O(log k) instead of the O(k) shift of a sorted list) in SHARED memory behind the barriers,
//                   entry i of epilogue thread t at word i * 128 + t (the 32 lanes of a warp touch consecutive
//                   words).  In local memory the same heap cost 4 % more (25M x 384, k = 100: 24.1 -> 23.1 ms);
//                   falls back to local memory for the rare shapes whose block does not fit next to the ring (one
//                   query tile, D = 768, k > 96).  A warp-cooperative SORTED list (all 32 lanes insert one
//                   candidate together, candidates of different lanes one after the other) was measured too and
//                   is slower (25.1 ms): a tile's ~7 candidates sit in different lanes, whose private sift-downs
//                   run side by side.
//   KL  > 128       the same heap in local memory (1024 x 8 B x 128 threads does not fit on chip).
constexpr int kEpiThreads = 128;
constexpr int kSmemListMax = 128;
template <int KL>
struct TopList {
  static constexpr bool kRegs = (KL <= 16);
  static constexpr bool kShared = (KL > 16 && KL <= kSmemListMax);
  uint64_t e[KL];                                      // (kShared: only used when the block did not fit, see launch_one)
  uint64_t* sh;                                        // kShared: this thread's column of the CTA's list block, or nullptr
  __device__ __forceinline__ uint64_t& at(int i) {
    if constexpr (kShared) return sh != nullptr ? sh[i * kEpiThreads] : e[i];
    else return e[i];
  }
  __device__ __forceinline__ void init(uint64_t* block, int tid, int k) {
    if constexpr (kShared) {
      sh = block != nullptr ? block + tid : nullptr;
      for (int i = 0; i < k; ++i) at(i) = kEmptyKey;
    } else {
      sh = nullptr;
#pragma unroll
      for (int i = 0; i < KL; ++i) e[i] = kEmptyKey;
    }
  }
  __device__ __forceinline__ uint64_t kth(int k) {
    if constexpr (kRegs) {
      uint64_t t = e[KL - 1];
#pragma unroll
      for (int i = 0; i < KL; ++i) t = (i == k - 1) ? e[i] : t;
      return t;
    } else {
      return at(0);
    }
  }
  // precondition: key < kth(k)
  __device__ __forceinline__ void insert(uint64_t key, int k) {
    if constexpr (kRegs) {
#pragma unroll
      for (int i = 0; i < KL; ++i) {
        const uint64_t cur = e[i];
        const bool lt = key < cur;
        e[i] = lt ? key : cur;
        key = lt ? cur : key;
      }
    } else {
      sift(key, k);
    }
  }
  __device__ __forceinline__ void sift(uint64_t key, int n) {      // sift `key` down from the root of a heap of n
    int i = 0;
    for (;;) {
      const int l = 2 * i + 1, r = l + 1;
      if (l >= n) break;
      const uint64_t lv = at(l);
      const uint64_t rv = (r < n) ? at(r) : 0ull;
      const int c = (rv > lv) ? r : l;
      const uint64_t cv = (rv > lv) ? rv : lv;
      if (cv <= key) break;
      at(i) = cv;
      i = c;
    }
    at(i) = key;
  }
  // ascending order in slots 0..k) (heap sort for the heap variants; the register list already is)
  __device__ __forceinline__ void finalize(int k) {
    if constexpr (!kRegs) {
      for (int n = k - 1; n > 0; --n) {
        const uint64_t key = at(n);    // move the max to its final slot, re-insert the displaced key
        at(n) = at(0);
        sift(key, n);
      }
    }
  }
};

// CL = CTAs per cluster.  The CL CTAs of a cluster serve CL different query tiles and walk
// the same corpus tiles; each loads 1/CL of every tile and TMA-multicasts it to all of
// them, so L2 -> SM traffic per CTA drops by CL (L2 bandwidth ~ HBM bandwidth on this chip,
// and it is what bounds the unicast version at ~0.9 PFLOP/s).
// MODE 2 -- cta_group::2: the same two CTAs form an MMA pair instead.  Each CTA loads ITS 32 rows of a
// tile into its own shared memory only; the leader (cluster rank 0) issues one M = 256 MMA per k-step
// for both (each tensor core contracts its own 128 queries with the 64 rows held by the pair), so a
// corpus byte is written to and read from shared memory once per PAIR -- half the per-SM traffic of
// the multicast scheme (which at 64 B/clk in + 64 B/clk out sat at the SM's shared-memory limit).
template <int KL, bool L2, int MODE, int NB, bool FILT>
__global__ void __launch_bounds__(kThreads, 1) gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap, const Args a) {
  static_assert(!FILT || KL <= 16, "the filter buffers sit where the shared-memory heaps of larger k would");
  constexpr int CL = (MODE == 0) ? 1 : 2;
  constexpr bool PAIR = (MODE == 2);
  constexpr int kStages = Ring<MODE, NB>::kStages;
  constexpr int kStageBytes = Ring<MODE, NB>::kStageBytes;
  constexpr int kAtomBytes = Ring<MODE, NB>::kAtomBytes;
  constexpr int kAccCols = NB;                              // fp32 accumulator columns per buffer
  constexpr int NW = NB / 32;                               // bitmap words per tile
  constexpr int NSUB = NB / 64;                             // 64-column halves the epilogue drains one after the other
  static_assert(NB == 64 || (NB == 128 && MODE == 2), "the wide tile exists for CTA pairs only");
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;             // 1024-byte aligned stage ring
  // ring stages in use: all of them, or a short ring when the per-thread top-k lists live in local memory
  // (k > 16): 128 threads x 1 KB of lists must stay L1-resident, and a 227 KB carve-out leaves L1 ~28 KB
  // (the depth of the ring does not limit the kernel: 4 stages measured as fast as 14)
  const uint32_t n_ring = static_cast<uint32_t>(a.stages_used);
  const uint32_t bar_base = base + n_ring * kStageBytes;
  // barrier layout (8 bytes each): full[kStages], empty[kStages], acc_full[2], acc_empty[2], tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto accf_bar = [&](int b) { return bar_base + 8u * (2 * kStages + b); };
  auto acce_bar = [&](int b) { return bar_base + 8u * (2 * kStages + kMaxAccBufs + b); };
  const uint32_t aready_bar = bar_base + 8u * (2 * kStages + 2 * kMaxAccBufs);     // PAIR: both CTAs' queries are in tensor memory
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 2 * kMaxAccBufs + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // clusters of CL CTAs lie along x (a cta_group::2 pair must): x = cj * CL + rank, y = group of CL query tiles
  const int mt = static_cast<int>(blockIdx.y) * CL + static_cast<int>(blockIdx.x % CL);   // query tile
  const int cj = static_cast<int>(blockIdx.x / CL);   // position among the CTAs of this query tile
  const int64_t n_tiles = (a.n_rows + NB - 1) / NB;
  const int kcols = ((a.row_elems + 15) / 16) * 8;         // TMEM columns of the A operand
  const int k_steps = (a.row_elems + 15) / 16;             // MMAs per tile
  const int n_stages_per_tile = (a.row_elems + kStageK - 1) / kStageK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), PAIR ? 1 : CL); }   // (unused ones included)
    for (int b = 0; b < kMaxAccBufs; ++b) { mbar_init(accf_bar(b), 1); mbar_init(acce_bar(b), PAIR ? 8 : 4); }   // one arrival per epilogue warp
    mbar_init(aready_bar, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) { tmem_alloc_pair(tmem_slot, kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // peers' barriers exist before anyone multicasts into them
  tc_fence_after();
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  constexpr uint16_t kMcMask = static_cast<uint16_t>((1u << CL) - 1u);
  const uint32_t tmem_base = *tmem_slot_ptr;
  // accumulator ring at the top of tensor memory: 2 x 64 columns behind a 768-wide A operand, 4 x 64 when
  // the queries leave room (D <= 512).  With two buffers the hand-off chain  MMA done -> commit -> epilogue
  // tcgen05.ld -> arrive -> next MMA  has to fit inside ONE tile's MMA time (768 cycles at D = 384), and it
  // does not -- least of all across a CTA pair; four buffers give it three tiles.
  const int nbuf = a.nbuf;
  const int nbuf_log2 = (nbuf == 4) ? 2 : 1;
  const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(kTmemCols - nbuf * kAccCols);

  // Walks this CTA's corpus tiles (cj, cj + cpm, ...) and yields the ones with at least one
  // live & filter-passing row.  Mask words of the NEXT tile are requested one step ahead so
  // that no role ever stalls on a dependent global load; dense stores compute them instead.
  struct TileWalker {
    const Args& a;
    int64_t t, n_tiles;
    uint32_t p[NW];       // prefetched mask words of tile t
    bool masks;           // false: the caller only needs the tile sequence (dense stores: nothing to compute)
    __device__ __forceinline__ void fetch() {
      if (t >= n_tiles) {
#pragma unroll
        for (int i = 0; i < NW; ++i) p[i] = 0u;
        return;
      }
      if (a.dense && !masks) {
#pragma unroll
        for (int i = 0; i < NW; ++i) p[i] = (i == 0) ? 1u : 0u;
        return;
      }
      if (a.dense) {
        const int64_t left = a.n_rows - t * NB;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          const int64_t li = left - 32 * i;
          p[i] = li >= 32 ? 0xffffffffu : (li > 0 ? ((1u << li) - 1u) : 0u);
        }
        return;
      }
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const int64_t w = static_cast<int64_t>(NW) * t + i;
        p[i] = __ldg(a.live + w);
        if (a.filter != nullptr) p[i] &= (w < a.filter_words) ? __ldg(a.filter + w) : 0u;
      }
    }
    __device__ __forceinline__ TileWalker(const Args& a_, int64_t t0, int64_t n, bool masks_ = true)
        : a(a_), t(t0), n_tiles(n), masks(masks_) { fetch(); }
    __device__ __forceinline__ bool next(int64_t& tile, uint32_t* w) {
      while (t < n_tiles) {
        tile = t;
        uint32_t any = 0u;
#pragma unroll
        for (int i = 0; i < NW; ++i) { w[i] = p[i]; any |= p[i]; }
        t += a.cpm;
        fetch();
        if (any != 0u) return true;
      }
      return false;
    }
  };

  if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
    // ================= epilogue warps: load A into TMEM, then drain accumulators =================
    const int lg = warp & 3;                       // TMEM lane group this warp may access
    const int m = lg * 32 + lane;                  // query row inside the tile
    const int b = mt * kM + m;                     // global query index
    const bool q_valid = b < a.B;
    {
      const uint4* qrow = reinterpret_cast<const uint4*>(a.q_bf16 + static_cast<size_t>(mt * kM + m) * a.row_elems);
      const int chunks = a.row_elems / 8;          // 16-byte chunks in the row (row_elems % 8 == 0)
      for (int c0 = 0; c0 < kcols; c0 += 8) {      // 8 TMEM columns = 16 bf16 = two 16-byte chunks
        uint32_t v[8];
        const int ch = c0 / 4;
        uint4 lo = (ch < chunks) ? __ldg(qrow + ch) : make_uint4(0, 0, 0, 0);
        uint4 hi = (ch + 1 < chunks) ? __ldg(qrow + ch + 1) : make_uint4(0, 0, 0, 0);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
        v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        tmem_st_x8(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + static_cast<uint32_t>(c0), v);
      }
      tmem_wait_st();
      tc_fence_before();
    }
    if constexpr (PAIR) {          // the leader's MMA warp needs BOTH CTAs' queries in place
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(aready_bar, 0));
    } else {
      asm volatile("bar.sync 1, 192;" ::: "memory");   // epilogue warps (128) + both MMA warps (64): A operand is in TMEM
    }

    const int k = a.k;
    TopList<KL> top;
    top.init(a.lists_in_smem ? reinterpret_cast<uint64_t*>(smem_raw + (bar_base + 512u - raw)) : nullptr, lg * 32 + lane, k);
    float tau = __int_as_float(0x7f800000);          // rejection threshold: min(own k-th best, shared bound); +inf at first
    float tau_g = __int_as_float(0x7f800000);        // last shared bound seen
    uint64_t kth_key = kEmptyKey;
    uint32_t* tau_slot = a.tau_shared + (mt * kM + m);
    float small[8];                                  // own 8 smallest distances, ascending (quantile bound, k > 16 only)
#pragma unroll
    for (int i = 0; i < 8; ++i) small[i] = __int_as_float(0x7f800000);
    float small_q = __int_as_float(0x7f800000);      // small[q - 1]
    const bool use_q = (KL > 16) && a.tau_q != nullptr;
    uint32_t* q_row = use_q ? a.tau_q + static_cast<size_t>(mt * kM + m) * a.cpm : nullptr;
    const float qn = (L2 && q_valid) ? a.q_norm2[b] : 0.0f;
    const float base_n = L2 ? qn + __ldg(a.x_min_norm2) : 0.0f;
    // FILT: rows within `margin` = 2 eps of the running bound are buffered (entry i of thread t at word i * 128 + t)
    float margin = 0.0f;
    uint64_t* fbuf = nullptr;
    int fcnt = 0;
    bool fover = false;
    if constexpr (FILT) {
      fbuf = reinterpret_cast<uint64_t*>(smem_raw + (bar_base + 512u - raw)) + (lg * 32 + lane);
      if (q_valid)
        margin = 2.0f * bf16_contraction_eps(a.q_norm2[b], a.q_lo_norm2[b], __ldg(a.x_max_norm2), __ldg(a.x_lo_max2),
                                             a.guard_rel, L2);
    }

    const bool use_grp = (KL <= 16) && a.tau_grp != nullptr;
    uint32_t* grp_row = use_grp ? a.tau_grp + static_cast<size_t>(mt * kM + m) * a.grp_n : nullptr;
    uint32_t* grp_slot = use_grp ? grp_row + (cj % a.grp_n) : nullptr;
    uint32_t grp_pub = 0xFFFFFFFFu;                  // what this thread last published to its group's slot

    uint32_t it = 0;
    TileWalker walk(a, cj, n_tiles);
    int64_t t;
    uint32_t wt[NW];
    while (walk.next(t, wt)) {                       // tiles with no passing row are skipped by every role
      const int buf = it & (nbuf - 1);
      const uint32_t par = (it >> nbuf_log2) & 1;
      ++it;
      // every 8th tile (and after the 1st, 2nd and 4th, while the bounds still move fast): pick up the bounds the
      // sibling CTAs have published (the loads overlap the wait below)
      const bool refresh = q_valid && ((it & 7u) == 1u || it == 2u || it == 3u || it == 5u);
      uint32_t tg_bits = 0xFFFFFFFFu;
      uint32_t grp_worst = 0xFFFFFFFFu;
      if (refresh) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(tg_bits) : "l"(tau_slot) : "memory");
        if constexpr (KL <= 16) {
          // like the quantile bound of k > 16: every refresh while the bound still moves fast (the first q_early tiles
          // of this CTA), every 128th tile afterwards -- refreshed every 8th tile throughout it cost the headline batch
          // (10M x 768, B = 1024, 8.7 k tiles per CTA) 3-4 %: 13.5-13.9 -> 14.2 ms, same box
          if (use_grp && (it < a.q_early || (it & a.q_late_mask) == 1u)) {
            grp_worst = 0u;
#pragma unroll
            for (int g = 0; g < 16; ++g) {
              if (g < a.grp_n) {
                uint32_t v;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(grp_row + g) : "memory");
                grp_worst = v > grp_worst ? v : grp_worst;
              }
            }
          }
        }
      }
      mbar_wait(accf_bar(buf), par);
      tc_fence_after();
      if (tg_bits != 0xFFFFFFFFu) {
        tau_g = fminf(tau_g, ordered_to_float(tg_bits));
        tau = fminf(tau, tau_g);
      }
      if (grp_worst != 0xFFFFFFFFu) {                // every group has published: k rows lie at or below the largest slot
        tau_g = fminf(tau_g, ordered_to_float(grp_worst));
        tau = fminf(tau, tau_g);
      }
      if constexpr (KL > 16) {
        // the quantile bound costs cpm global loads per THREAD whose latency nothing hides: every 8th tile only while
        // the bound still moves fast (the first 512 tiles of this CTA), every 128th afterwards.  Refreshing it every
        // 8th tile throughout was the single largest cost of k > 16: 25M x 384 l2 k = 100 23.1-24.9 -> 19.2-20.0 ms,
        // 10M x 768 cosine k = 100 16.5-17.3 -> 14.4 ms (the ncu source page had it as a 5.6 % stall on one VIMNMX3,
        // profiles/r02_c5_k100_source_top.txt line 946 -- but every such stall also held a buffer of the pair)
        if (use_q && refresh && (it < a.q_early || (it & a.q_late_mask) == 1u)) {      // T = max over the CTAs' q-th best (all must have one)
          uint32_t worst = 0u;
          for (int c = 0; c < a.cpm; ++c) {
            uint32_t v;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(q_row + c) : "memory");
            worst = v > worst ? v : worst;
          }
          if (worst != 0xFFFFFFFFu) {
            tau_g = fminf(tau_g, ordered_to_float(worst));
            tau = fminf(tau, tau_g);
          }
        }
      }
#pragma unroll
      for (int sub = 0; sub < NSUB; ++sub) {
        // both halves of the 64-column accumulator are requested back to back, then the buffer is
        // handed back to the MMA warp before any score is looked at
        const uint32_t w0 = wt[2 * sub], w1 = wt[2 * sub + 1];
        const bool last_sub = (sub == NSUB - 1);
        uint32_t vv[2][32];
        const uint32_t acc_addr = tmem_acc + (static_cast<uint32_t>(lg * 32) << 16) + static_cast<uint32_t>(buf * kAccCols + sub * 64);
        // tcgen05.ld / tcgen05.wait are .sync.aligned: the lanes, which leave the candidate loops of the previous
        // tile one by one, must be back together here
        __syncwarp();
        if ((w0 | w1) != 0u) {                           // a half with no passing row is not even read
          tmem_ld_x32(acc_addr, vv[0]);
          tmem_ld_x32(acc_addr + 32, vv[1]);
          tmem_wait_ld();
        }
        if (last_sub) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(mapa_shared(acce_bar(buf), 0));   // the leader issues for both
            else mbar_arrive(acce_bar(buf));
          }
        }
        if ((w0 | w1) == 0u) continue;
        if (a.stats && lane == 0) atomicAdd(&g_tensor_stats[0], 1ull);
        // Fast reject: once the lists have warmed up almost no tile holds a candidate for any of the
        // warp's 32 queries.  The largest of the 64 dot products (a tree of 32 three-input max
        // instructions) is tested against a bound that no candidate can fall below; only if some lane
        // passes does the warp go on, and then each lane DESCENDS the same tree (7 subtree maxima -> 22
        // group maxima -> scores) to the few scores that pass instead of testing all 64.
        float m32[11], n32[11];                        // maxima of groups of 3 scores (columns 0-31 / 32-63)
  #pragma unroll
        for (int j = 0; j < 10; ++j)
          m32[j] = fmax3(__uint_as_float(vv[0][3 * j]), __uint_as_float(vv[0][3 * j + 1]), __uint_as_float(vv[0][3 * j + 2]));
        m32[10] = fmaxf(__uint_as_float(vv[0][30]), __uint_as_float(vv[0][31]));
  #pragma unroll
        for (int j = 0; j < 10; ++j)
          n32[j] = fmax3(__uint_as_float(vv[1][3 * j]), __uint_as_float(vv[1][3 * j + 1]), __uint_as_float(vv[1][3 * j + 2]));
        n32[10] = fmaxf(__uint_as_float(vv[1][30]), __uint_as_float(vv[1][31]));
        const float t0 = fmax3(m32[0], m32[1], m32[2]), t1 = fmax3(m32[3], m32[4], m32[5]);
        const float t2 = fmax3(m32[6], m32[7], m32[8]), t3 = fmax3(m32[9], m32[10], n32[0]);
        const float t4 = fmax3(n32[1], n32[2], n32[3]), t5 = fmax3(n32[4], n32[5], n32[6]);
        const float t6 = fmax3(n32[7], n32[8], n32[9]);
        const float best = fmax3(fmax3(t0, t1, t2), fmax3(t3, t4, t5), fmaxf(t6, n32[10]));
        const float tau_r = FILT ? tau + margin : tau; // FILT: everything within `margin` of the bound is wanted
        float lb;                                      // lower bound on the dot product of any candidate of this lane
        if constexpr (L2) {
          // d = |q|^2 + |x|^2 - 2 q.x <= tau  =>  q.x >= (|q|^2 + min|x|^2 - tau) / 2 (minus rounding slack),
          // with min|x|^2 over everything the store ever held (tracked by the upsert kernel): no per-tile loads
          lb = 0.5f * (base_n - tau_r) - 4e-7f * (fabsf(base_n) + fabsf(tau_r));
        } else {
          // cosine / ip: fl(1 - dot) <= tau implies dot >= 1 - tau - 2^-24
          lb = (1.0f - tau_r) - 1.2e-7f;
        }
        if (!__any_sync(0xffffffffu, q_valid && best >= lb)) continue;
        if (a.stats && lane == 0) atomicAdd(&g_tensor_stats[1], 1ull);
        if constexpr (L2) {
          // second-level reject with THIS tile's smallest |x|^2 (the store-wide minimum above is loose when the
          // rows' norms vary; unit-norm embeddings never get further with it)
          const int64_t r0 = t * NB + sub * 64 + lane, r1 = r0 + 32;
          const float xn0 = (r0 < a.n_rows) ? __ldg(a.x_norm2 + r0) : __int_as_float(0x7f800000);
          const float xn1 = (r1 < a.n_rows) ? __ldg(a.x_norm2 + r1) : __int_as_float(0x7f800000);
          float xmin = fminf(xn0, xn1);
  #pragma unroll
          for (int off = 16; off > 0; off >>= 1) xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, off));
          const float bn = qn + xmin;
          lb = 0.5f * (bn - tau_r) - 4e-7f * (fabsf(bn) + fabsf(tau_r));
          if (!__any_sync(0xffffffffu, q_valid && best >= lb)) continue;
          if (a.stats && lane == 0) atomicAdd(&g_tensor_stats[2], 1ull);
        }
        // ---- descent: candidate bits of this lane (a superset; the key comparison below is the exact test) ----
        uint32_t c0 = 0u, c1 = 0u;
        if (q_valid && best >= lb) {
  #define RAG_GROUP(GMAX, CBITS, H, G)                                                              \
          if ((GMAX) >= lb) {                                                                       \
            if (__uint_as_float(vv[H][3 * (G)]) >= lb) CBITS |= 1u << (3 * (G));                     \
            if (__uint_as_float(vv[H][3 * (G) + 1]) >= lb) CBITS |= 1u << (3 * (G) + 1);             \
            if (3 * (G) + 2 < 32 && __uint_as_float(vv[H][(3 * (G) + 2) & 31]) >= lb) CBITS |= 1u << ((3 * (G) + 2) & 31); \
          }
          if (t0 >= lb) { RAG_GROUP(m32[0], c0, 0, 0) RAG_GROUP(m32[1], c0, 0, 1) RAG_GROUP(m32[2], c0, 0, 2) }
          if (t1 >= lb) { RAG_GROUP(m32[3], c0, 0, 3) RAG_GROUP(m32[4], c0, 0, 4) RAG_GROUP(m32[5], c0, 0, 5) }
          if (t2 >= lb) { RAG_GROUP(m32[6], c0, 0, 6) RAG_GROUP(m32[7], c0, 0, 7) RAG_GROUP(m32[8], c0, 0, 8) }
          if (t3 >= lb) { RAG_GROUP(m32[9], c0, 0, 9) RAG_GROUP(m32[10], c0, 0, 10) RAG_GROUP(n32[0], c1, 1, 0) }
          if (t4 >= lb) { RAG_GROUP(n32[1], c1, 1, 1) RAG_GROUP(n32[2], c1, 1, 2) RAG_GROUP(n32[3], c1, 1, 3) }
          if (t5 >= lb) { RAG_GROUP(n32[4], c1, 1, 4) RAG_GROUP(n32[5], c1, 1, 5) RAG_GROUP(n32[6], c1, 1, 6) }
          if (t6 >= lb) { RAG_GROUP(n32[7], c1, 1, 7) RAG_GROUP(n32[8], c1, 1, 8) RAG_GROUP(n32[9], c1, 1, 9) }
          RAG_GROUP(n32[10], c1, 1, 10)
  #undef RAG_GROUP
          c0 &= w0;                                    // live & filter bits of the tile's rows
          c1 &= w1;
          if (a.stats && (c0 | c1)) atomicAdd(&g_tensor_stats[3], static_cast<unsigned long long>(__popc(c0) + __popc(c1)));
        }
  #pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t* v = vv[half];
          uint32_t cand = half ? c1 : c0;
          while (cand) {
            const int j = __ffs(cand) - 1;
            cand &= cand - 1;
            float sc = 0.0f;
  #pragma unroll
            for (int jj = 0; jj < 32; ++jj) sc = (jj == j) ? __uint_as_float(v[jj]) : sc;
            const uint32_t row = static_cast<uint32_t>(t * NB + sub * 64 + half * 32 + j);
            float dj;
            if constexpr (L2) dj = fmaxf(fmaf(-2.0f, sc, qn + __ldg(a.x_norm2 + row)), 0.0f);
            else dj = 1.0f - sc;
            const uint64_t key = make_key(dj, row);
            if constexpr (KL > 16) {
              if (use_q && dj < small_q) {             // keep the own q smallest distances and publish the q-th
                if (a.stats) atomicAdd(&g_tensor_stats[5], 1ull);
                float x = dj;
  #pragma unroll
                for (int i = 0; i < 8; ++i) { const float c = small[i]; const bool lt = x < c; small[i] = lt ? x : c; x = lt ? c : x; }
                small_q = small[0];
  #pragma unroll
                for (int i = 1; i < 8; ++i) small_q = (i == a.tau_q_rank - 1) ? small[i] : small_q;
                if (small_q < __int_as_float(0x7f800000))
                  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(q_row + cj), "r"(float_to_ordered(small_q)) : "memory");
              }
            }
            if constexpr (FILT) {
              if (dj <= tau + margin) {                // (tau = +inf until the list is full: everything is kept)
                if (fcnt == kFiltCap) {                // full: drop what the bound has moved past since it was buffered
                  const float thr = tau + margin;
                  int w = 0;
                  for (int i = 0; i < kFiltCap; ++i) {
                    const uint64_t e = fbuf[i * kEpiThreads];
                    if (key_dist(e) <= thr) { fbuf[w * kEpiThreads] = e; ++w; }
                  }
                  fcnt = w;
                }
                if (fcnt < kFiltCap) { fbuf[fcnt * kEpiThreads] = key; ++fcnt; }
                else fover = true;                     // more than kFiltCap rows within the margin: exact re-run
              }
            }
            if (key < kth_key && dj <= tau_g) {
              if (a.stats) atomicAdd(&g_tensor_stats[4], 1ull);
              top.insert(key, k);
              if constexpr (KL <= 16) {
                if (use_grp) {                         // own r-th best -> the group's slot, when it has improved
                  const uint64_t rk = top.kth(a.grp_rank);
                  const uint32_t rv = static_cast<uint32_t>(rk >> 32);
                  if (rk != kEmptyKey && rv < grp_pub) { atomicMin(grp_slot, rv); grp_pub = rv; }
                }
              }
              kth_key = top.kth(k);
              if (kth_key != kEmptyKey) {              // list is full: its k-th best bounds the global k-th best
                const float own = key_dist(kth_key);
                atomicMin(tau_slot, float_to_ordered(own));
                tau = fminf(own, tau_g);
              }
            }
          }
        }
      }
    }
    if constexpr (FILT) {
      if (q_valid) {           // the buffered rows still within the margin of the final bound
        const size_t slot = static_cast<size_t>(cj) * a.B + b;
        const float thr = tau + margin;
        uint64_t* out = a.extra + slot * kFiltCap;
        int w = 0;
        for (int i = 0; i < fcnt; ++i) {
          const uint64_t e = fbuf[i * kEpiThreads];
          if (key_dist(e) <= thr) out[w++] = e;
        }
        a.extra_cnt[slot] = fover ? -1 : w;
      }
    }
    // ---- emit this thread's list: partial[cj][b][0..k) ----
    if (q_valid) {
      top.finalize(k);
      uint64_t* out = a.partial + (static_cast<size_t>(cj) * a.B + b) * k;
      if constexpr (KL <= 16) {
#pragma unroll
        for (int i = 0; i < KL; ++i) if (i < k) out[i] = top.e[i];
      } else {
        for (int i = 0; i < k; ++i) out[i] = top.at(i);
      }
    }
    tc_fence_before();
  } else if (warp == 0) {
    // ================= TMA producer (whole warp walks the loop, one elected lane issues) =========
    uint32_t s = 0, ph = 0;
    TileWalker walk(a, cj, n_tiles, false);
    int64_t t;
    uint32_t wt[NW];
    const int n_sib = (a.progress != nullptr) ? static_cast<int>(gridDim.y) : 1;   // clusters sharing this cj's tiles
    uint32_t* prog = a.progress + static_cast<size_t>(cj) * n_sib;
    uint32_t ti = 0;                                           // tiles this CTA has issued
    while (walk.next(t, wt)) {
      if (n_sib > 1 && (ti & 3u) == 0u) {
        if (crank == 0 && lane == 0)
          asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(prog + blockIdx.y), "r"(0xFFFFFFFFu - ti) : "memory");
        if (ti >= a.pace_window) {
          for (int spins = 0; spins < 100000; ++spins) {       // bounded: pacing is an optimisation, never a dependency
            uint32_t v = 0u;
            if (lane < n_sib) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(prog + lane) : "memory");
            const uint32_t slowest = __reduce_min_sync(0xffffffffu, 0xFFFFFFFFu - v);
            if (slowest + a.pace_window >= ti) break;
            __nanosleep(200);
          }
        }
      }
      ++ti;
      const int row0 = static_cast<int>(t * NB);
      // optionally pull this CTA's slice of a tile a.prefetch steps ahead into L2 (hides DRAM latency
      // when the CTAs sharing a corpus tile have drifted apart and it is no longer L2-resident)
      {
        const int64_t tp = t + static_cast<int64_t>(a.prefetch) * a.cpm;
        if (a.prefetch > 0 && tp < n_tiles && elect_one()) {
          const int rp = static_cast<int>(tp * NB) + static_cast<int>(crank) * (NB / CL);
          for (int kc = 0; kc < a.row_elems; kc += kAtomK) tma_prefetch_l2_2d(&tmap, kc, rp);
        }
        __syncwarp();
      }
      for (int ks = 0; ks < n_stages_per_tile; ++ks) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const int k0 = ks * kStageK;
        int atoms = (a.row_elems - k0 + kAtomK - 1) / kAtomK;       // atoms that hold real columns
        atoms = atoms > kAtomsPerStage ? kAtomsPerStage : atoms;
        if (elect_one()) {
          const int r0 = row0 + static_cast<int>(crank) * (NB / CL);
          if constexpr (PAIR) {
            // the leader's barrier collects the bytes of both halves; each CTA fills its own ring
            const uint32_t lbar = mapa_shared(full_bar(s), 0);
            if (crank == 0) mbar_arrive_expect_tx(full_bar(s), static_cast<uint32_t>(atoms) * kAtomBytes * 2);
            const uint32_t dst = base + s * kStageBytes;
#pragma unroll
            for (int at = 0; at < kAtomsPerStage; ++at)
              if (at < atoms) tma_load_2d_pair(dst + at * kAtomBytes, &tmap, lbar, k0 + at * kAtomK, r0);
          } else {
            mbar_arrive_expect_tx(full_bar(s), static_cast<uint32_t>(atoms) * kAtomBytes);   // bytes from all CL loaders
            const uint32_t dst = base + s * kStageBytes + crank * (kAtomBytes / CL);
#pragma unroll
            for (int at = 0; at < kAtomsPerStage; ++at) {
              if (at < atoms) {
                if constexpr (CL == 1) tma_load_2d(dst + at * kAtomBytes, &tmap, full_bar(s), k0 + at * kAtomK, r0);
                else tma_load_2d_mc(dst + at * kAtomBytes, &tmap, full_bar(s), k0 + at * kAtomK, r0, kMcMask);
              }
            }
          }
        }
        __syncwarp();
        if (++s == n_ring) { s = 0; ph ^= 1u; }
      }
    }
    if (n_sib > 1 && crank == 0 && lane == 0)                 // finished: nobody should ever wait for this cluster
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(prog + blockIdx.y), "r"(0u) : "memory");
  } else {
    // ================= MMA issuers (warps 1 and 6; in PAIR mode only the leader CTA's) =================
    // TWO issuing warps take alternate tiles.  A tile costs the issuing thread ~310 (D = 384) to ~450
    // (D = 768) mostly dependent scalar/uniform instructions (barrier polls, descriptor and column
    // arithmetic, 3 per MMA) at ~4.7 cycles each -- 1.5-2.1 k cycles, MORE than the tile's 768-1536
    // cycles on the tensor pipe (ncu: the single issuer was never back-pressured by UTCHMMA, it was simply
    // busy).  The MMAs of consecutive tiles touch different accumulators and ring stages, so two threads
    // can issue them independently; each commits what it issued.
    if constexpr (PAIR) {
      if (crank == 0) mbar_wait(aready_bar, 0);      // 8 epilogue warps of the pair have stored their queries
    } else {
      asm volatile("bar.sync 1, 192;" ::: "memory");   // wait for the A operand
    }
    tc_fence_after();
    uint32_t s = 0, ph = 0, it = 0;
    const uint64_t desc0 = make_b_desc(base);        // descriptor of stage 0, atom 0, k = 0
    constexpr uint32_t idesc = make_idesc(NB, PAIR ? 2 * kM : kM);
    auto mma = [](uint32_t d, uint32_t at, uint64_t bd, uint32_t acc) {
      if constexpr (PAIR) umma_ts_pair(d, at, bd, idesc, acc);
      else umma_ts(d, at, bd, idesc, acc);
    };
    TileWalker walk(a, cj, (PAIR && crank != 0) ? 0 : n_tiles, false);     // the peer's MMA warps issue nothing
    const uint32_t issuer = (warp == kMmaWarpB) ? 1u : 0u;
    int64_t t;
    uint32_t wt[NW];
    while (walk.next(t, wt)) {
      const int buf = it & (nbuf - 1);
      const uint32_t par = (it >> nbuf_log2) & 1;
      // one issuer where the ring geometry does not let two issuers run without watching each other's stages
      // (launch_one decides, with the reasons)
      const bool mine = a.one_issuer ? (issuer == 0u) : ((it & 1u) == issuer);
      ++it;
      if (!mine) {                                  // the other issuer's tile: step over its ring stages
        s += static_cast<uint32_t>(n_stages_per_tile);
        if (s >= n_ring) { s -= n_ring; ph ^= 1u; }
        continue;
      }
      mbar_wait(acce_bar(buf), par ^ 1u);           // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d_tmem = tmem_acc + static_cast<uint32_t>(buf * kAccCols);
      for (int ks = 0; ks < n_stages_per_tile; ++ks) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          // descriptors advance by plain adds: stage (32 KB), atom (8 KB), k-step inside an atom (32 B), in 16-byte units
          const uint64_t dstage = desc0 + static_cast<uint64_t>((s * kStageBytes) >> 4);
          const uint32_t a0 = tmem_base + static_cast<uint32_t>(ks * kMmasPerStage * 8);
          const int left = k_steps - ks * kMmasPerStage;
          if (a.split_steps > 0) {
            // fp32 store contracted as bf16 pairs: q.x ~ qh.xh + ql.xh + qh.xl  (ql.xl < 2^-18 |q||x| dropped).
            // B k-step g of the [hi | lo] shadow row pairs with A columns of q_hi (and q_lo while g is a hi step).
            // fully unrolled so that the descriptor offsets are immediates and everything stays in uniform
            // registers (a rolled loop issued one MMA per ~100 cycles: 3.0 ms instead of ~1.5 for 1M x 384)
            const int g0 = ks * kMmasPerStage;
            const int sp = a.split_steps;
            const uint32_t a_hi = tmem_base + static_cast<uint32_t>(g0 * 8);          // q_hi columns of this stage's first k-step
            const uint32_t a_lo = a_hi + static_cast<uint32_t>(sp * 8);              // matching q_lo columns
            const uint32_t a_hx = a_hi - static_cast<uint32_t>(sp * 8);              // q_hi columns for an x_lo k-step
#pragma unroll
            for (int j = 0; j < kMmasPerStage; ++j) {
              if (j < left) {
                const uint64_t bdesc = dstage + (((j >> 2) * kAtomBytes + (j & 3) * 32) >> 4);
                if (g0 + j < sp) {
                  mma(d_tmem, a_hi + j * 8, bdesc, (g0 + j > 0) ? 1u : 0u);
                  mma(d_tmem, a_lo + j * 8, bdesc, 1u);
                } else {
                  mma(d_tmem, a_hx + j * 8, bdesc, 1u);
                }
              }
            }
          } else if (left >= kMmasPerStage) {
#pragma unroll
            for (int j = 0; j < kMmasPerStage; ++j)
              mma(d_tmem, a0 + j * 8, dstage + (((j >> 2) * kAtomBytes + (j & 3) * 32) >> 4),
                  (j > 0) ? 1u : (ks > 0 ? 1u : 0u));
          } else {
#pragma unroll
            for (int j = 0; j < kMmasPerStage; ++j)
              if (j < left)
                mma(d_tmem, a0 + j * 8, dstage + (((j >> 2) * kAtomBytes + (j & 3) * 32) >> 4),
                    (j > 0) ? 1u : (ks > 0 ? 1u : 0u));
          }
          if constexpr (PAIR) {
            umma_commit_pair(empty_bar(s), kMcMask);           // both CTAs' producers may refill their half
            if (ks == n_stages_per_tile - 1) umma_commit_pair(accf_bar(buf), kMcMask);   // both epilogues may drain
          } else {
            if constexpr (CL == 1) umma_commit(empty_bar(s));   // stage reusable once these MMAs retire
            else umma_commit_mc(empty_bar(s), kMcMask);          // ... in every CTA that loads into it
            if (ks == n_stages_per_tile - 1) umma_commit(accf_bar(buf));
          }
        }
        __syncwarp();
        if (++s == n_ring) { s = 0; ph ^= 1u; }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // no peer may still multicast into / arrive on this CTA
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host side --------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct Layout {
  int n_mtiles, cpm, cl, mode, nb;
  size_t off_qbf16, off_qnorm, off_qlo, off_qf32, off_qexact, off_partial, off_tau, off_prog, off_tauq, off_merged, off_extra, off_extracnt, total;
};

// row_elems: width of the bf16 rows the kernel streams (2 x the store's for split precision); k: list length kept
Layout make_layout(int row_elems, int B, int k, int sm_count) {
  Layout L{};
  L.n_mtiles = (B + kM - 1) / kM;
  L.cl = (L.n_mtiles >= 2) ? 2 : 1;
  if (const char* e = getenv("RAG_B200_TENSOR_CL")) { if (atoi(e) == 1) L.cl = 1; }
  // two query tiles or more: CTA pairs (cta_group::2); RAG_B200_TENSOR_MODE=1 selects the older multicast pair
  L.mode = (L.cl == 1) ? 0 : 2;
  if (const char* e = getenv("RAG_B200_TENSOR_MODE")) { if (L.cl == 2 && atoi(e) == 1) L.mode = 1; }
  // wide (128-row) corpus tiles, possible where tensor memory has room for two 128-column accumulators behind the
  // A operand (D <= 512) and the CTAs run as pairs.  Measured: 3 % faster at k = 10 on 25M x 384,
  // 4 % SLOWER at k = 100 and on 10M x 384 -- the per-instruction issue cost it halves is not the bound -- so it is
  // an experiment switch (RAG_B200_TENSOR_NB=128), off by default.
  L.nb = kNB;
  if (L.mode == 2 && ((row_elems + 15) / 16) * 8 + 2 * kNBWide <= kTmemCols &&
      getenv("RAG_B200_TENSOR_NB") && atoi(getenv("RAG_B200_TENSOR_NB")) == 128)
    L.nb = kNBWide;
  L.n_mtiles = (L.n_mtiles + L.cl - 1) / L.cl * L.cl;   // pad with idle query tiles to whole clusters
  L.cpm = sm_count / L.n_mtiles;
  if (L.cpm < 1) L.cpm = 1;
  size_t off = 0;
  L.off_qbf16 = off; off += align256(static_cast<size_t>(L.n_mtiles) * kM * row_elems * 2);
  L.off_qnorm = off; off += align256(static_cast<size_t>(L.n_mtiles) * kM * 4);
  L.off_qlo = off; off += align256(static_cast<size_t>(L.n_mtiles) * kM * 4);
  L.off_qf32 = off;  off += align256(static_cast<size_t>(B) * row_elems * 4);
  L.off_qexact = off; off += align256(static_cast<size_t>(B) * row_elems * 4);   // un-rounded queries (exact re-ranking)
  L.off_partial = off; off += align256(static_cast<size_t>(L.cpm) * B * k * 8);
  L.off_tau = off; off += align256(static_cast<size_t>(L.n_mtiles) * kM * 4);      // directly behind `partial`: one memset
  L.off_prog = off; off += align256(static_cast<size_t>(L.cpm) * L.n_mtiles * 4);  // ... which also covers the pacing counters
  L.off_tauq = off; off += align256(static_cast<size_t>(L.n_mtiles) * kM * L.cpm * 4);   // ... and the quantile bounds
  L.off_merged = off; off += align256(static_cast<size_t>(B) * k * 8);
  L.off_extra = off; off += align256(static_cast<size_t>(L.cpm) * B * kFiltCap * 8);      // FILT buffers (always sized: cheap)
  L.off_extracnt = off; off += align256(static_cast<size_t>(L.cpm) * B * 4);
  L.total = off;
  return L;
}

template <int KL, bool L2, int MODE, int NB, bool FILT = false>
cudaError_t launch_one(const CUtensorMap& tmap, Args a, dim3 grid, cudaStream_t st) {
  constexpr int CL = (MODE == 0) ? 1 : 2;
  using R = Ring<MODE, NB>;
  auto kern = gemm_topk_kernel<KL, L2, MODE, NB, FILT>;
  // k <= 16: lists in registers, the whole 224 KB ring.  16 < k <= 128: the heaps take k x 128 x 8 bytes of shared
  // memory behind the barriers, the ring gets what is left (>= 96 KB; its depth does not limit the kernel: 4 stages
  // measured as fast as 14).  Larger k: heaps in local memory -> a short ring and the rest of the SM left to L1.
  const int per_tile = (a.row_elems + kStageK - 1) / kStageK;
  int list_bytes = TopList<KL>::kShared ? a.k * kEpiThreads * 8 : (FILT ? kFiltCap * kEpiThreads * 8 : 0);
  if (!FILT && list_bytes > 0 && (kRingBytes - list_bytes) / R::kStageBytes < per_tile + 1) list_bytes = 0;   // no room: heaps in local memory
  a.lists_in_smem = (!FILT && list_bytes > 0) ? 1 : 0;
  a.stages_used = R::kStages;
  if (list_bytes > 0) a.stages_used = std::min(R::kStages, (kRingBytes - list_bytes) / R::kStageBytes);
  else if (KL > 16) a.stages_used = std::min(4, (96 * 1024) / R::kStageBytes);
  if (const char* ev = getenv("RAG_B200_TENSOR_STAGES")) {
    const int v = atoi(ev);
    if (v >= 2 && v <= a.stages_used) a.stages_used = v;
  }
  // the ring must hold at least one whole tile plus a stage, or producer and issuer wait for each other
  if (a.stages_used < per_tile + 1) a.stages_used = std::min(per_tile + 1, R::kStages);
  plan_ring(a.stages_used, per_tile, a.nbuf, &a.stages_used, &a.one_issuer);
  if (const char* ev = getenv("RAG_B200_TENSOR_ONE_ISSUER")) { if (atoi(ev) == 1) a.one_issuer = 1; }
  const int smem_bytes = a.stages_used * R::kStageBytes + 1024 /*align*/ + 512 /*barriers*/ + list_bytes;
  if (smem_bytes > kSmemBytes) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                           (KL > 16 && list_bytes == 0) ? std::min(100, (smem_bytes + 8 * 1024) * 100 / (228 * 1024) + 1) : 100);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tmap, a);
}

// the hi-only filter of fp32 stores (k <= 16): single CTAs or cta_group::2 pairs, 64-row tiles
cudaError_t launch_filter(const CUtensorMap& tmap, const Args& a, bool l2, int mode, dim3 grid, cudaStream_t st) {
  if (mode == 2) return l2 ? launch_one<16, true, 2, kNB, true>(tmap, a, grid, st) : launch_one<16, false, 2, kNB, true>(tmap, a, grid, st);
  if (mode == 0) return l2 ? launch_one<16, true, 0, kNB, true>(tmap, a, grid, st) : launch_one<16, false, 0, kNB, true>(tmap, a, grid, st);
  return cudaErrorNotSupported;
}

template <int KL>
cudaError_t launch_kl(const CUtensorMap& tmap, const Args& a, bool l2, int mode, int nb, dim3 grid, cudaStream_t st) {
  if (mode == 2 && nb == kNBWide)
    return l2 ? launch_one<KL, true, 2, kNBWide>(tmap, a, grid, st) : launch_one<KL, false, 2, kNBWide>(tmap, a, grid, st);
  if (mode == 2) return l2 ? launch_one<KL, true, 2, kNB>(tmap, a, grid, st) : launch_one<KL, false, 2, kNB>(tmap, a, grid, st);
  if (mode == 1) return l2 ? launch_one<KL, true, 1, kNB>(tmap, a, grid, st) : launch_one<KL, false, 1, kNB>(tmap, a, grid, st);
  return l2 ? launch_one<KL, true, 0, kNB>(tmap, a, grid, st) : launch_one<KL, false, 0, kNB>(tmap, a, grid, st);
}

}  // namespace

// Ring stages the kernel may use with `avail` stages of shared memory, p = per_tile stages per corpus tile and nbuf
// accumulators, and whether ONE warp has to issue every tile.
// Two MMA issuers take alternate tiles and never look at the full barriers of each other's stages.  An issuer about
// to poll a stage must therefore know, by other means, that the stage's PREVIOUS use has landed (TMA loads complete
// out of order) -- else its parity test passes on the phase before and everything derails (launch failures with
// 4-5 stages at D = 768; tools/ring_protocol_model.py fails the same way; tests/test_ring_protocol.py holds this
// function to that model).  What an issuer of tile t does know: its own tile t-2 (it polled those stages) and
// every tile the epilogue has drained, i.e. up to t - nbuf (it waits for that accumulator).  The previous use of a
// stage of tile t lies n stages back, so:
//   n == 2p            every stage always belongs to the same issuer                      -> safe
//   n >= 2p, nbuf == 2 the previous use lies in tile t-2 or before                        -> safe
//   n >= 4p            ... in tile t-4 or before                                          -> safe
//   2p < n < 4p with 4 accumulators: it can lie in the OTHER issuer's tile t-3            -> clamp to 2p
//   n < 2p             ... in the other issuer's tile t-1                                 -> one issuer
void plan_ring(int avail, int per_tile, int nbuf, int* stages, int* one_issuer) {
  const int p2 = 2 * per_tile;
  const bool safe_two = (avail == p2) || (avail >= p2 && nbuf == 2) || (avail >= 2 * p2);
  *stages = avail;
  *one_issuer = 0;
  if (!safe_two) {
    if (avail > p2) *stages = p2;
    else *one_issuer = 1;
  }
}

static bool split_enabled() {
  static const bool on = !(getenv("RAG_B200_F32_TENSOR") && atoi(getenv("RAG_B200_F32_TENSOR")) == 0);
  return on;
}

int default_shadow_kind(int row_elems) {
  const bool hi_ok = row_elems >= 8 && row_elems % 8 == 0 && ((row_elems + 15) / 16) * 8 <= kMaxKCols;
  const bool hilo_ok = row_elems >= 16 && row_elems % 16 == 0 && row_elems <= kMaxKCols;
  if (const char* e = getenv("RAG_B200_F32_SHADOW")) {
    if (!strcmp(e, "hilo") && hilo_ok) return kShadowHiLo;
    if (!strcmp(e, "hi") && hi_ok) return kShadowHi;
  }
  // measured (1M x 384 fp32, B = 32 / 1024, top-10): hi/lo split 0.36 / 1.87 ms, hi-only filter 0.28 / 1.15 ms.
  // (k > 16 on a hi-only store keeps 64-128 candidates in per-thread heaps instead of the filter buffers, which is
  // slower than the split; stores queried that way can be pinned with rag_store_set_f32_shadow.)
  return hi_ok ? kShadowHi : (hilo_ok ? kShadowHiLo : kShadowNone);
}

static bool filter_enabled() {
  static const bool on = !(getenv("RAG_B200_F32_FILTER") && atoi(getenv("RAG_B200_F32_FILTER")) == 0);
  return on;
}

int candidates_kept(int dtype, int k, int rerank, int shadow_kind) {
  if (dtype == 1 && !rerank) return k;
  // hi-only filter, k <= 16: the top-k list itself stays in registers; the rows within 2 eps of it go to the
  // per-thread buffers (Args::extra), not into a longer list
  if (dtype != 1 && shadow_kind == kShadowHi && k <= 16 && filter_enabled()) return 16;
  // hi-only filter of an fp32 store: bf16 rounding moves a unit-norm dot product by ~1e-3 worst case, so the list
  // must reach that far beyond the k-th neighbour for the guard to certify it (1M x 384 unit-norm: ~rank 20-40)
  if (dtype != 1 && shadow_kind == kShadowHi) return k <= 32 ? 64 : (k <= 100 ? 128 : k + 64);
  return k <= 10 ? 16 : k + 16;      // slack for the exact re-ranking of approximately ranked rows
}

bool supported(int dtype, int row_elems, int k, int space, int rerank, int shadow_kind) {
  (void)space;
  if (get_encode() == nullptr || k < 1) return false;
  if (dtype == 1) return row_elems >= 8 && ((row_elems + 15) / 16) * 8 <= kMaxKCols && candidates_kept(dtype, k, rerank, 0) <= 1024;
  if (!split_enabled() || candidates_kept(dtype, k, 0, shadow_kind) > 1024) return false;
  // fp32 rows as bf16(x): an ordinary bf16 contraction over the shadow (TMA needs 16-byte row pitches)
  if (shadow_kind == kShadowHi) return row_elems >= 8 && row_elems % 8 == 0 && ((row_elems + 15) / 16) * 8 <= kMaxKCols;
  // fp32 rows as [hi | lo] bf16: the A operand takes row_elems TMEM columns; k-steps must not straddle hi/lo
  if (shadow_kind == kShadowHiLo) return row_elems >= 16 && row_elems % 16 == 0 && row_elems <= kMaxKCols;
  return false;
}

size_t scratch_bytes(int dtype, int row_elems, int B, int k, int sm_count, int rerank, int shadow_kind) {
  if (!supported(dtype, row_elems, k, 0, rerank, shadow_kind)) return 0;
  const int width = (dtype != 1 && shadow_kind == kShadowHiLo) ? 2 * row_elems : row_elems;
  return make_layout(width, B, candidates_kept(dtype, k, rerank, shadow_kind), sm_count).total;
}

cudaError_t launch(const Problem& p, cudaStream_t st, Result* out, int* launches) {
  if (!supported(p.dtype, p.row_elems, p.k, p.space, p.rerank, p.shadow_kind)) return cudaErrorNotSupported;
  const bool f32 = (p.dtype != 1);
  const bool split = f32 && p.shadow_kind == kShadowHiLo;        // [hi | lo] rows, three MMAs per k-step
  const bool hi_only = f32 && p.shadow_kind == kShadowHi;        // bf16(x) rows: an ordinary bf16 contraction
  const int width = split ? 2 * p.row_elems : p.row_elems;       // bf16 elements per streamed row
  const int kk = candidates_kept(p.dtype, p.k, p.rerank, p.shadow_kind);
  if (f32 && p.shadow == nullptr) return cudaErrorInvalidValue;
  const Layout L = make_layout(width, p.B, kk, p.sm_count);
  __nv_bfloat16* q_bf16 = reinterpret_cast<__nv_bfloat16*>(p.scratch + L.off_qbf16);
  float* q_norm = reinterpret_cast<float*>(p.scratch + L.off_qnorm);
  float* q_f32 = reinterpret_cast<float*>(p.scratch + L.off_qf32);
  uint64_t* part = reinterpret_cast<uint64_t*>(p.scratch + L.off_partial);

  // queries: zero the padded tile rows, then normalise / round / convert
  cudaError_t e = cudaMemsetAsync(q_bf16, 0, static_cast<size_t>(L.n_mtiles) * kM * width * 2, st);
  if (e != cudaSuccess) return e;
  PrepArgs pa{};
  pa.src = p.queries_raw; pa.B = p.B; pa.dim = p.dim; pa.row_elems = p.row_elems;
  // fp32 stores: q_f32 and |q|^2 stay un-rounded (the refinement and the norms are exact; only the A operand is bf16)
  pa.normalise = (p.space == 1); pa.round_bf16 = f32 ? 0 : 1; pa.split = split ? 1 : 0;
  pa.q_f32 = q_f32; pa.q_bf16 = q_bf16; pa.q_norm2 = q_norm;
  float* q_lo = reinterpret_cast<float*>(p.scratch + L.off_qlo);
  pa.q_lo_norm2 = hi_only ? q_lo : nullptr;
  float* q_exact = reinterpret_cast<float*>(p.scratch + L.off_qexact);
  pa.q_exact = p.rerank ? q_exact : nullptr; pa.exact_elems = p.exact_elems;
  e = launch_prep_queries(pa, st);
  if (e != cudaSuccess) return e;

  // TMA descriptor over the live part of the corpus: [n_rows][row_elems] bf16, box 64 x 64, 128B swizzle
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(p.n_rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(width) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kAtomK), static_cast<cuuint32_t>(L.nb / L.cl)};
  const cuuint32_t estride[2] = {1, 1};
  // 128-byte promotion = exactly the box row (64 bf16); 256 B fetched the neighbouring k-slice early and cost the
  // HBM-bound small batches 4 % (10M x 768, B = 32: 2.565 -> 2.469 ms); no effect at B = 1024
  CUtensorMapL2promotion l2promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (const char* ev = getenv("RAG_B200_TENSOR_L2PROMO")) {
    const int v = atoi(ev);
    l2promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
              : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  }
  CUresult r = get_encode()(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                            const_cast<void*>(f32 ? static_cast<const void*>(p.shadow) : p.vectors), gdim, gstride, box,
                            estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;

  Args a{};
  a.q_bf16 = q_bf16; a.q_norm2 = q_norm; a.x_norm2 = p.norms2; a.x_min_norm2 = p.min_norm2;
  a.live = p.live; a.filter = p.filter; a.filter_words = p.filter_words;
  a.n_rows = p.n_rows; a.row_elems = width; a.B = p.B; a.k = kk; a.cpm = L.cpm; a.partial = part;
  a.split_steps = split ? p.row_elems / 16 : 0;
  a.nbuf = (L.nb == kNB && ((width + 15) / 16) * 8 + 4 * kNB <= kTmemCols) ? 4 : 2;
  if (const char* e = getenv("RAG_B200_TENSOR_NBUF")) { if (atoi(e) == 2) a.nbuf = 2; }
  a.dense = p.dense;
  a.q_early = 512u;
  a.q_late_mask = 127u;
  if (const char* ev = getenv("RAG_B200_TENSOR_QEARLY")) a.q_early = static_cast<uint32_t>(atoi(ev));
  if (const char* ev = getenv("RAG_B200_TENSOR_QLATE")) a.q_late_mask = (1u << atoi(ev)) - 1u;      // log2 of the late interval (>= 3)
  static const int stats_on = (getenv("RAG_B200_TENSOR_STATS") && atoi(getenv("RAG_B200_TENSOR_STATS")) == 1) ? 1 : 0;
  a.stats = stats_on;
  a.prefetch = 0;
  if (const char* e = getenv("RAG_B200_TENSOR_PF")) a.prefetch = atoi(e);
  // lists of CTAs that never see a tile must still read as empty
  a.tau_shared = reinterpret_cast<uint32_t*>(p.scratch + L.off_tau);
  a.progress = reinterpret_cast<uint32_t*>(p.scratch + L.off_prog);
  {
    const size_t tile_bytes = static_cast<size_t>(L.nb) * width * 2;
    size_t w = kPaceBytes / (static_cast<size_t>(L.cpm) * tile_bytes);
    a.pace_window = static_cast<uint32_t>(w < 8 ? 8 : (w > 4096 ? 4096 : w));
  }
  // Off unless RAG_B200_TENSOR_PACE=<window in tiles, or 1 for the default window> is set.  Measured on B200
  // (10M x 768, B = 1024): pacing cuts the kernel's DRAM reads from 30.5 GB to 18.0 GB (corpus 15.4 GB) but
  // not its time (12.9 -> 13.0-13.2 ms: the re-reads run at a third of the HBM rate and are not the bound),
  // and at D = 384 the coupling of the sibling clusters costs 20-30 %.  Kept as an experiment switch.
  {
    const char* e = getenv("RAG_B200_TENSOR_PACE");
    const int v = e ? atoi(e) : 0;
    if (v <= 0) a.progress = nullptr;
    else if (v > 1) a.pace_window = static_cast<uint32_t>(v);
  }
  const bool filt = hi_only && kk <= 16 && p.k <= 16 && filter_enabled() && (L.mode == 0 || L.mode == 2) && L.nb == kNB;
  a.q_lo_norm2 = q_lo; a.x_max_norm2 = p.max_norm2; a.x_lo_max2 = p.lo_max2; a.guard_rel = kGuardRel;
  a.extra = reinterpret_cast<uint64_t*>(p.scratch + L.off_extra);
  a.extra_cnt = reinterpret_cast<int*>(p.scratch + L.off_extracnt);
  if (filt && (p.max_norm2 == nullptr || p.lo_max2 == nullptr)) return cudaErrorInvalidValue;
  // filter mode: the list only has to give the k-th best (the slack lives in the buffers), so it is k long, not kk
  const int klist = filt ? p.k : kk;
  a.k = klist;
  a.tau_grp = nullptr; a.grp_n = 0; a.grp_rank = 0;
  if (klist <= 16 && !(getenv("RAG_B200_TENSOR_GRP") && atoi(getenv("RAG_B200_TENSOR_GRP")) == 0)) {
    a.grp_n = std::min(std::min(L.cpm, klist), 16);
    a.grp_rank = (klist + a.grp_n - 1) / a.grp_n;
    a.tau_grp = reinterpret_cast<uint32_t*>(p.scratch + L.off_tauq);      // the quantile bounds' block (k > 16 only): stride grp_n <= cpm
  }
  a.tau_q = nullptr;
  a.tau_q_rank = (kk + L.cpm - 1) / L.cpm;
  if (kk > 16 && L.cpm <= 32 && a.tau_q_rank <= 8 && !(getenv("RAG_B200_TENSOR_TAUQ") && atoi(getenv("RAG_B200_TENSOR_TAUQ")) == 0))
    a.tau_q = reinterpret_cast<uint32_t*>(p.scratch + L.off_tauq);
  e = cudaMemsetAsync(part, 0xFF, (L.off_tauq - L.off_partial) + static_cast<size_t>(L.n_mtiles) * kM * L.cpm * 4, st);
  if (e != cudaSuccess) return e;
  dim3 grid(L.cpm * L.cl, L.n_mtiles / L.cl, 1);
  const bool l2 = (p.space == 0);
  if (filt) e = launch_filter(tmap, a, l2, L.mode, grid, st);
  else if (kk <= 16) e = launch_kl<16>(tmap, a, l2, L.mode, L.nb, grid, st);
  else if (kk <= 128) e = launch_kl<128>(tmap, a, l2, L.mode, L.nb, grid, st);
  else e = launch_kl<1024>(tmap, a, l2, L.mode, L.nb, grid, st);
  if (e != cudaSuccess) return e;
  out->partial = part;
  out->S = L.cpm;
  out->k_kept = klist;
  out->q_norm2 = q_norm;
  out->q_lo_norm2 = hi_only ? q_lo : nullptr;
  out->q_f32 = q_f32;
  out->q_exact = q_exact;
  out->merged = reinterpret_cast<uint64_t*>(p.scratch + L.off_merged);
  out->filt = filt ? 1 : 0;
  out->extra = a.extra; out->extra_cnt = a.extra_cnt; out->extra_cap = kFiltCap;
  if (launches) *launches += 2;
  return cudaSuccess;
}

int read_stats(unsigned long long* out8, int reset) {
  if (cudaMemcpyFromSymbol(out8, g_tensor_stats, 8 * sizeof(unsigned long long)) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
  if (reset) {
    const unsigned long long z[8] = {};
    if (cudaMemcpyToSymbol(g_tensor_stats, z, sizeof(z)) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
  }
  return 0;
}

}  // namespace tensor
}  // namespace rag
