// Microbenchmark: what does the tensor-regime MAINLOOP schedule cost per tcgen05.mma, with no TMA, no
// epilogue and no data dependencies?  One CTA per SM issues M = 128, N = 64, K = 16 bf16 MMAs (A from
// TMEM) the way gemm_topk_kernel does: `tile_mmas` MMAs per tile into one of `nbuf` accumulator
// buffers (first MMA of a tile overwrites), a tcgen05.commit every `commit_every` MMAs (to a barrier
// nobody waits on), the B descriptor cycling over `stages` shared-memory stages, optionally while the
// other warps keep writing shared memory (a stand-in for the TMA fill).   nvcc -arch=sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct P { int tiles, nbuf, commit, writers; };

// TILE_STAGES x 16 MMAs per tile (24 MMAs: second stage half full, as D = 384 in the kernel), fully unrolled
// issue with immediate descriptor offsets -- the same instruction stream as gemm_topk_kernel's MMA warp
template <int TILE_MMAS>
__global__ void __launch_bounds__(192, 1) sched_kernel(P p, long long* out_cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ volatile int stop;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr int kStagesPerTile = (TILE_MMAS + 15) / 16;
  constexpr int kRing = 6;
  if (warp == 1) {
    const uint32_t acc0 = tmem + 512 - p.nbuf * 64;
    const uint64_t desc0 = make_desc(base);
    const long long t0 = clock64();
    uint32_t s = 0;
    for (int t = 0; t < p.tiles; ++t) {
      const uint32_t d = acc0 + (t & (p.nbuf - 1)) * 64;
      for (int ks = 0; ks < kStagesPerTile; ++ks) {
        if (elect_one()) {
          const uint64_t dstage = desc0 + (uint64_t)((s * 32768u) >> 4);
          const uint32_t a0 = tmem + ks * 128;
          const int left = TILE_MMAS - ks * 16;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < left)
              asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, q;\n\t}"
                           ::"r"(d), "r"(a0 + j * 8), "l"(dstage + (uint64_t)(((j >> 2) * 8192 + (j & 3) * 32) >> 4)), "r"(idesc),
                             "r"((j > 0) ? 1u : (ks > 0 ? 1u : 0u)) : "memory");
          if (p.commit) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
            if (ks == kStagesPerTile - 1 && p.commit > 1)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
          }
        }
        __syncwarp();
        if (++s == kRing) s = 0;
      }
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    __syncwarp();
    asm volatile(
        "{\n\t.reg .pred q;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], 0;\n\t@q bra D1;\n\tbra W1;\n\tD1:\n\t}"
        ::"r"(smem_u32(&bar[0])) : "memory");
    const long long t1 = clock64();
    stop = 1;
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
  } else if (warp >= 2 && warp < 2 + p.writers) {
    // stand-in for the TMA fill: stream 16-byte stores over the stage ring until the MMA warp is done
    uint4* dst = reinterpret_cast<uint4*>(smem + (base - smem_u32(smem)));
    const int lanes = p.writers * 32;
    uint32_t i = (warp - 2) * 32 + (threadIdx.x & 31);
    const uint32_t words = kRing * 32768 / 16;
    while (!stop) {
#pragma unroll
      for (int u = 0; u < 8; ++u) { dst[i % words] = make_uint4(i, i, i, i); i += lanes; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int TILE_MMAS>
static void run(P p, int grid) {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 6 * 32768 + 2048;
  cudaFuncSetAttribute(sched_kernel<TILE_MMAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  P warm = p; warm.tiles = 50;
  sched_kernel<TILE_MMAS><<<grid, 192, smem>>>(warm, d);
  cudaDeviceSynchronize();
  sched_kernel<TILE_MMAS><<<grid, 192, smem>>>(p, d);
  cudaError_t err = cudaDeviceSynchronize();
  long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  const double mmas = (double)p.tiles * TILE_MMAS;
  printf("tile_mmas=%2d nbuf=%d commits=%d writers=%d : %6.1f cycles/MMA (%s)\n", TILE_MMAS, p.nbuf, p.commit, p.writers,
         cyc / mmas, cudaGetErrorString(err));
  cudaFree(d);
}

template <int TILE_MMAS>
static void sweep() {
  const int grid = 148, tiles = 20000;
  run<TILE_MMAS>({tiles, 1, 0, 0}, grid);      // one accumulator, no commits: the plain issue rate of this instruction stream
  run<TILE_MMAS>({tiles, 2, 0, 0}, grid);      // accumulators rotate
  run<TILE_MMAS>({tiles, 4, 0, 0}, grid);
  run<TILE_MMAS>({tiles, 4, 1, 0}, grid);      // + one commit per stage
  run<TILE_MMAS>({tiles, 4, 2, 0}, grid);      // + the accumulator commit at the end of a tile
  run<TILE_MMAS>({tiles, 4, 2, 2}, grid);      // + shared-memory writes in the background (2 / 4 warps)
  run<TILE_MMAS>({tiles, 4, 2, 4}, grid);
}

int main() {
  sweep<16>();
  sweep<24>();
  sweep<32>();
  sweep<48>();
  return 0;
}
