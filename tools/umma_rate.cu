// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16 -> fp32, M = 128) for
// N in {64,128,256}, A operand from TMEM (TS) or shared memory (SS).  One CTA per SM.
// Used to size the tensor-regime kernel's tile.  nvcc -arch=sm_100a.
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

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out_cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint64_t bdesc = make_desc(base + (j & 3) * 32 + (j >> 2) * 16384);
          if (TS) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(tmem + 256), "r"(tmem + j * 8), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
          } else {
            const uint64_t adesc = make_desc(base + 65536 + (j & 3) * 32 + (j >> 2) * 16384);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem + 256), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
          }
        }
      }
      __syncwarp();
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    asm volatile(
        "{\n\t.reg .pred p;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D1;\n\tbra W1;\n\tD1:\n\t}"
        ::"r"(smem_u32(&bar)) : "memory");
    t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, bool TS>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 160 * 1024;
  cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4000;
  rate_kernel<N, TS><<<grid, 128, smem>>>(100, d);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<N, TS><<<grid, 128, smem>>>(iters, d);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
  const double mmas = (double)iters * 8;
  const double flops = mmas * 2.0 * 128 * N * 16 * grid;
  printf("%-10s N=%3d grid=%3d: %7.1f cycles/MMA  %8.1f TFLOP/s  (%s)\n", name, N, grid, cyc / mmas, flops / (ms * 1e-3) / 1e12,
         cudaGetErrorString(err));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<64, true>("TS", grid);  run<128, true>("TS", grid);  run<256, true>("TS", grid);
    run<64, false>("SS", grid); run<128, false>("SS", grid); run<256, false>("SS", grid);
  }
  return 0;
}
