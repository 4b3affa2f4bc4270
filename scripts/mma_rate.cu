// Micro-benchmark: clocks per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N and of where A comes from (shared
// memory descriptor or tensor memory), one issuing warp per CTA, one CTA per SM, operands resident in shared memory.
// It answers what bounds the attention-linearisation kernels, whose MMAs are small (N = 48 / 64):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I diffusion_pullback_b200/csrc -o gpurun_out/mma_rate scripts/mma_rate.cu && gpurun_out/mma_rate
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "pb_tc.cuh"
using namespace pbtc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

// mode 0: A and B from shared memory; 1: A from tensor memory; nacc: accumulators the chain rotates over (independent chains)
__global__ void __launch_bounds__(64, 1) rate_kernel(int N, int mode, int iters, int nacc, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (threadIdx.x < 32) {
    const uint32_t idesc = (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint64_t adesc = make_smem_desc(smem_u32(smem)), bdesc = make_smem_desc(smem_u32(smem) + 16384);
    const uint32_t t = tbase;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t d = t + 256 + uint32_t((i * 4 + k) & (nacc - 1)) * 64u;
          if (mode == 0) mma_f16(d, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, 1u);
          else mma_ts(d, t + 8 * k, bdesc + uint64_t(2 * k), idesc, 1u);
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one()) tcgen05_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
  }
}

// CTA pair: M = 256 over two SMs (cta_group::2), leader issues; each CTA holds 128 rows of A and N / 2 rows of B
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1) rate2_kernel(int N, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  if (threadIdx.x < 32) {
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(256 >> 4) << 24);
      const uint64_t adesc = make_smem_desc(smem_u32(smem)), bdesc = make_smem_desc(smem_u32(smem) + 16384);
      const uint32_t t = tbase;
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_f16_2sm(t + 256, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, 1u);
        }
        __syncwarp();
      }
      long long t1 = clock64();
      if (elect_one()) tcgen05_commit_2sm(&bar);
      __syncwarp();
      mbar_wait(&bar, 0);
      long long t2 = clock64();
      if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else {
      mbar_wait(&bar, 0);
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  printf("tcgen05.mma kind::f16 M=128 K=16, %d MMAs back to back per SM, 148 CTAs: clocks per MMA (issue loop | until the last retires)\n", 4 * iters);
  const int Ns[] = {16, 32, 48, 64, 96, 128, 160, 240, 256};
  for (int mode = 0; mode < 2; ++mode)
    for (int nacc = 1; nacc <= 4; nacc += 3)
      for (int N : Ns) {
        if (nacc * 64 < N && nacc > 1) continue;
        if (nacc == 4 && N > 64) continue;
        rate_kernel<<<148, 64, smem>>>(N, mode, iters, nacc, d);
        rate_kernel<<<148, 64, smem>>>(N, mode, iters, nacc, d);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("A from %s  N=%3d  accumulators %d:  %6.1f | %6.1f clk/MMA   (peak-rate math %5.1f)\n", mode ? "TMEM" : "smem", N, nacc,
               double(h[0]) / (4 * iters), double(h[1]) / (4 * iters), 128.0 * N * 16 * 2 / 8192);
      }
  cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int N : {32, 48, 64, 96, 128, 160, 256}) {
    rate2_kernel<<<148, 64, smem>>>(N, iters, d);
    rate2_kernel<<<148, 64, smem>>>(N, iters, d);
    long long h[2];
    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    printf("CTA pair (M=256 over two SMs, A and B from smem)  N=%3d:  %6.1f | %6.1f clk/MMA   (peak-rate math per SM %5.1f)\n", N,
           double(h[0]) / (4 * iters), double(h[1]) / (4 * iters), 128.0 * N * 16 * 2 / 8192);
  }
  return 0;
}
