// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16, M=128, K=16) for several N, A from smem or TMEM,
// one issuing thread, accumulators rotated over `nacc` TMEM regions.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include "../../nn-fac_b200/csrc/tc_common.cuh"

__global__ void __launch_bounds__(128, 1) mma_rate(int N, int a_tmem, int nacc, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  if (threadIdx.x == 32) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  tc::fence_proxy_async_smem();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 32) {
    const uint32_t idesc = tc::umma_idesc_bf16(128, N);
    const uint64_t a = tc::umma_desc_k_sw128(tc::smem_u32(smem)), b = tc::umma_desc_k_sw128(tc::smem_u32(smem) + 32768);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tm + (uint32_t)((i % nacc) * N);
      if (a_tmem) tc::umma_bf16_ts(d, tm + 448 + (i & 3) * 8, b + (uint64_t)((i & 3) * 2), idesc, true);
      else tc::umma_bf16(d, a + (uint64_t)((i & 3) * 2), b + (uint64_t)((i & 3) * 2), idesc, true);
    }
    const long long t1 = clock64();
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 512);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4096;
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int nacc = 1; nacc <= 2; ++nacc)
      for (int N : {16, 32, 48, 64, 128, 256}) {
        if (N * nacc > 448) continue;
        for (int grid : {1, 148}) {
          mma_rate<<<grid, 128, 64 * 1024>>>(N, a_tmem, nacc, iters, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("A=%s nacc=%d N=%3d grid=%3d : issue %.1f clk/mma, complete %.1f clk/mma\n", a_tmem ? "tmem" : "smem", nacc, N,
                 grid, (double)out[0] / iters, (double)out[1] / iters);
        }
      }
  return 0;
}
