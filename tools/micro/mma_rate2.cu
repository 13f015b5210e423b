// Micro-benchmark 2: tcgen05.mma (kind::f16, bf16, M=128, K=16) issue rate with a fully unrolled issue loop
// (no address arithmetic between MMAs), and the issue -> commit -> mbarrier round trip of a short batch.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate2 mma_rate2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../nn-fac_b200/csrc/tc_common.cuh"

template <int N, int BATCH, bool TS>
__global__ void __launch_bounds__(128, 1) kern(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
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
    uint32_t phase = 0;
    // (1) back-to-back issue of iters*BATCH MMAs, one commit at the end
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        if (TS) tc::umma_bf16_ts(tm, tm + 448 + (j & 3) * 8, b + (uint64_t)((j & 3) * 2), idesc, true);
        else tc::umma_bf16(tm, a + (uint64_t)((j & 3) * 2), b + (uint64_t)((j & 3) * 2), idesc, true);
      }
    }
    long long t1 = clock64();
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, phase); phase ^= 1;
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
    // (2) round trips: BATCH MMAs, commit, wait
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        if (TS) tc::umma_bf16_ts(tm, tm + 448 + (j & 3) * 8, b + (uint64_t)((j & 3) * 2), idesc, true);
        else tc::umma_bf16(tm, a + (uint64_t)((j & 3) * 2), b + (uint64_t)((j & 3) * 2), idesc, true);
      }
      tc::umma_commit(&bar);
      tc::mbar_wait(&bar, phase); phase ^= 1;
    }
    t1 = clock64();
    out[2] = t1 - t0;
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 512);
}

template <int N, int BATCH, bool TS>
void run(long long* out) {
  const int iters = 512;
  cudaFuncSetAttribute(kern<N, BATCH, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  kern<N, BATCH, TS><<<1, 128, 64 * 1024>>>(iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("A=%s N=%3d batch=%2d : issue %.1f clk/mma, complete %.1f clk/mma, round trip %.0f clk per batch\n", TS ? "tmem" : "smem", N,
         BATCH, (double)out[0] / (iters * BATCH), (double)out[1] / (iters * BATCH), (double)out[2] / iters);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  run<16, 1, false>(out); run<16, 4, false>(out); run<16, 12, false>(out);
  run<48, 1, false>(out); run<48, 4, false>(out); run<48, 12, false>(out);
  run<64, 1, false>(out); run<64, 4, false>(out); run<64, 12, false>(out); run<64, 12, true>(out);
  run<128, 4, false>(out); run<128, 12, false>(out);
  run<256, 4, false>(out); run<256, 12, false>(out); run<256, 12, true>(out);
  return 0;
}
