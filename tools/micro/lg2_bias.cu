// Mean error of lg2.approx.ftz.f32 per bucket of its argument (run on the GPU box): is the bias a constant?
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__global__ void k(float lo, float hi, int n, double* out) {
  double s = 0, s2 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float q = lo + (hi - lo) * ((i + 0.37f) / n);
    const double e = (double)lg2a(q) - log2((double)q);
    s += e; s2 += e * e;
  }
  atomicAdd(&out[0], s); atomicAdd(&out[1], s2);
}
int main() {
  const float edges[] = {0.01f, 0.1f, 0.25f, 0.5f, 0.6f, 0.7f, 0.8f, 0.9f, 0.95f, 0.99f, 1.0f, 1.01f, 1.05f, 1.1f, 1.2f, 1.4f, 1.7f, 2.0f, 4.0f, 10.f, 100.f};
  const int n = 1 << 22;
  double* d; cudaMalloc(&d, 16);
  for (int b = 0; b + 1 < (int)(sizeof(edges) / sizeof(float)); ++b) {
    cudaMemset(d, 0, 16);
    k<<<296, 256>>>(edges[b], edges[b + 1], n, d);
    double h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("q in [%6.2f, %6.2f): mean err %+.3e  rms %.3e\n", edges[b], edges[b + 1], h[0] / n, sqrt(h[1] / n));
  }
  return 0;
}
