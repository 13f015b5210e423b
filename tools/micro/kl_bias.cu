// Where does the bias of sum x*log(x/k) come from: rcp.approx, lg2.approx, or both?  (run on the GPU box)
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__global__ void k(const float* x, const float* kk, int n, double* out) {
  double s[5] = {0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float a = x[i], b = kk[i];
    const float i0 = rcpa(b);
    const float i1 = fmaf(i0, fmaf(-b, i0, 1.f), i0);
    s[0] += (double)(a * lg2a(a * i0));                  // current kernel
    s[1] += (double)(a * lg2a(a * i1));                  // refined reciprocal
    s[2] += (double)(a * lg2a(a / b));                   // IEEE division
    s[3] += (double)(a * log2f(a / b));                  // accurate log2f
    const float q = a * i1, d = (a - b) * i1;            // log1p-style: lg2(q) replaced near 1 by a short series in d
    const float ser = d * (1.f - d * (0.5f - d * (1.f / 3.f - d * 0.25f)));
    s[4] += (double)(a * (fabsf(d) < 0.03125f ? ser * 1.4426950408889634f : lg2a(q)));
  }
  for (int j = 0; j < 5; ++j) atomicAdd(&out[j], s[j]);
}
int main() {
  const int n = 1 << 24;
  std::vector<float> x(n), kk(n);
  double ref = 0;
  unsigned long long st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
  for (int amp = 0; amp < 2; ++amp) {
    const double A = amp == 0 ? 0.1 : 0.5;
    ref = 0;
    for (int i = 0; i < n; ++i) {
      kk[i] = (float)(1000.0 + 2000.0 * rnd());
      x[i] = (float)(kk[i] * (1.0 + A * (2.0 * rnd() - 1.0)));
      ref += (double)x[i] * std::log2((double)x[i] / (double)kk[i]);
    }
    float *dx, *dk; double* dout;
    cudaMalloc(&dx, n * 4); cudaMalloc(&dk, n * 4); cudaMalloc(&dout, 40); cudaMemset(dout, 0, 40);
    cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dk, kk.data(), n * 4, cudaMemcpyHostToDevice);
    k<<<592, 256>>>(dx, dk, n, dout);
    double h[5]; cudaMemcpy(h, dout, 40, cudaMemcpyDeviceToHost);
    double sx = 0; for (int i = 0; i < n; ++i) sx += x[i];
    printf("|d| <= %.1f: ref %.6e  sum x %.3e\n", A, ref, sx);
    const char* name[5] = {"rcp.approx+lg2.approx", "refined rcp+lg2.approx", "div+lg2.approx", "div+log2f", "refined rcp + series near 1"};
    for (int j = 0; j < 5; ++j) printf("  %-28s err %+.3e  (%.2e of sum x, %.2e of ref)\n", name[j], h[j] - ref, (h[j] - ref) / sx, (h[j] - ref) / ref);
    cudaFree(dx); cudaFree(dk); cudaFree(dout);
  }
  return 0;
}
