// Counter-based synthetic data ("measurement"): every element of a synthetic matrix is a pure function of
// (seed, stream, global row, global column), so any shard on any number of GPUs regenerates exactly the same data
// without generating -- or communicating -- anything else.  Philox4x32-10 (Salmon et al., SC'11): counter =
// (row, column, stream, 0), key = (seed low, seed high); the first output word gives a uniform in [0, 1) with 24 bits.
// Not on the factorisation path: bench.py and the tests use it to build X, U0 and V0 blocks on the device.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t philox_first_word(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return c0;
}

__global__ void __launch_bounds__(256) philox_uniform_kernel(float* __restrict__ out, int64_t ld, int64_t rows, int64_t cols, int64_t row0,
                                                             int64_t col0, uint32_t k0, uint32_t k1, uint32_t stream_id, float scale,
                                                             int accumulate) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const uint32_t w = philox_first_word((uint32_t)(row0 + r), (uint32_t)(col0 + c), stream_id, 0u, k0, k1);
    const float u = (float)(w >> 8) * (1.0f / 16777216.0f);
    float* p = out + r * ld + c;
    *p = accumulate ? *p + scale * u : scale * u;
  }
}

}  // namespace

extern "C" int nnfac_philox_uniform(nnfac_ctx* ctx, float* out, int64_t ld, int64_t rows, int64_t cols, int64_t row0, int64_t col0,
                                    uint64_t seed, uint32_t stream_id, double scale, int accumulate, void* stream) {
  NNFAC_ARG(ctx && out && rows > 0 && cols > 0 && ld >= cols && row0 >= 0 && col0 >= 0, "nnfac_philox_uniform: bad argument");
  NNFAC_ARG(row0 + rows <= 0xffffffffll && col0 + cols <= 0xffffffffll, "nnfac_philox_uniform: index beyond 2^32");
  const int64_t total = rows * cols;
  const int64_t want = ceil_div64(total, 256 * 4);
  const int grid = (int)(want < (int64_t)ctx->sm_count * 16 ? (want < 1 ? 1 : want) : (int64_t)ctx->sm_count * 16);
  philox_uniform_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, ld, rows, cols, row0, col0, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                                stream_id, (float)scale, accumulate);
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}
