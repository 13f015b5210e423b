// Strided / batched / K-blocked GEMM on CUDA cores, fp32 or fp64.
//
// This is the general-shape dense kernel of the library: it serves the fp64 validation mode, odd
// shapes, and every small contraction of the tensor models (Grams, mode products, MTTKRP on a
// materialised Khatri-Rao block).  The two X-streaming cross products and the fused multiplicative
// update of the fp32 headline path have their own tcgen05 kernels (tc_*.cu).
//
// Replaces the numpy dgemm call sites listed in include/nnfac_b200.h (nmf.py:407-408,432-433,
// mu.py:82, ntf.py:442-449, mu.py:141,159, ntd.py:672).
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename T>
struct GemmArgs {
  T* C;
  const T* A;
  const T* B;
  T* partial;  // [splits][batch][M][N] when splits > 1
  int64_t ldc, sc_b;
  int64_t sa_i, sa_k, sa_q, sa_b;
  int64_t sb_k, sb_j, sb_q, sb_b;
  int64_t M, N, K, kb, batch;
  int splits;
  int64_t kblocks_per_q;   // ceil(K / BK)
  int64_t blocks_per_split;
};

template <typename T>
__global__ void __launch_bounds__(NT) gemm_strided_kernel(GemmArgs<T> g) {
  __shared__ T As[BK][BM + 4];
  __shared__ T Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t bz = blockIdx.z;
  const int64_t b = bz / g.splits;
  const int split = (int)(bz % g.splits);
  const T* __restrict__ A = g.A + b * g.sa_b;
  const T* __restrict__ B = g.B + b * g.sb_b;

  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

  const int64_t nblk = g.kb * g.kblocks_per_q;
  const int64_t blk0 = (int64_t)split * g.blocks_per_split;
  int64_t blk1 = blk0 + g.blocks_per_split;
  if (blk1 > nblk) blk1 = nblk;
  const bool a_kcontig = (g.sa_k == 1);
  const bool b_jcontig = (g.sb_j == 1);

  for (int64_t blk = blk0; blk < blk1; ++blk) {
    const int64_t q = blk / g.kblocks_per_q;
    const int64_t k0 = (blk % g.kblocks_per_q) * BK;
    const T* Aq = A + q * g.sa_q;
    const T* Bq = B + q * g.sb_q;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int i, k;
      if (a_kcontig) { k = t & 15; i = (t >> 4) + 16 * j; }
      else           { i = t & 63; k = (t >> 6) + 4 * j; }
      const int64_t gi = m0 + i, gk = k0 + k;
      As[k][i] = (gi < g.M && gk < g.K) ? Aq[gi * g.sa_i + gk * g.sa_k] : T(0);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c, k;
      if (b_jcontig) { c = t & 63; k = (t >> 6) + 4 * j; }
      else           { k = t & 15; c = (t >> 4) + 16 * j; }
      const int64_t gc = n0 + c, gk = k0 + k;
      Bs[k][c] = (gc < g.N && gk < g.K) ? Bq[gk * g.sb_k + gc * g.sb_j] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      T a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }

  T* out;
  int64_t ld;
  if (g.splits == 1) { out = g.C + b * g.sc_b; ld = g.ldc; }
  else { out = g.partial + ((int64_t)split * g.batch + b) * g.M * g.N; ld = g.N; }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gi = m0 + ty * 4 + i;
    if (gi >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gj = n0 + tx * 4 + j;
      if (gj < g.N) out[gi * ld + gj] = acc[i][j];
    }
  }
}

template <typename T>
__global__ void splitk_reduce_kernel(T* C, int64_t ldc, int64_t sc_b, const T* partial, int64_t M,
                                     int64_t N, int64_t batch, int splits) {
  const int64_t total = batch * M * N;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    T s = T(0);
    for (int sp = 0; sp < splits; ++sp) s += partial[(int64_t)sp * total + idx];  // fixed order
    const int64_t b = idx / (M * N), rem = idx % (M * N);
    C[b * sc_b + (rem / N) * ldc + (rem % N)] = s;
  }
}

template <typename T>
int run_gemm(nnfac_ctx* ctx, T* C, int64_t ldc, int64_t sc_b, const T* A, int64_t sa_i, int64_t sa_k,
             int64_t sa_q, int64_t sa_b, const T* B, int64_t sb_k, int64_t sb_j, int64_t sb_q,
             int64_t sb_b, int64_t M, int64_t N, int64_t K, int64_t kb, int64_t batch, cudaStream_t st) {
  GemmArgs<T> g;
  g.C = C; g.A = A; g.B = B; g.ldc = ldc; g.sc_b = sc_b;
  g.sa_i = sa_i; g.sa_k = sa_k; g.sa_q = sa_q; g.sa_b = sa_b;
  g.sb_k = sb_k; g.sb_j = sb_j; g.sb_q = sb_q; g.sb_b = sb_b;
  g.M = M; g.N = N; g.K = K; g.kb = kb; g.batch = batch;
  g.kblocks_per_q = ceil_div64(K, BK);
  const int64_t nblk = kb * g.kblocks_per_q;
  const int64_t tm = ceil_div64(M, BM), tn = ceil_div64(N, BN);
  const int64_t tiles = tm * tn * batch;
  int64_t splits = 1;
  if (tiles < 2 * (int64_t)ctx->sm_count) {
    splits = ceil_div64(2 * (int64_t)ctx->sm_count, tiles);
    const int64_t max_by_work = nblk / 8 > 0 ? nblk / 8 : 1;  // at least 8 k-blocks per split
    if (splits > max_by_work) splits = max_by_work;
    if (splits > 256) splits = 256;
  }
  g.blocks_per_split = ceil_div64(nblk, splits);
  splits = ceil_div64(nblk, g.blocks_per_split);
  g.splits = (int)splits;
  g.partial = nullptr;
  if (splits > 1) {
    const size_t need = (size_t)splits * batch * M * N * sizeof(T);
    int rc = nnfac_guard_enter(ctx, NNFAC_GUARD_WS, st);
    if (!rc) rc = nnfac_ws_reserve(ctx, need, st);
    if (rc) return rc;
    g.partial = (T*)ctx->ws;
  }
  NNFAC_ARG(tm <= 65535 && batch * splits <= 65535, "nnfac_gemm_strided: grid too large (M tiles %lld, batch*splits %lld)",
            (long long)tm, (long long)(batch * splits));
  dim3 grid((unsigned)tn, (unsigned)tm, (unsigned)(batch * splits));
  gemm_strided_kernel<T><<<grid, NT, 0, st>>>(g);
  NNFAC_LAUNCH_CHECK(ctx);
  if (splits > 1) {
    const int64_t total = batch * M * N;
    int blocks = (int)(ceil_div64(total, 256) < 4096 ? ceil_div64(total, 256) : 4096);
    splitk_reduce_kernel<T><<<blocks, 256, 0, st>>>(C, ldc, sc_b, g.partial, M, N, batch, (int)splits);
    NNFAC_LAUNCH_CHECK(ctx);
  }
  return NNFAC_OK;
}

// ---- Gram of a rank-major factor: out (r x r) = F F^T, F (r x len), r <= 64 -------------------------------
// nmf.py:407 (V V^T) and nmf.py:432 (U^T U): a short-and-wide product (K = len up to 10^5..10^6, M = N = r).
// Every CTA owns a contiguous range of columns, stages 64-column slabs of F in shared memory (transposed, so
// that a thread reads its 4 + 4 operands with two 16-byte loads) and keeps a 4 x 4 block of the Gram in
// registers; the per-CTA partials are added in fixed order by a second kernel (fp64 adder).
constexpr int GR = 64, GC = 64, GPADF = 68;

template <typename T>
__global__ void __launch_bounds__(256, 4) gram_partial_kernel(const T* __restrict__ F, int64_t ld_f, int r, int64_t len,
                                                           int64_t cols_per_cta, T* __restrict__ part) {
  __shared__ __align__(16) T tile[GC][GPADF];
  const int t = threadIdx.x, ti = t >> 4, tj = t & 15;
  const int64_t c_begin = (int64_t)blockIdx.x * cols_per_cta;
  int64_t c_end = c_begin + cols_per_cta;
  if (c_end > len) c_end = len;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int64_t c0 = c_begin; c0 < c_end; c0 += GC) {
    __syncthreads();
    for (int idx = t; idx < GR * GC; idx += 256) {
      const int k = idx >> 6, cx = idx & 63;
      const int64_t c = c0 + cx;
      tile[cx][k] = (k < r && c < c_end) ? F[(int64_t)k * ld_f + c] : T(0);
    }
    __syncthreads();
#pragma unroll 8
    for (int cx = 0; cx < GC; ++cx) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = tile[cx][4 * ti + i]; b[i] = tile[cx][4 * tj + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
  }
  T* out = part + (size_t)blockIdx.x * GR * GR;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(4 * ti + i) * GR + 4 * tj + j] = acc[i][j];
}

template <typename T>
__global__ void gram_reduce_kernel(const T* __restrict__ part, int nparts, int r, T* __restrict__ out, int64_t ld_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= r * r) return;
  const int i = idx / r, j = idx % r;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += (double)part[(size_t)p * GR * GR + i * GR + j];   // fixed order
  out[(int64_t)i * ld_out + j] = (T)s;
}

template <typename T>
int run_gram(nnfac_ctx* ctx, T* out, int64_t ld_out, const T* F, int64_t ld_f, int r, int64_t len, cudaStream_t st) {
  // four CTAs per SM hide the latency of the slab loads; at most 4 * sm_count partials for the fixed-order adder
  int64_t cols = ceil_div64(len, (int64_t)ctx->sm_count * 4);
  cols = ceil_div64(cols, GC) * GC;
  const int grid = (int)ceil_div64(len, cols);
  const size_t need = (size_t)grid * GR * GR * sizeof(T);
  int rc = nnfac_guard_enter(ctx, NNFAC_GUARD_WS, st);
  if (!rc) rc = nnfac_ws_reserve(ctx, need, st);
  if (rc != NNFAC_OK) return rc;
  gram_partial_kernel<T><<<grid, 256, 0, st>>>(F, ld_f, r, len, cols, (T*)ctx->ws);
  NNFAC_LAUNCH_CHECK(ctx);
  gram_reduce_kernel<T><<<(r * r + 255) / 256, 256, 0, st>>>((const T*)ctx->ws, grid, r, out, ld_out);
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

// ---- the same for 64 < r <= 128 (fp32): the Gram as 64 x 64 blocks (0,0), (0,1), (1,1) -- blockIdx.y -- of the row blocks
// F[0:64] and F[64:128]; block (1,0) is the mirror of (0,1).  32-column slabs (17 KiB of shared memory for the two tiles), so
// that a CTA still fits beside the tcgen05 X pass it runs under.
constexpr int GC2 = 32;

__global__ void __launch_bounds__(256, 4) gram2_partial_kernel(const float* __restrict__ F, int64_t ld_f, int r, int64_t len,
                                                               int64_t cols_per_cta, float* __restrict__ part) {
  __shared__ __align__(16) float ta[GC2][GPADF], tb[GC2][GPADF];
  const int t = threadIdx.x, ti = t >> 4, tj = t & 15;
  const int pair = blockIdx.y, ra = pair == 2 ? 64 : 0, rb = pair == 0 ? 0 : 64;       // row blocks of the two operands
  const int64_t c_begin = (int64_t)blockIdx.x * cols_per_cta;
  int64_t c_end = c_begin + cols_per_cta;
  if (c_end > len) c_end = len;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t c0 = c_begin; c0 < c_end; c0 += GC2) {
    __syncthreads();
    for (int idx = t; idx < GR * GC2; idx += 256) {
      const int k = idx >> 5, cx = idx & 31;
      const int64_t c = c0 + cx;
      ta[cx][k] = (ra + k < r && c < c_end) ? F[(int64_t)(ra + k) * ld_f + c] : 0.f;
      tb[cx][k] = (rb + k < r && c < c_end) ? F[(int64_t)(rb + k) * ld_f + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int cx = 0; cx < GC2; ++cx) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = ta[cx][4 * ti + i]; b[i] = tb[cx][4 * tj + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
  }
  float* out = part + ((size_t)pair * gridDim.x + blockIdx.x) * GR * GR;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(4 * ti + i) * GR + 4 * tj + j] = acc[i][j];
}

__global__ void gram2_reduce_kernel(const float* __restrict__ part, int nparts, int r, float* __restrict__ out, int64_t ld_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= r * r) return;
  const int i = idx / r, j = idx % r;
  const int bi = i >> 6, bj = j >> 6;
  const int pair = bi + bj;                                         // (0,0) -> 0, (0,1) / (1,0) -> 1, (1,1) -> 2
  const int li = (bi <= bj ? i : j) & 63, lj = (bi <= bj ? j : i) & 63;   // block (1,0) reads block (0,1) transposed
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += (double)part[((size_t)pair * nparts + p) * GR * GR + li * GR + lj];   // fixed order
  out[(int64_t)i * ld_out + j] = (float)s;
}

int run_gram2(nnfac_ctx* ctx, float* out, int64_t ld_out, const float* F, int64_t ld_f, int r, int64_t len, cudaStream_t st) {
  int64_t cols = ceil_div64(len, (int64_t)ctx->sm_count * 2);
  cols = ceil_div64(cols, GC2) * GC2;
  const int grid = (int)ceil_div64(len, cols);
  const size_t need = (size_t)3 * grid * GR * GR * sizeof(float);
  int rc = nnfac_guard_enter(ctx, NNFAC_GUARD_WS, st);
  if (!rc) rc = nnfac_ws_reserve(ctx, need, st);
  if (rc != NNFAC_OK) return rc;
  gram2_partial_kernel<<<dim3((unsigned)grid, 3), 256, 0, st>>>(F, ld_f, r, len, cols, (float*)ctx->ws);
  NNFAC_LAUNCH_CHECK(ctx);
  gram2_reduce_kernel<<<(r * r + 255) / 256, 256, 0, st>>>((const float*)ctx->ws, grid, r, out, ld_out);
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

}  // namespace

extern "C" int nnfac_gram(nnfac_ctx* ctx, int dtype, void* out, int64_t ld_out, const void* F, int64_t ld_f, int r,
                          int64_t len, void* stream) {
  NNFAC_ARG(ctx && out && F && r > 0 && len > 0 && ld_out >= r && ld_f >= len, "nnfac_gram: bad argument");
  NNFAC_ARG(dtype == NNFAC_F32 || dtype == NNFAC_F64, "nnfac_gram: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (r > GR && r <= 2 * GR && dtype == NNFAC_F32) return run_gram2(ctx, (float*)out, ld_out, (const float*)F, ld_f, r, len, st);
  if (r > GR) {   // wider factors go through the general kernel
    if (dtype == NNFAC_F32)
      return run_gemm<float>(ctx, (float*)out, ld_out, 0, (const float*)F, ld_f, 1, 0, 0, (const float*)F, 1, ld_f, 0, 0, r, r, len, 1, 1, st);
    return run_gemm<double>(ctx, (double*)out, ld_out, 0, (const double*)F, ld_f, 1, 0, 0, (const double*)F, 1, ld_f, 0, 0, r, r, len, 1, 1, st);
  }
  if (dtype == NNFAC_F32) return run_gram<float>(ctx, (float*)out, ld_out, (const float*)F, ld_f, r, len, st);
  return run_gram<double>(ctx, (double*)out, ld_out, (const double*)F, ld_f, r, len, st);
}

extern "C" int nnfac_gemm_strided(nnfac_ctx* ctx, int dtype, void* C, int64_t ldc, int64_t sc_b,
                                  const void* A, int64_t sa_i, int64_t sa_k, int64_t sa_q,
                                  int64_t sa_b, const void* B, int64_t sb_k, int64_t sb_j,
                                  int64_t sb_q, int64_t sb_b, int64_t M, int64_t N, int64_t K,
                                  int64_t kb, int64_t batch, void* stream) {
  NNFAC_ARG(ctx && C && A && B, "nnfac_gemm_strided: NULL argument");
  NNFAC_ARG(M > 0 && N > 0 && K > 0 && kb > 0 && batch > 0, "nnfac_gemm_strided: empty dimension");
  NNFAC_ARG(dtype == NNFAC_F32 || dtype == NNFAC_F64, "nnfac_gemm_strided: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == NNFAC_F32)
    return run_gemm<float>(ctx, (float*)C, ldc, sc_b, (const float*)A, sa_i, sa_k, sa_q, sa_b,
                           (const float*)B, sb_k, sb_j, sb_q, sb_b, M, N, K, kb, batch, st);
  return run_gemm<double>(ctx, (double*)C, ldc, sc_b, (const double*)A, sa_i, sa_k, sa_q, sa_b,
                          (const double*)B, sb_k, sb_j, sb_q, sb_b, M, N, K, kb, batch, st);
}
