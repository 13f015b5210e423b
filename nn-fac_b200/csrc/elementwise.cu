// Element-wise terms, reductions and small layout kernels of the factor-update path.
// Reference call sites: mu.py:84-97,143-159 (ratio / power terms and the multiplicative apply),
// beta_divergence.py:42-52, nmf.py:452 (Frobenius term, matrix 1-norm), ntf.py:445,448,466,470,
// ntd.py:676-681.
#include "common.cuh"

namespace {

constexpr int RB = 256;        // reduction block size
constexpr int RMAX_BLOCKS = 1184;  // 8 x 148: two-stage reductions use at most this many partials

template <typename T> __device__ __forceinline__ T t_pow(T a, T b);
template <> __device__ __forceinline__ float t_pow<float>(float a, float b) { return powf(a, b); }
template <> __device__ __forceinline__ double t_pow<double>(double a, double b) { return pow(a, b); }
template <typename T> __device__ __forceinline__ T t_log(T a);
template <> __device__ __forceinline__ float t_log<float>(float a) { return logf(a); }
template <> __device__ __forceinline__ double t_log<double>(double a) { return log(a); }

// ---- mu terms --------------------------------------------------------------------------------
// mode: 0 beta==1 (P = X / K), 1 beta==2 (P = X, Q = K), 2 beta==3 (P = K X, Q = K^2), 3 general
template <typename T>
__global__ void mu_terms_kernel(int mode, T beta, const T* __restrict__ K, const T* __restrict__ X,
                                T* P, T* Q, int64_t count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const T k = K[i];
    const T x = X ? X[i] : T(1);
    T p, q;
    if (mode == 0) { p = x / k; q = T(1); }
    else if (mode == 1) { p = x; q = k; }
    else if (mode == 2) { p = k * x; q = k * k; }
    else { p = t_pow<T>(k, beta - T(2)) * x; q = t_pow<T>(k, beta - T(1)); }
    if (P) P[i] = p;
    if (Q) Q[i] = q;
  }
}

template <typename T>
__global__ void mu_apply_kernel(T* out, const T* __restrict__ F, const T* __restrict__ num,
                                const T* __restrict__ den_mat, const T* __restrict__ den_vec,
                                int vec_per_row, int64_t rows, int64_t cols, T gamma, T floor_value) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const T d = den_vec ? den_vec[vec_per_row ? i / cols : i % cols] : den_mat[i];
    T ratio = num[i] / d;
    if (gamma != T(1)) ratio = t_pow<T>(ratio, gamma);
    const T v = F[i] * ratio;
    out[i] = v > floor_value ? v : floor_value;   // np.maximum(., epsilon); NaN propagates like numpy
    if (v != v) out[i] = v;
  }
}

// ---- reductions --------------------------------------------------------------------------------
// op: 0 beta-div beta==1, 1 beta==0, 2 general beta, 3 squared difference, 4 dot, 5 sum of squares
template <typename T>
__device__ __forceinline__ double red_term(int op, double beta, const T* A, const T* B, int64_t i) {
  const double a = (double)A[i];
  if (op == 5) return a * a;
  const double b = (double)B[i];
  switch (op) {
    // beta_divergence.py:45-50 masks the logarithm where its argument is 0 (`where=`): an entry with a == 0 contributes
    // b (beta = 1) or a / b - 1 (beta = 0), not 0 * log(0) = NaN / +inf -- sparse and count data have exact zeros
    case 0: return (a != 0.0 ? a * log(a / b) : 0.0) - a + b;
    case 1: { const double q = a / b; return q - (a != 0.0 ? log(q) : 0.0) - 1.0; }
    case 2: return (pow(a, beta) + (beta - 1.0) * pow(b, beta) - beta * a * pow(b, beta - 1.0)) / (beta * (beta - 1.0));
    case 3: { const double d = a - b; return d * d; }
    default: return a * b;
  }
}

template <typename T>
__global__ void __launch_bounds__(RB) reduce_stage1(int op, double beta, const T* __restrict__ A,
                                                    const T* __restrict__ B, int64_t count, double* part) {
  __shared__ double sh[33];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < count; i += (int64_t)gridDim.x * RB)
    s += red_term<T>(op, beta, A, B, i);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void __launch_bounds__(RB) reduce_stage2(const double* part, int nparts, double* out, int is_max) {
  __shared__ double sh[33];
  double s = is_max ? -1.0e300 : 0.0;
  for (int i = threadIdx.x; i < nparts; i += RB) s = is_max ? fmax(s, part[i]) : s + part[i];
  s = is_max ? block_max(s, sh) : block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = s;
}

template <typename T>
int run_reduce(nnfac_ctx* ctx, int op, double beta, const T* A, const T* B, int64_t count, double* out,
               cudaStream_t st) {
  int blocks = (int)(ceil_div64(count, RB * 4) < RMAX_BLOCKS ? ceil_div64(count, RB * 4) : RMAX_BLOCKS);
  if (blocks < 1) blocks = 1;
  const int grc = nnfac_guard_enter(ctx, NNFAC_GUARD_RED, st);
  if (grc) return grc;
  reduce_stage1<T><<<blocks, RB, 0, st>>>(op, beta, A, B, count, ctx->red);
  NNFAC_LAUNCH_CHECK(ctx);
  reduce_stage2<<<1, RB, 0, st>>>(ctx->red, blocks, out, 0);
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

// row sums: one warp per row
template <typename T>
__global__ void row_sums_kernel(const T* __restrict__ A, int64_t lda, int64_t rows, int64_t cols, T* out) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  double s = 0.0;
  for (int64_t j = threadIdx.x & 31; j < cols; j += 32) s += (double)A[row * lda + j];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) out[row] = (T)s;
}

// long rows: one CTA per row, fixed-order block reduction
template <typename T>
__global__ void __launch_bounds__(256) row_sums_wide_kernel(const T* __restrict__ A, int64_t lda, int64_t cols, T* out) {
  __shared__ double sh[33];
  const T* row = A + (int64_t)blockIdx.x * lda;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < cols; j += 256) s += (double)row[j];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = (T)s;
}

// column abs sums -> per-block maxima (matrix 1-norm)
template <typename T>
__global__ void __launch_bounds__(RB) norm1_stage1(const T* __restrict__ A, int64_t lda, int64_t rows,
                                                   int64_t cols, double* part) {
  __shared__ double sh[33];
  double best = 0.0;
  for (int64_t j = (int64_t)blockIdx.x * RB + threadIdx.x; j < cols; j += (int64_t)gridDim.x * RB) {
    double s = 0.0;
    for (int64_t i = 0; i < rows; ++i) s += fabs((double)A[i * lda + j]);
    best = fmax(best, s);
  }
  best = block_max(best, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = best;
}

template <typename T>
__global__ void transpose_kernel(T* out, int64_t ld_out, const T* __restrict__ in, int64_t ld_in,
                                 int64_t rows, int64_t cols) {
  __shared__ T tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[r * ld_in + c] : T(0);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * ld_out + r] = tile[threadIdx.x][i];
  }
}

template <typename T>
__global__ void khatri_rao_kernel(T* out, const T* __restrict__ A, int64_t I, const T* __restrict__ B,
                                  int64_t J, int64_t r) {
  const int64_t total = I * J * r;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = idx % r, ij = idx / r;
    out[idx] = A[(ij / J) * r + q] * B[(ij % J) * r + q];
  }
}

template <typename T>
__global__ void hadamard_kernel(T* out, const T* A, const T* B, int64_t count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = A[i] * B[i];
}

// out = a X + b Y (the coupled right-hand side UtM + mu Vtarget of nnls.py:318, and the extra-row correction of nnls.py:163)
template <typename T>
__global__ void axpby_kernel(T* out, T a, const T* X, T b, const T* Y, int64_t count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a * X[i] + b * Y[i];
}

// one block per row: row /= ||row||_2 when the norm is non-zero (ntd.py:678-680)
template <typename T>
__global__ void __launch_bounds__(RB) normalize_rows_kernel(T* A, int64_t lda, int64_t cols) {
  __shared__ double sh[33];
  T* row = A + (int64_t)blockIdx.x * lda;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < cols; j += RB) s += (double)row[j] * (double)row[j];
  s = block_sum(s, sh);
  const double nrm = sqrt(s);
  if (nrm != 0.0)
    for (int64_t j = threadIdx.x; j < cols; j += RB) row[j] = (T)((double)row[j] / nrm);
}

inline int grid_for(int64_t count, int sm) {
  int64_t b = ceil_div64(count, 256);
  const int64_t cap = (int64_t)sm * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---- projected-gradient step on the Tucker core (ntd.py:607-617) -------------------------------------------
// state = {upd_0, upd, cnt, done}: the loop `while cnt <= 300 and upd >= delta * upd_0` runs on the device; once
// `done` is set the remaining launches of a batch return at once (the host looks at the flag only every few steps).
template <typename T>
__global__ void __launch_bounds__(RB) core_pg_step_kernel(T* __restrict__ core, const T* __restrict__ MtX, const T* __restrict__ P,
                                                          int64_t count, T step, T sparse, int dev_scalars, const double* state,
                                                          double* part) {
  __shared__ double sh[33];
  if (state[3] != 0.0) return;
  if (dev_scalars) {                                     // step and sparsity from state[4], state[5] (graph-replayable launch)
    step = (T)state[4];
    sparse = (T)state[5];
  }
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < count; i += (int64_t)gridDim.x * RB) {
    const T g = -MtX[i] + P[i] + sparse;                 // gradient, ntd.py:608
    const T c = core[i];
    const T st = step * g;
    const T d = st < c ? st : c;                         // np.minimum(gradient_step * gradient, core)
    core[i] = c - d;
    s += (double)d * (double)d;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void core_pg_finish_kernel(const double* part, int nparts, double delta, double* state) {
  if (threadIdx.x != 0 || state[3] != 0.0) return;
  double s = 0.0;
  for (int i = 0; i < nparts; ++i) s += part[i];         // fixed order
  const double upd = sqrt(s);
  if (state[2] == 1.0) state[0] = upd;                   // upd_0, ntd.py:614-615
  state[1] = upd;
  state[2] += 1.0;
  if (!(state[2] <= 300.0 && upd >= delta * state[0])) state[3] = 1.0;
  // A first step that moves nothing (typically: the reference's round(step, 6) gave a step of 0, ntd.py:594) leaves the
  // loop condition `0 >= delta * 0` true, and every further step repeats the same no-op on the same core until cnt
  // exceeds 300: those steps are counted, not executed.
  if (state[2] == 2.0 && upd == 0.0) {
    state[2] = 301.0;
    state[3] = 1.0;
  }
}

// ---- projected-gradient step on a 3-way Tucker core, every rank <= 64 (ntd.py:607-617) -----------------------------
// P = core x_0 M0 x_1 M1 x_2 M2 (ntd.py:610, the M are the r x r Grams of the factors) in two kernels instead of three
// generic strided GEMMs of ~9 us each: (1) one CTA per slab i contracts modes 2 and 1 in shared memory,
// Z[i] = M1 (core[i] M2^T); (2) one thread per column (b, c) contracts mode 0, P[:, col] = M0 Z[:, col], and applies
// the step to its r0 elements at once (gradient, clamp, squared-step partial).  The step is ~3 MFLOP on 128 KB: what
// matters is the number of launches per step, because the loop runs up to 300 steps per outer iteration.
template <typename T>
__global__ void __launch_bounds__(256) core_modes12_kernel(const T* __restrict__ core, const T* __restrict__ M1,
                                                           const T* __restrict__ M2, T* __restrict__ Z, int r1, int r2,
                                                           const double* state) {
  if (state[3] != 0.0) return;
  extern __shared__ __align__(16) unsigned char pg_smem[];
  T* Gs = reinterpret_cast<T*>(pg_smem);        // [r1][r2]
  T* Ys = Gs + r1 * r2;                         // [r1][r2]
  T* M1s = Ys + r1 * r2;                        // [r1][r1]
  T* M2s = M1s + r1 * r1;                       // [r2][r2 + 1]  (padded: lanes read a column)
  const int i = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  for (int x = tid; x < r1 * r2; x += nt) Gs[x] = core[(int64_t)i * r1 * r2 + x];
  for (int x = tid; x < r1 * r1; x += nt) M1s[x] = M1[x];
  for (int x = tid; x < r2 * r2; x += nt) M2s[(x / r2) * (r2 + 1) + x % r2] = M2[x];
  __syncthreads();
  for (int x = tid; x < r1 * r2; x += nt) {     // Y[j][c] = sum_k core[i][j][k] M2[c][k]
    const int j = x / r2, c = x % r2;
    T acc = 0;
    for (int k = 0; k < r2; ++k) acc += Gs[j * r2 + k] * M2s[c * (r2 + 1) + k];
    Ys[x] = acc;
  }
  __syncthreads();
  for (int x = tid; x < r1 * r2; x += nt) {     // Z[i][b][c] = sum_j M1[b][j] Y[j][c]
    const int b = x / r2, c = x % r2;
    T acc = 0;
    for (int j = 0; j < r1; ++j) acc += M1s[b * r1 + j] * Ys[j * r2 + c];
    Z[(int64_t)i * r1 * r2 + x] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(128) core_mode0_step_kernel(T* __restrict__ core, const T* __restrict__ MtX,
                                                              const T* __restrict__ Z, const T* __restrict__ M0, int r0, int ncol,
                                                              T step, T sparse, int dev_scalars, const double* state,
                                                              double* part) {
  __shared__ double sh[33];
  extern __shared__ __align__(16) unsigned char pg_smem[];
  if (state[3] != 0.0) return;
  if (dev_scalars) {
    step = (T)state[4];
    sparse = (T)state[5];
  }
  T* M0s = reinterpret_cast<T*>(pg_smem);       // [r0][r0]
  T* Zs = M0s + r0 * r0;                        // [r0][128]
  const int tid = threadIdx.x, col = blockIdx.x * 128 + tid;
  for (int x = tid; x < r0 * r0; x += 128) M0s[x] = M0[x];
  if (col < ncol)
    for (int i = 0; i < r0; ++i) Zs[i * 128 + tid] = Z[(int64_t)i * ncol + col];
  __syncthreads();
  double s = 0.0;
  if (col < ncol) {
    for (int a = 0; a < r0; ++a) {
      T p = 0;
      for (int i = 0; i < r0; ++i) p += M0s[a * r0 + i] * Zs[i * 128 + tid];
      const int64_t idx = (int64_t)a * ncol + col;
      const T g = -MtX[idx] + p + sparse;                 // gradient, ntd.py:608
      const T c = core[idx];
      const T st = step * g;
      const T d = st < c ? st : c;                         // np.minimum(gradient_step * gradient, core)
      core[idx] = c - d;
      s += (double)d * (double)d;
    }
  }
  s = block_sum(s, sh);
  if (tid == 0) part[blockIdx.x] = s;
}

#define DISPATCH_T(dtype, CALL_F, CALL_D) \
  if ((dtype) == NNFAC_F32) { CALL_F; } else if ((dtype) == NNFAC_F64) { CALL_D; } else { \
    nnfac_set_error("bad dtype %d", (int)(dtype)); return NNFAC_ERR_ARG; }

}  // namespace

extern "C" {

int nnfac_mu_terms(nnfac_ctx* ctx, int dtype, double beta, const void* K, const void* X, void* P,
                   void* Q, int64_t count, void* stream) {
  NNFAC_ARG(ctx && K && count > 0, "nnfac_mu_terms: bad argument");
  NNFAC_ARG(beta >= 0, "nnfac_mu_terms: negative beta");
  const int mode = beta == 1.0 ? 0 : beta == 2.0 ? 1 : beta == 3.0 ? 2 : 3;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(count, ctx->sm_count);
  DISPATCH_T(dtype,
             (mu_terms_kernel<float><<<grid, 256, 0, st>>>(mode, (float)beta, (const float*)K, (const float*)X, (float*)P, (float*)Q, count)),
             (mu_terms_kernel<double><<<grid, 256, 0, st>>>(mode, beta, (const double*)K, (const double*)X, (double*)P, (double*)Q, count)));
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

int nnfac_mu_apply(nnfac_ctx* ctx, int dtype, void* F_out, const void* F_in, const void* num,
                   const void* den_mat, const void* den_vec, int vec_per_row, int64_t rows,
                   int64_t cols, double gamma, double floor_value, void* stream) {
  NNFAC_ARG(ctx && F_out && F_in && num && (den_mat || den_vec) && rows > 0 && cols > 0, "nnfac_mu_apply: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(rows * cols, ctx->sm_count);
  DISPATCH_T(dtype,
             (mu_apply_kernel<float><<<grid, 256, 0, st>>>((float*)F_out, (const float*)F_in, (const float*)num, (const float*)den_mat, (const float*)den_vec, vec_per_row, rows, cols, (float)gamma, (float)floor_value)),
             (mu_apply_kernel<double><<<grid, 256, 0, st>>>((double*)F_out, (const double*)F_in, (const double*)num, (const double*)den_mat, (const double*)den_vec, vec_per_row, rows, cols, gamma, floor_value)));
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

int nnfac_beta_divergence(nnfac_ctx* ctx, int dtype, double beta, const void* A, const void* B,
                          int64_t count, double* out, void* stream) {
  NNFAC_ARG(ctx && A && B && out && count > 0, "nnfac_beta_divergence: bad argument");
  NNFAC_ARG(beta >= 0, "nnfac_beta_divergence: negative beta");
  const int op = beta == 1.0 ? 0 : beta == 0.0 ? 1 : 2;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, return run_reduce<float>(ctx, op, beta, (const float*)A, (const float*)B, count, out, st),
             return run_reduce<double>(ctx, op, beta, (const double*)A, (const double*)B, count, out, st));
}

int nnfac_sq_diff(nnfac_ctx* ctx, int dtype, const void* A, const void* B, int64_t count,
                  double* out, void* stream) {
  NNFAC_ARG(ctx && A && out && count > 0, "nnfac_sq_diff: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int op = B ? 3 : 5;
  DISPATCH_T(dtype, return run_reduce<float>(ctx, op, 0.0, (const float*)A, (const float*)B, count, out, st),
             return run_reduce<double>(ctx, op, 0.0, (const double*)A, (const double*)B, count, out, st));
}

int nnfac_dot(nnfac_ctx* ctx, int dtype, const void* A, const void* B, int64_t count, double* out,
              void* stream) {
  NNFAC_ARG(ctx && A && B && out && count > 0, "nnfac_dot: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, return run_reduce<float>(ctx, 4, 0.0, (const float*)A, (const float*)B, count, out, st),
             return run_reduce<double>(ctx, 4, 0.0, (const double*)A, (const double*)B, count, out, st));
}

int nnfac_row_sums(nnfac_ctx* ctx, int dtype, const void* A, int64_t lda, int64_t rows,
                   int64_t cols, void* out, void* stream) {
  NNFAC_ARG(ctx && A && out && rows > 0 && cols > 0, "nnfac_row_sums: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (cols >= 2048 && rows <= 65535) {
    DISPATCH_T(dtype, (row_sums_wide_kernel<float><<<(unsigned)rows, 256, 0, st>>>((const float*)A, lda, cols, (float*)out)),
               (row_sums_wide_kernel<double><<<(unsigned)rows, 256, 0, st>>>((const double*)A, lda, cols, (double*)out)));
  } else {
    const int grid = (int)ceil_div64(rows, 8);
    DISPATCH_T(dtype, (row_sums_kernel<float><<<grid, 256, 0, st>>>((const float*)A, lda, rows, cols, (float*)out)),
               (row_sums_kernel<double><<<grid, 256, 0, st>>>((const double*)A, lda, rows, cols, (double*)out)));
  }
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

int nnfac_norm1(nnfac_ctx* ctx, int dtype, const void* A, int64_t lda, int64_t rows, int64_t cols,
                double* out, void* stream) {
  NNFAC_ARG(ctx && A && out && rows > 0 && cols > 0, "nnfac_norm1: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)(ceil_div64(cols, RB) < RMAX_BLOCKS ? ceil_div64(cols, RB) : RMAX_BLOCKS);
  const int grc = nnfac_guard_enter(ctx, NNFAC_GUARD_RED, st);
  if (grc) return grc;
  DISPATCH_T(dtype, (norm1_stage1<float><<<blocks, RB, 0, st>>>((const float*)A, lda, rows, cols, ctx->red)),
             (norm1_stage1<double><<<blocks, RB, 0, st>>>((const double*)A, lda, rows, cols, ctx->red)));
  NNFAC_LAUNCH_CHECK(ctx);
  reduce_stage2<<<1, RB, 0, st>>>(ctx->red, blocks, out, 1);
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

int nnfac_transpose(nnfac_ctx* ctx, int dtype, void* out, int64_t ld_out, const void* in,
                    int64_t ld_in, int64_t rows, int64_t cols, void* stream) {
  NNFAC_ARG(ctx && out && in && rows > 0 && cols > 0, "nnfac_transpose: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  NNFAC_ARG(ceil_div64(rows, 32) <= 65535, "nnfac_transpose: too many rows");
  dim3 grid((unsigned)ceil_div64(cols, 32), (unsigned)ceil_div64(rows, 32)), block(32, 8);
  DISPATCH_T(dtype, (transpose_kernel<float><<<grid, block, 0, st>>>((float*)out, ld_out, (const float*)in, ld_in, rows, cols)),
             (transpose_kernel<double><<<grid, block, 0, st>>>((double*)out, ld_out, (const double*)in, ld_in, rows, cols)));
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

int nnfac_khatri_rao(nnfac_ctx* ctx, int dtype, void* out, const void* A, int64_t I, const void* B,
                     int64_t J, int64_t r, void* stream) {
  NNFAC_ARG(ctx && out && A && B && I > 0 && J > 0 && r > 0, "nnfac_khatri_rao: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(I * J * r, ctx->sm_count);
  DISPATCH_T(dtype, (khatri_rao_kernel<float><<<grid, 256, 0, st>>>((float*)out, (const float*)A, I, (const float*)B, J, r)),
             (khatri_rao_kernel<double><<<grid, 256, 0, st>>>((double*)out, (const double*)A, I, (const double*)B, J, r)));
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

int nnfac_hadamard(nnfac_ctx* ctx, int dtype, void* out, const void* A, const void* B,
                   int64_t count, void* stream) {
  NNFAC_ARG(ctx && out && A && B && count > 0, "nnfac_hadamard: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(count, ctx->sm_count);
  DISPATCH_T(dtype, (hadamard_kernel<float><<<grid, 256, 0, st>>>((float*)out, (const float*)A, (const float*)B, count)),
             (hadamard_kernel<double><<<grid, 256, 0, st>>>((double*)out, (const double*)A, (const double*)B, count)));
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

int nnfac_axpby(nnfac_ctx* ctx, int dtype, void* out, double a, const void* X, double b, const void* Y, int64_t count, void* stream) {
  NNFAC_ARG(ctx && out && X && Y && count > 0, "nnfac_axpby: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(count, ctx->sm_count);
  DISPATCH_T(dtype, (axpby_kernel<float><<<grid, 256, 0, st>>>((float*)out, (float)a, (const float*)X, (float)b, (const float*)Y, count)),
             (axpby_kernel<double><<<grid, 256, 0, st>>>((double*)out, a, (const double*)X, b, (const double*)Y, count)));
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

static int core_pg_step_impl(nnfac_ctx* ctx, int dtype, void* core, const void* MtX, const void* P, int64_t count, double step,
                             double sparse, int dev_scalars, double delta, double* state, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)(ceil_div64(count, RB) < RMAX_BLOCKS ? ceil_div64(count, RB) : RMAX_BLOCKS);
  // the partials live behind the mailbox area of the sweep so that they never collide with other reductions in flight
  double* part = ctx->red + 32768;
  const int grc = nnfac_guard_enter(ctx, NNFAC_GUARD_RED, st);
  if (grc) return grc;
  DISPATCH_T(dtype, (core_pg_step_kernel<float><<<blocks, RB, 0, st>>>((float*)core, (const float*)MtX, (const float*)P, count,
                                                                        (float)step, (float)sparse, dev_scalars, state, part)),
             (core_pg_step_kernel<double><<<blocks, RB, 0, st>>>((double*)core, (const double*)MtX, (const double*)P, count, step,
                                                                  sparse, dev_scalars, state, part)));
  NNFAC_LAUNCH_CHECK(ctx);
  core_pg_finish_kernel<<<1, 32, 0, st>>>(part, blocks, delta, state);
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

int nnfac_core_pg_step(nnfac_ctx* ctx, int dtype, void* core, const void* MtX, const void* P, int64_t count, double step,
                       double sparse, double delta, double* state, void* stream) {
  NNFAC_ARG(ctx && core && MtX && P && state && count > 0, "nnfac_core_pg_step: bad argument");
  return core_pg_step_impl(ctx, dtype, core, MtX, P, count, step, sparse, 0, delta, state, stream);
}

}  // extern "C"

template <typename T>
static int core_pg_step3_launch(nnfac_ctx* ctx, T* core, const T* MtX, const T* M0, const T* M1, const T* M2, int r0, int r1, int r2,
                                T* Z, double step, double sparse, int dev_scalars, double delta, double* state, cudaStream_t st) {
  const size_t smem_a = sizeof(T) * ((size_t)2 * r1 * r2 + (size_t)r1 * r1 + (size_t)r2 * (r2 + 1));
  const size_t smem_b = sizeof(T) * ((size_t)r0 * r0 + (size_t)r0 * 128);
  NNFAC_CUDA(cudaFuncSetAttribute(core_modes12_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
  NNFAC_CUDA(cudaFuncSetAttribute(core_mode0_step_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  const int ncol = r1 * r2, blocks = (ncol + 127) / 128;
  double* part = ctx->red + 32768;
  const int grc = nnfac_guard_enter(ctx, NNFAC_GUARD_RED, st);
  if (grc) return grc;
  core_modes12_kernel<T><<<r0, 256, smem_a, st>>>(core, M1, M2, Z, r1, r2, state);
  NNFAC_LAUNCH_CHECK(ctx);
  core_mode0_step_kernel<T><<<blocks, 128, smem_b, st>>>(core, MtX, Z, M0, r0, ncol, (T)step, (T)sparse, dev_scalars, state, part);
  NNFAC_LAUNCH_CHECK(ctx);
  core_pg_finish_kernel<<<1, 32, 0, st>>>(part, blocks, delta, state);
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

extern "C" {

int nnfac_core_pg_step3(nnfac_ctx* ctx, int dtype, void* core, const void* MtX, const void* M0, const void* M1, const void* M2,
                        int r0, int r1, int r2, void* Z, double step, double sparse, int dev_scalars, double delta, double* state,
                        void* stream) {
  NNFAC_ARG(ctx && core && MtX && M0 && M1 && M2 && Z && state, "nnfac_core_pg_step3: bad argument");
  NNFAC_ARG(r0 >= 1 && r1 >= 1 && r2 >= 1 && r0 <= 64 && r1 <= 64 && r2 <= 64, "nnfac_core_pg_step3: ranks must be in 1..64");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == NNFAC_F32)
    return core_pg_step3_launch<float>(ctx, (float*)core, (const float*)MtX, (const float*)M0, (const float*)M1, (const float*)M2, r0,
                                       r1, r2, (float*)Z, step, sparse, dev_scalars, delta, state, st);
  if (dtype == NNFAC_F64)
    return core_pg_step3_launch<double>(ctx, (double*)core, (const double*)MtX, (const double*)M0, (const double*)M1,
                                        (const double*)M2, r0, r1, r2, (double*)Z, step, sparse, dev_scalars, delta, state, st);
  nnfac_set_error("bad dtype %d", dtype);
  return NNFAC_ERR_ARG;
}

int nnfac_core_pg_step_dev(nnfac_ctx* ctx, int dtype, void* core, const void* MtX, const void* P, int64_t count, double delta,
                           double* state, void* stream) {
  NNFAC_ARG(ctx && core && MtX && P && state && count > 0, "nnfac_core_pg_step_dev: bad argument");
  return core_pg_step_impl(ctx, dtype, core, MtX, P, count, 0.0, 0.0, 1, delta, state, stream);
}

int nnfac_normalize_rows(nnfac_ctx* ctx, int dtype, void* A, int64_t lda, int64_t rows,
                         int64_t cols, void* stream) {
  NNFAC_ARG(ctx && A && rows > 0 && cols > 0, "nnfac_normalize_rows: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, (normalize_rows_kernel<float><<<(unsigned)rows, RB, 0, st>>>((float*)A, lda, cols)),
             (normalize_rows_kernel<double><<<(unsigned)rows, RB, 0, st>>>((double*)A, lda, cols)));
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

}  // extern "C"
