// General HALS NNLS sweep (nn_fac/update_rules/nnls.py:156-198, deterministic rule): ANY rank, ANY number of columns, every
// option, fp32 or fp64.  This is the fallback behind the specialised kernels (tensor-core sweep: fp32, rank <= 128, no
// row-wise options; register-resident CUDA-core sweep: rank <= 128, row-wise options only while all columns are resident):
// nothing is kept on chip across rows, so nothing limits the shape -- V round-trips through L1/L2 for every row
// (r^2 loads per column and sweep) and every sweep is a handful of launches whose stop decision is taken on the device.
// It is correct, deterministic and slow; it exists so that no call the reference accepts is refused.
//
// Without row-wise options one sweep is ONE kernel (one thread per column walks the rows in order) plus a one-block kernel
// that adds the per-block partial sums in a fixed order and evaluates nnls.py:156.  With `normalize` / `nonzero` a row
// needs sums over ALL columns before the next row may start (nnls.py:173-185): three launches per row (update + partial
// sums; one block: totals, zero-row and zero-diagonal rules; fix-up of the row).  Every kernel returns at once when the
// device-side state says the solve has ended, so the host enqueues maxiter sweeps without ever synchronising.
#include "common.cuh"

namespace {

// device-side state of one solve (doubles): eps0, eps, cnt, done, zero_row, nd (running sum of the sweep), row_ss, row_mx,
// row_action (0 none, 1 fill with row_fill), row_fill
enum { S_EPS0 = 0, S_EPS, S_CNT, S_DONE, S_ZROW, S_ND, S_SS, S_MX, S_ACT, S_FILL, S_COUNT };

template <typename T>
struct GenArgs {
  const T* b;
  const T* G;
  T* V;
  int64_t ld_b, ld_g, ld_v, n;
  int r, maxiter;
  double delta;
  T sp;
  unsigned flags;
  double* st;      // [S_COUNT]
  double* part;    // [3 * grid] per-block partials (nd, ss, mx)
  double* result;
};

__global__ void gen_init_kernel(double* st) {
  if (threadIdx.x == 0) {
    st[S_EPS0] = 0.0; st[S_EPS] = 1.0; st[S_CNT] = 1.0; st[S_DONE] = 0.0; st[S_ZROW] = -1.0; st[S_ND] = 0.0;
    st[S_SS] = 0.0; st[S_MX] = 0.0; st[S_ACT] = 0.0; st[S_FILL] = 0.0;
  }
}

// dot(G[k, :], V[:, c]) with V as it is NOW (rows < k already updated: Gauss-Seidel); four independent partial sums
template <typename T>
__device__ __forceinline__ T row_dot(const GenArgs<T>& a, int k, int64_t c) {
  const T* g = a.G + (int64_t)k * a.ld_g;
  const T* v = a.V + c;
  T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int j = 0;
  for (; j + 4 <= a.r; j += 4) {
    s0 = fma(__ldg(g + j), v[(int64_t)j * a.ld_v], s0);
    s1 = fma(__ldg(g + j + 1), v[(int64_t)(j + 1) * a.ld_v], s1);
    s2 = fma(__ldg(g + j + 2), v[(int64_t)(j + 2) * a.ld_v], s2);
    s3 = fma(__ldg(g + j + 3), v[(int64_t)(j + 3) * a.ld_v], s3);
  }
  for (; j < a.r; ++j) s0 = fma(__ldg(g + j), v[(int64_t)j * a.ld_v], s0);
  return (s0 + s1) + (s2 + s3);
}

// one whole sweep, no row-wise options: thread = column (its column of V is private, so the rows can be walked in order
// without any synchronisation)
template <typename T>
__global__ void __launch_bounds__(256) gen_sweep_kernel(GenArgs<T> a) {
  __shared__ double sh[40];
  if (a.st[S_DONE] != 0.0) return;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double nd = 0.0;
  if (c < a.n) {
    for (int k = 0; k < a.r; ++k) {
      const T dk = a.G[(int64_t)k * a.ld_g + k];
      if (dk != T(0)) {                                          // nnls.py:160
        const T cur = a.V[(int64_t)k * a.ld_v + c];
        T d = (a.b[(int64_t)k * a.ld_b + c] - row_dot(a, k, c) - a.sp) / dk;   // nnls.py:163 / :167
        d = d > -cur ? d : -cur;
        a.V[(int64_t)k * a.ld_v + c] = cur + d;
        nd += (double)d * (double)d;                             // nnls.py:170
      }
    }
  }
  nd = block_sum(nd, sh);
  if (threadIdx.x == 0) a.part[blockIdx.x] = nd;
}

// row k of a sweep with row-wise options: update + per-block partials of (squared step, squared norm, max |.|) of the new row
template <typename T>
__global__ void __launch_bounds__(256) gen_row_update_kernel(GenArgs<T> a, int k, int nblocks) {
  __shared__ double sh[40];
  if (a.st[S_DONE] != 0.0) return;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const T dk = a.G[(int64_t)k * a.ld_g + k];
  double nd = 0.0, ss = 0.0, mx = 0.0;
  if (c < a.n) {
    T x = a.V[(int64_t)k * a.ld_v + c];
    if (dk != T(0)) {
      T d = (a.b[(int64_t)k * a.ld_b + c] - row_dot(a, k, c) - a.sp) / dk;
      d = d > -x ? d : -x;
      x += d;
      a.V[(int64_t)k * a.ld_v + c] = x;
      nd = (double)d * (double)d;
    }
    ss = (double)x * (double)x;
    mx = fabs((double)x);
  }
  nd = block_sum(nd, sh);
  ss = block_sum(ss, sh);
  mx = block_max(mx, sh);
  if (threadIdx.x == 0) {
    a.part[blockIdx.x] = nd;
    a.part[nblocks + blockIdx.x] = ss;
    a.part[2 * nblocks + blockIdx.x] = mx;
  }
}

// one block: totals of the row in block order; nnls.py:173-177 (all-zero row -> 1e-16 max(V); zero diagonal -> error)
template <typename T>
__global__ void __launch_bounds__(256) gen_row_stats_kernel(GenArgs<T> a, int k, int nblocks) {
  __shared__ double sh[40];
  if (a.st[S_DONE] != 0.0) return;
  const bool nonzero = (a.flags & NNFAC_HALS_NONZERO) != 0;
  const T dk = a.G[(int64_t)k * a.ld_g + k];
  __shared__ double tot[3];
  if (threadIdx.x == 0) {
    double nd = 0.0, ss = 0.0, mx = 0.0;
    for (int i = 0; i < nblocks; ++i) { nd += a.part[i]; ss += a.part[nblocks + i]; mx = fmax(mx, a.part[2 * nblocks + i]); }
    tot[0] = nd; tot[1] = ss; tot[2] = mx;
  }
  __syncthreads();
  const bool fill = nonzero && dk != T(0) && tot[2] == 0.0;
  double vm = -1.0e300;
  if (fill) {                                                     // np.max(V) over the whole matrix (rare)
    for (int64_t i = threadIdx.x; i < (int64_t)a.r * a.n; i += blockDim.x)
      vm = fmax(vm, (double)a.V[(i / a.n) * a.ld_v + (i % a.n)]);
    vm = block_max(vm, sh);
  }
  if (threadIdx.x == 0) {
    a.st[S_ND] += tot[0];
    a.st[S_SS] = tot[1]; a.st[S_MX] = tot[2]; a.st[S_ACT] = 0.0;
    if (dk == T(0) && nonzero) {                                  // nnls.py:176-177
      a.st[S_ZROW] = (double)k;
      a.st[S_DONE] = 1.0;
    } else if (fill) {
      const T f = (T)(1e-16 * vm);
      a.st[S_ACT] = 1.0; a.st[S_FILL] = (double)f;
      a.st[S_SS] = (double)a.n * (double)f * (double)f;
    }
  }
}

// nnls.py:174 (fill) and :179-185 (normalise the row, or the constant row 1/sqrt(n) when its norm is zero)
template <typename T>
__global__ void __launch_bounds__(256) gen_row_fix_kernel(GenArgs<T> a, int k) {
  if (a.st[S_DONE] != 0.0) return;
  const bool normalize = (a.flags & NNFAC_HALS_NORMALIZE) != 0;
  const bool fill = a.st[S_ACT] != 0.0;
  if (!fill && !normalize) return;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.n) return;
  T x = fill ? (T)a.st[S_FILL] : a.V[(int64_t)k * a.ld_v + c];
  if (normalize) {
    const double nrm = sqrt(a.st[S_SS]);
    x = nrm != 0.0 ? (T)((double)x / nrm) : (T)(1.0 / sqrt((double)a.n));
  }
  a.V[(int64_t)k * a.ld_v + c] = x;
}

// end of a sweep (one block): nnls.py:187-196 and the loop test of :156 for the next one
template <typename T>
__global__ void gen_sweep_end_kernel(GenArgs<T> a, int nblocks, int rowops) {
  if (threadIdx.x != 0 || a.st[S_DONE] != 0.0) return;
  double tot = 0.0;
  if (rowops) tot = a.st[S_ND];
  else for (int i = 0; i < nblocks; ++i) tot += a.part[i];
  a.st[S_ND] = 0.0;
  double cnt = a.st[S_CNT];
  if (cnt == 1.0) a.st[S_EPS0] = tot;
  a.st[S_EPS] = tot;
  cnt += 1.0;
  const bool normalize = (a.flags & NNFAC_HALS_NORMALIZE) != 0;
  bool done = !(tot >= a.delta * a.st[S_EPS0] && cnt <= (double)a.maxiter);
  if (tot == 0.0 && !normalize) {
    // remaining sweeps are no-ops.  nnls.py:156 keeps looping on `0 >= delta * 0` only when the first sweep already moved
    // nothing (eps0 == 0: it then burns all maxiter sweeps, cnt = maxiter + 1); otherwise the test fails and cnt stays
    if (a.st[S_EPS0] == 0.0 && cnt < (double)a.maxiter + 1.0) cnt = (double)a.maxiter + 1.0;
    done = true;
  }
  a.st[S_CNT] = cnt;
  if (done) a.st[S_DONE] = 1.0;
}

__global__ void gen_result_kernel(const double* st, double* result) {
  if (threadIdx.x == 0) {
    result[0] = st[S_EPS];
    result[1] = st[S_CNT];
    result[2] = st[S_ZROW];
    result[3] = st[S_CNT] - 1.0;
  }
}

template <typename T>
int run_general(nnfac_ctx* ctx, const void* UtM, int64_t ld_utm, const void* UtU, int64_t ld_utu, void* V, int64_t ld_v, int r,
                int64_t n, int maxiter, double delta, double sparsity, unsigned flags, double* result, cudaStream_t st) {
  const int64_t nblocks = ceil_div64(n, 256);
  if ((size_t)(3 * nblocks + S_COUNT) > ctx->red_count) {
    nnfac_set_error("nnfac_hals_nnls: %lld columns exceed the reduction scratch of the general sweep", (long long)n);
    return NNFAC_ERR_UNSUPPORTED;
  }
  const int grc = nnfac_guard_enter(ctx, NNFAC_GUARD_RED, st);
  if (grc) return grc;
  GenArgs<T> a;
  a.b = (const T*)UtM; a.G = (const T*)UtU; a.V = (T*)V; a.ld_b = ld_utm; a.ld_g = ld_utu; a.ld_v = ld_v; a.n = n;
  a.r = r; a.maxiter = maxiter; a.delta = delta; a.sp = (T)sparsity; a.flags = flags;
  a.st = ctx->red; a.part = ctx->red + S_COUNT; a.result = result;
  const bool rowops = (flags & (NNFAC_HALS_NORMALIZE | NNFAC_HALS_NONZERO)) != 0;
  gen_init_kernel<<<1, 32, 0, st>>>(a.st);
  NNFAC_LAUNCH_CHECK(ctx);
  // The kernels of a finished solve return at once, but enqueueing maxiter * (3 r + 1) of them is not free: every few sweeps the
  // host looks at the device-side flag (one 8-byte copy + a stream synchronisation; skipped while the stream is being captured)
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  const int check_every = rowops ? 4 : 16;
  for (int sweep = 0; sweep < maxiter; ++sweep) {
    if (cap == cudaStreamCaptureStatusNone && sweep > 0 && sweep % check_every == 0) {
      double done = 0.0;
      NNFAC_CUDA(cudaMemcpyAsync(&done, a.st + S_DONE, sizeof(double), cudaMemcpyDeviceToHost, st));
      NNFAC_CUDA(cudaStreamSynchronize(st));
      if (done != 0.0) break;
    }
    if (!rowops) {
      gen_sweep_kernel<T><<<(unsigned)nblocks, 256, 0, st>>>(a);
    } else {
      for (int k = 0; k < r; ++k) {
        gen_row_update_kernel<T><<<(unsigned)nblocks, 256, 0, st>>>(a, k, (int)nblocks);
        gen_row_stats_kernel<T><<<1, 256, 0, st>>>(a, k, (int)nblocks);
        gen_row_fix_kernel<T><<<(unsigned)nblocks, 256, 0, st>>>(a, k);
      }
    }
    gen_sweep_end_kernel<T><<<1, 32, 0, st>>>(a, (int)nblocks, rowops ? 1 : 0);
    NNFAC_LAUNCH_CHECK(ctx);
  }
  gen_result_kernel<<<1, 32, 0, st>>>(a.st, result);
  NNFAC_LAUNCH_CHECK(ctx);
  return NNFAC_OK;
}

}  // namespace

int nnfac_sweep_general(nnfac_ctx* ctx, int dtype, const void* UtM, int64_t ld_utm, const void* UtU, int64_t ld_utu, void* V,
                        int64_t ld_v, int r, int64_t n, int maxiter, double delta, double sparsity, unsigned flags, double* result,
                        cudaStream_t st) {
  if (dtype == NNFAC_F32) return run_general<float>(ctx, UtM, ld_utm, UtU, ld_utu, V, ld_v, r, n, maxiter, delta, sparsity, flags, result, st);
  if (dtype == NNFAC_F64) return run_general<double>(ctx, UtM, ld_utm, UtU, ld_utu, V, ld_v, r, n, maxiter, delta, sparsity, flags, result, st);
  nnfac_set_error("nnfac_hals_nnls: bad dtype %d", dtype);
  return NNFAC_ERR_ARG;
}
