// Accelerated-HALS NNLS sweep (nn_fac/update_rules/nnls.py:156-198, deterministic rule).
//
// One persistent cooperative kernel runs ALL sweeps of one hals_nnls_acc call:
//   * the r x r Gram (UtU) sits in shared memory, zero padded to RP x RP;
//   * each group of L adjacent lanes owns C consecutive columns of V; a lane keeps RP/L rows of
//     those columns in registers for the whole call (V touches HBM twice: load, store);
//   * row k of the Gauss-Seidel sweep = RP/L * C FMAs per lane against a broadcast-loaded slice
//     of Gram row k, an xor-shuffle reduction over the L lanes, and the clamped update by the
//     lane that owns row k;
//   * the stop test needs sum ||dV||^2 over every column: per-CTA partials (warp shuffles + one
//     smem hop), a grid barrier, and a fixed-order re-summation so that every CTA (and every
//     replica on other GPUs) takes the same decision.
// When n exceeds what the grid can keep in registers the kernel walks column batches and V
// round-trips through global memory once per sweep.  Rank > 128, and row-wise options (normalize / nonzero) on more columns
// than the grid keeps resident, go to the general sweep of csrc/hals_general.cu.
#pragma once
#include "common.cuh"

namespace hals {

template <typename T> struct VecT;
template <> struct VecT<float> { using type = float4; static constexpr int N = 4; };
template <> struct VecT<double> { using type = double2; static constexpr int N = 2; };

template <typename T>
struct SweepArgs {
  const T* b;   // UtM, r x n
  const T* G;   // UtU, r x r
  T* V;         // r x n, in place
  int64_t ld_b, ld_g, ld_v, n;
  int r, maxiter;
  double delta;
  T sp;
  unsigned flags;
  int64_t nbatch;
  int slab_ld;        // > 0: the CTA keeps its r x slab_ld slice of UtM in shared memory
  double* part;       // [2][2 * gridDim.x] barrier partials
  unsigned* counter;  // zeroed before launch
  double* result;     // {eps, cnt, zero_diag_row, sweeps}
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide reduction of (sum, max) with a fixed summation order; needs a cooperative launch.
__device__ __forceinline__ void grid_reduce(double& vsum, double& vmax, double* part, unsigned* counter,
                                            unsigned& epoch, double* sh) {
  const unsigned nb = gridDim.x;
  if (nb == 1) return;
  double* slot = part + (size_t)(epoch & 1u) * 2u * nb;
  if (threadIdx.x == 0) {
    slot[2 * blockIdx.x] = vsum;
    slot[2 * blockIdx.x + 1] = vmax;
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned target = (epoch + 1u) * nb;
    while (ld_acquire_u32(counter) < target) {}
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = 0.0, m = -1.0e300;
    for (unsigned i = threadIdx.x; i < nb; i += 32) {
      s += __ldcg(slot + 2 * i);
      m = fmax(m, __ldcg(slot + 2 * i + 1));
    }
    s = warp_sum(s);
    m = warp_max(m);
    if (threadIdx.x == 0) { sh[34] = s; sh[35] = m; }
  }
  __syncthreads();
  vsum = sh[34];
  vmax = sh[35];
  ++epoch;
}

template <typename T, int RP, int L, int C, bool ROWOPS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) hals_sweep_kernel(SweepArgs<T> a) {
  constexpr int VEC = VecT<T>::N;
  constexpr int NR = RP / L;
  constexpr int NCH = NR / VEC;
  static_assert(NCH >= 1 && NCH * VEC * L == RP, "bad sweep tiling");
  using V4 = typename VecT<T>::type;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Gs = reinterpret_cast<T*>(smem_raw);
  T* diag = Gs + RP * RP;
  double* sh = reinterpret_cast<double*>(diag + RP);  // 40 doubles
  T* bs = reinterpret_cast<T*>(sh + 40);              // optional r x slab_ld slice of UtM

  const int t = threadIdx.x;
  const int lane_l = t % L;
  const int gid = t / L;
  const int GP = blockDim.x / L;
  const int r = a.r;
  const bool normalize = ROWOPS && (a.flags & NNFAC_HALS_NORMALIZE);
  const bool nonzero = ROWOPS && (a.flags & NNFAC_HALS_NONZERO);

  for (int idx = t; idx < RP * RP; idx += blockDim.x) {
    const int i = idx / RP, j = idx % RP;
    Gs[idx] = (i < r && j < r) ? a.G[(int64_t)i * a.ld_g + j] : T(0);
  }
  for (int i = t; i < RP; i += blockDim.x) diag[i] = i < r ? a.G[(int64_t)i * a.ld_g + i] : T(0);
  __syncthreads();

  T v[NCH][VEC][C];
  const bool single = (a.nbatch == 1);
  const bool vec_b = (C * sizeof(T) == 16) && (a.ld_b % C == 0) && ((reinterpret_cast<uintptr_t>(a.b) & 15) == 0);

  auto col_of = [&](int64_t batch) -> int64_t {
    return ((batch * gridDim.x + blockIdx.x) * (int64_t)GP + gid) * C;
  };
  auto load_v = [&](int64_t col0) {
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int row = VEC * (lane_l + L * i) + e;
#pragma unroll
        for (int c = 0; c < C; ++c)
          v[i][e][c] = (row < r && col0 + c < a.n) ? a.V[(int64_t)row * a.ld_v + col0 + c] : T(0);
      }
  };
  auto store_v = [&](int64_t col0) {
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int row = VEC * (lane_l + L * i) + e;
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (row < r && col0 + c < a.n) a.V[(int64_t)row * a.ld_v + col0 + c] = v[i][e][c];
      }
  };

  if (a.slab_ld > 0) {
    // stage this CTA's columns of UtM once: the sweep re-reads them every sweep, and a global
    // load in the middle of the row loop is a ~700-cycle bubble that 7 warps cannot hide
    const int64_t cta_col0 = (int64_t)blockIdx.x * GP * C;
    for (int idx = t; idx < r * a.slab_ld; idx += blockDim.x) {
      const int k = idx / a.slab_ld, c = idx % a.slab_ld;
      bs[idx] = (cta_col0 + c < a.n) ? a.b[(int64_t)k * a.ld_b + cta_col0 + c] : T(0);
    }
    __syncthreads();
  }

  unsigned epoch = 0;
  double eps0 = 0.0, eps = 1.0;
  int cnt = 1;
  int zero_row = -1;
  if (single) load_v(col_of(0));

  while (eps >= a.delta * eps0 && cnt <= a.maxiter) {
    T nd = T(0);
    for (int64_t batch = 0; batch < a.nbatch; ++batch) {
      const int64_t col0 = col_of(batch);
      if (!single) load_v(col0);
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
#pragma unroll 1
        for (int lo = 0; lo < L; ++lo) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const int k = VEC * (lo + L * i) + e;
            if (k >= r) continue;
            const T dk = diag[k];
            const bool owner = (lane_l == lo);
            if (dk != T(0)) {
              T bk[C];
              if (owner && a.slab_ld > 0) {
                const T* bp = bs + k * a.slab_ld + gid * C;
                if (C * sizeof(T) == 16) {
                  const V4 q = *reinterpret_cast<const V4*>(bp);
                  const T* qq = reinterpret_cast<const T*>(&q);
#pragma unroll
                  for (int c = 0; c < C; ++c) bk[c] = qq[c];
                } else {
#pragma unroll
                  for (int c = 0; c < C; ++c) bk[c] = bp[c];
                }
              } else if (owner) {
                const T* bp = a.b + (int64_t)k * a.ld_b + col0;
                if (vec_b && col0 + C <= a.n) {
                  const V4 q = *reinterpret_cast<const V4*>(bp);
                  const T* qq = reinterpret_cast<const T*>(&q);
#pragma unroll
                  for (int c = 0; c < C; ++c) bk[c] = qq[c];
                } else {
#pragma unroll
                  for (int c = 0; c < C; ++c) bk[c] = (col0 + c < a.n) ? bp[c] : T(0);
                }
              }
              T acc[C];
#pragma unroll
              for (int c = 0; c < C; ++c) acc[c] = T(0);
              const T* grow = Gs + k * RP + VEC * lane_l;
#pragma unroll
              for (int i2 = 0; i2 < NCH; ++i2) {
                const V4 g4 = *reinterpret_cast<const V4*>(grow + VEC * L * i2);
                const T* gg = reinterpret_cast<const T*>(&g4);
#pragma unroll
                for (int e2 = 0; e2 < VEC; ++e2)
#pragma unroll
                  for (int c = 0; c < C; ++c) acc[c] = fma(gg[e2], v[i2][e2][c], acc[c]);
              }
#pragma unroll
              for (int off = L / 2; off > 0; off >>= 1)
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
              if (owner) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                  const T cur = v[i][e][c];
                  T d = (bk[c] - acc[c] - a.sp) / dk;         // nnls.py:163 / :167
                  d = d > -cur ? d : -cur;                     // np.maximum(., -V[k,:])
                  if (col0 + c >= a.n) d = T(0);
                  v[i][e][c] = cur + d;
                  nd = fma(d, d, nd);                          // nnls.py:170
                }
              }
            } else if (nonzero) {
              zero_row = k;                                    // nnls.py:176-177
            }
            if (ROWOPS) {
              if (zero_row >= 0) break;
              double ss = 0.0, mx = 0.0;
              if (owner) {
#pragma unroll
                for (int c = 0; c < C; ++c)
                  if (col0 + c < a.n) {
                    const double x = (double)v[i][e][c];
                    ss += x * x;
                    mx = fmax(mx, fabs(x));
                  }
              }
              ss = block_sum(ss, sh);
              mx = block_max(mx, sh);
              grid_reduce(ss, mx, a.part, a.counter, epoch, sh);
              if (nonzero && dk != T(0) && mx == 0.0) {         // nnls.py:173-174
                double vm = -1.0e300, dummy = 0.0;
#pragma unroll
                for (int i3 = 0; i3 < NCH; ++i3)
#pragma unroll
                  for (int e3 = 0; e3 < VEC; ++e3)
#pragma unroll
                    for (int c = 0; c < C; ++c)
                      if (VEC * (lane_l + L * i3) + e3 < r && col0 + c < a.n) vm = fmax(vm, (double)v[i3][e3][c]);
                vm = block_max(vm, sh);
                grid_reduce(dummy, vm, a.part, a.counter, epoch, sh);
                const T fill = (T)(1e-16 * vm);
                if (owner)
#pragma unroll
                  for (int c = 0; c < C; ++c) v[i][e][c] = fill;
                ss = (double)a.n * (double)fill * (double)fill;
              }
              if (normalize) {                                  // nnls.py:179-185
                const double nrm = sqrt(ss);
                if (owner) {
#pragma unroll
                  for (int c = 0; c < C; ++c)
                    v[i][e][c] = nrm != 0.0 ? (T)((double)v[i][e][c] / nrm) : (T)(1.0 / sqrt((double)a.n));
                }
              }
            }
          }
          if (ROWOPS && zero_row >= 0) break;
        }
        if (ROWOPS && zero_row >= 0) break;
      }
      if (!single) store_v(col0);
    }
    if (ROWOPS && zero_row >= 0) break;
    double tot = block_sum((double)nd, sh), dummy = 0.0;
    grid_reduce(tot, dummy, a.part, a.counter, epoch, sh);
    if (cnt == 1) eps0 = tot;                                   // nnls.py:187-188
    eps = tot;
    ++cnt;
    if (tot == 0.0 && !normalize) {
      // remaining sweeps are no-ops.  nnls.py:156 keeps looping on `0 >= delta * 0` only when the first sweep already moved
      // nothing (eps0 == 0: it then burns all maxiter sweeps, cnt = maxiter + 1); otherwise the test fails and cnt stays
      if (eps0 == 0.0 && cnt < a.maxiter + 1) cnt = a.maxiter + 1;
      break;
    }
  }
  if (single) store_v(col_of(0));
  if (blockIdx.x == 0 && t == 0) {
    a.result[0] = eps;
    a.result[1] = (double)cnt;
    a.result[2] = (double)zero_row;
    a.result[3] = (double)(cnt - 1);
  }
}


// --------------------------------------------------------------------------------------------
// Chunk-blocked sweep (the production path when no per-row global operation is requested).
//
// Rows are processed VEC (= 16 bytes / sizeof(T)) at a time; the VEC rows of a chunk belong to
// one lane.  Phase 1: every lane forms its partial dot products for the VEC rows against the
// values V had when the chunk started (VEC * NR * C independent FMAs, one shuffle reduction).
// Phase 2: the owning lane walks the VEC rows in order and adds the in-chunk corrections
// G[k,k'] * dV[k'] for k' < k, which restores the exact Gauss-Seidel recurrence
//   dV[k] = max((UtM[k] - UtU[k,:] V - sp) / UtU[k,k], -V[k])            (nnls.py:163/167)
// while the sequential dependency chain is paid once per chunk instead of once per row.
// The row loop is a run-time loop (the code stays inside the instruction cache); the owner
// copies its chunk in and out of the register tile through a switch on the chunk slot.
// --------------------------------------------------------------------------------------------
#define NNFAC_SLOT_CASES(OP) \
  switch (slot) {            \
    case 0: OP(0) break;  case 1: OP(1) break;  case 2: OP(2) break;  case 3: OP(3) break;     \
    case 4: OP(4) break;  case 5: OP(5) break;  case 6: OP(6) break;  case 7: OP(7) break;     \
    case 8: OP(8) break;  case 9: OP(9) break;  case 10: OP(10) break; case 11: OP(11) break;  \
    case 12: OP(12) break; case 13: OP(13) break; case 14: OP(14) break; default: OP(15) break; \
  }

template <typename T, int RP, int L, int C, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) hals_sweep_chunked_kernel(SweepArgs<T> a) {
  constexpr int VEC = VecT<T>::N;
  constexpr int NR = RP / L;
  constexpr int NCH = NR / VEC;
  static_assert(NCH >= 1 && NCH <= 16 && NCH * VEC * L == RP, "bad sweep tiling");
  using V4 = typename VecT<T>::type;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Gs = reinterpret_cast<T*>(smem_raw);
  T* diag = Gs + RP * RP;
  double* sh = reinterpret_cast<double*>(diag + RP);
  T* bs = reinterpret_cast<T*>(sh + 40);

  const int t = threadIdx.x;
  const int lane_l = t % L;
  const int gid = t / L;
  const int GP = blockDim.x / L;
  const int r = a.r;

  for (int idx = t; idx < RP * RP; idx += blockDim.x) {
    const int i = idx / RP, j = idx % RP;
    Gs[idx] = (i < r && j < r) ? a.G[(int64_t)i * a.ld_g + j] : T(0);
  }
  for (int i = t; i < RP; i += blockDim.x) diag[i] = i < r ? a.G[(int64_t)i * a.ld_g + i] : T(0);
  if (a.slab_ld > 0) {
    const int64_t cta_col0 = (int64_t)blockIdx.x * GP * C;
    for (int idx = t; idx < r * a.slab_ld; idx += blockDim.x) {
      const int k = idx / a.slab_ld, c = idx % a.slab_ld;
      bs[idx] = (cta_col0 + c < a.n) ? a.b[(int64_t)k * a.ld_b + cta_col0 + c] : T(0);
    }
  }
  __syncthreads();

  T v[NCH][VEC][C];
  const bool single = (a.nbatch == 1);
  auto col_of = [&](int64_t batch) -> int64_t {
    return ((batch * gridDim.x + blockIdx.x) * (int64_t)GP + gid) * C;
  };
  auto load_v = [&](int64_t col0) {
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int row = VEC * (lane_l + L * i) + e;
#pragma unroll
        for (int c = 0; c < C; ++c)
          v[i][e][c] = (row < r && col0 + c < a.n) ? a.V[(int64_t)row * a.ld_v + col0 + c] : T(0);
      }
  };
  auto store_v = [&](int64_t col0) {
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int row = VEC * (lane_l + L * i) + e;
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (row < r && col0 + c < a.n) a.V[(int64_t)row * a.ld_v + col0 + c] = v[i][e][c];
      }
  };

  unsigned epoch = 0;
  double eps0 = 0.0, eps = 1.0;
  int cnt = 1;
  const int nchunks = (r + VEC - 1) / VEC;
  if (single) load_v(col_of(0));

  while (eps >= a.delta * eps0 && cnt <= a.maxiter) {
    T nd = T(0);
    for (int64_t batch = 0; batch < a.nbatch; ++batch) {
      const int64_t col0 = col_of(batch);
      if (!single) load_v(col0);
#pragma unroll 1
      for (int q = 0; q < nchunks; ++q) {
        const int lo = q % L, slot = q / L, k0 = q * VEC;
        // ---- phase 1: partial dots of the chunk's rows against V as it is now ----
        T acc[VEC][C];
#pragma unroll
        for (int e = 0; e < VEC; ++e)
#pragma unroll
          for (int c = 0; c < C; ++c) acc[e][c] = T(0);
        const T* grow = Gs + k0 * RP + VEC * lane_l;
#pragma unroll
        for (int i2 = 0; i2 < NCH; ++i2) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const V4 g4 = *reinterpret_cast<const V4*>(grow + e * RP + VEC * L * i2);
            const T* gg = reinterpret_cast<const T*>(&g4);
#pragma unroll
            for (int e2 = 0; e2 < VEC; ++e2)
#pragma unroll
              for (int c = 0; c < C; ++c) acc[e][c] = fma(gg[e2], v[i2][e2][c], acc[e][c]);
          }
        }
#pragma unroll
        for (int off = L / 2; off > 0; off >>= 1)
#pragma unroll
          for (int e = 0; e < VEC; ++e)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[e][c] += __shfl_xor_sync(0xffffffffu, acc[e][c], off);
        // ---- phase 2: the owner applies the rows in order ----
        if (lane_l == lo) {
          T w[VEC][C], d[VEC][C];
#define NNFAC_COPY_IN(S)                                        \
  if (S < NCH) {                                                \
    _Pragma("unroll") for (int e = 0; e < VEC; ++e)             \
      _Pragma("unroll") for (int c = 0; c < C; ++c) w[e][c] = v[S < NCH ? S : 0][e][c]; \
  }
          NNFAC_SLOT_CASES(NNFAC_COPY_IN)
#undef NNFAC_COPY_IN
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const int k = k0 + e;
            const T dk = diag[k];
            if (k < r && dk != T(0)) {
              T bk[C];
              if (a.slab_ld > 0) {
                const T* bp = bs + k * a.slab_ld + gid * C;
                if (C * sizeof(T) == 16) {
                  const V4 qv = *reinterpret_cast<const V4*>(bp);
                  const T* qq = reinterpret_cast<const T*>(&qv);
#pragma unroll
                  for (int c = 0; c < C; ++c) bk[c] = qq[c];
                } else {
#pragma unroll
                  for (int c = 0; c < C; ++c) bk[c] = bp[c];
                }
              } else {
                const T* bp = a.b + (int64_t)k * a.ld_b + col0;
#pragma unroll
                for (int c = 0; c < C; ++c) bk[c] = (col0 + c < a.n) ? bp[c] : T(0);
              }
              // in-chunk corrections: rows k0 .. k-1 of this chunk have already moved by d
              T gk[VEC];
              {
                const V4 g4 = *reinterpret_cast<const V4*>(Gs + k * RP + k0);
                const T* gg = reinterpret_cast<const T*>(&g4);
#pragma unroll
                for (int e2 = 0; e2 < VEC; ++e2) gk[e2] = gg[e2];
              }
#pragma unroll
              for (int c = 0; c < C; ++c) {
                T s = acc[e][c];
#pragma unroll
                for (int e2 = 0; e2 < VEC; ++e2)
                  if (e2 < e) s = fma(gk[e2], d[e2][c], s);
                const T cur = w[e][c];
                T dd = (bk[c] - s - a.sp) / dk;
                dd = dd > -cur ? dd : -cur;
                if (col0 + c >= a.n) dd = T(0);
                w[e][c] = cur + dd;
                d[e][c] = dd;
                nd = fma(dd, dd, nd);
              }
            } else {
#pragma unroll
              for (int c = 0; c < C; ++c) d[e][c] = T(0);
            }
          }
#define NNFAC_COPY_OUT(S)                                       \
  if (S < NCH) {                                                \
    _Pragma("unroll") for (int e = 0; e < VEC; ++e)             \
      _Pragma("unroll") for (int c = 0; c < C; ++c) v[S < NCH ? S : 0][e][c] = w[e][c]; \
  }
          NNFAC_SLOT_CASES(NNFAC_COPY_OUT)
#undef NNFAC_COPY_OUT
        }
      }
      if (!single) store_v(col0);
    }
    double tot = block_sum((double)nd, sh), dummy = 0.0;
    grid_reduce(tot, dummy, a.part, a.counter, epoch, sh);
    if (cnt == 1) eps0 = tot;
    eps = tot;
    ++cnt;
    if (tot == 0.0) {
      if (eps0 == 0.0 && cnt < a.maxiter + 1) cnt = a.maxiter + 1;     // see above (nnls.py:156)
      break;
    }
  }
  if (single) store_v(col_of(0));
  if (blockIdx.x == 0 && t == 0) {
    a.result[0] = eps;
    a.result[1] = (double)cnt;
    a.result[2] = -1.0;
    a.result[3] = (double)(cnt - 1);
  }
}

template <typename T, int RP, int L, int C, bool ROWOPS, int MAXT>
int launch(nnfac_ctx* ctx, SweepArgs<T> a, cudaStream_t st) {
  void (*kern)(SweepArgs<T>) = hals_sweep_kernel<T, RP, L, C, true, MAXT>;
  if (!ROWOPS) kern = hals_sweep_chunked_kernel<T, RP, L, C, MAXT>;
  const size_t smem_base = (size_t)(RP * RP + RP) * sizeof(T) + 40 * sizeof(double);
  int per_sm = 0;
  NNFAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_base));
  NNFAC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MAXT, smem_base));
  if (per_sm < 1) {
    nnfac_set_error("hals sweep kernel <RP=%d,L=%d,C=%d> does not fit on an SM", RP, L, C);
    return NNFAC_ERR_UNSUPPORTED;
  }
  const int64_t max_grid = ctx->sm_count;  // one CTA per SM: the barrier cost grows with the grid
  const int64_t groups = ceil_div64(a.n, C);
  // spread the column groups over the SMs, then round the CTA up to whole warps
  int64_t grid = groups * L >= max_grid * 32 ? max_grid : ceil_div64(groups * L, 32);
  if (grid < 1) grid = 1;
  int64_t threads = ceil_div64(ceil_div64(groups, grid) * L, 32) * 32;
  int64_t nbatch = 1;
  if (threads > MAXT) {
    threads = MAXT;
    const int64_t per_cta = (MAXT / L) * (int64_t)C;
    nbatch = ceil_div64(a.n, per_cta * grid);
  }
  if (ROWOPS && nbatch > 1) {
    nnfac_set_error("hals_nnls: normalize/nonzero need n <= %lld columns for rank <= %d on this device (got %lld)",
                    (long long)((MAXT / L) * (int64_t)C * grid), RP, (long long)a.n);
    return NNFAC_ERR_UNSUPPORTED;
  }
  a.nbatch = nbatch;
  // UtM slice in shared memory when it fits next to the Gram (single batch only)
  size_t smem = smem_base;
  a.slab_ld = 0;
  if (nbatch == 1) {
    const int64_t slab_ld = (threads / L) * (int64_t)C;     // multiple of C, so 16-byte rows stay aligned
    const size_t slab = (size_t)a.r * slab_ld * sizeof(T);
    if (smem_base + slab <= (size_t)200 * 1024) { a.slab_ld = (int)slab_ld; smem += slab; }
  }
  NNFAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if ((size_t)(4 * grid) > ctx->red_count) {
    nnfac_set_error("hals_nnls: barrier scratch too small");
    return NNFAC_ERR_UNSUPPORTED;
  }
  const int grc = nnfac_guard_enter(ctx, NNFAC_GUARD_RED, st);
  if (grc) return grc;
  a.part = ctx->red;
  a.counter = ctx->sync;
  NNFAC_CUDA(cudaMemsetAsync(ctx->sync, 0, sizeof(unsigned), st));
  void* params[] = {&a};
  NNFAC_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3((unsigned)threads), params, smem, st));
  ctx->launches++;
  return NNFAC_OK;
}

// Wide tiling (FMA-efficient, for many columns) and narrow tiling (more lanes per column, for few).
template <typename T, int RP> struct Tiling;
template <> struct Tiling<float, 16>  { static constexpr int WL = 1, WC = 4, NL = 4,  NC = 1; };
template <> struct Tiling<float, 32>  { static constexpr int WL = 1, WC = 4, NL = 8,  NC = 1; };
template <> struct Tiling<float, 64>  { static constexpr int WL = 2, WC = 4, NL = 8,  NC = 1; };
template <> struct Tiling<float, 128> { static constexpr int WL = 4, WC = 4, NL = 16, NC = 1; };
template <> struct Tiling<double, 16>  { static constexpr int WL = 1, WC = 2, NL = 4,  NC = 1; };
template <> struct Tiling<double, 32>  { static constexpr int WL = 1, WC = 2, NL = 8,  NC = 1; };
template <> struct Tiling<double, 64>  { static constexpr int WL = 2, WC = 2, NL = 8,  NC = 1; };
template <> struct Tiling<double, 128> { static constexpr int WL = 4, WC = 2, NL = 16, NC = 1; };

template <typename T, int RP>
int run_rank(nnfac_ctx* ctx, SweepArgs<T> a, cudaStream_t st) {
  using TL = Tiling<T, RP>;
  const bool rowops = (a.flags & (NNFAC_HALS_NORMALIZE | NNFAC_HALS_NONZERO)) != 0;
  if (rowops) return launch<T, RP, TL::WL, TL::WC, true, 256>(ctx, a, st);
  const int64_t wide_threads = ceil_div64(a.n, TL::WC) * TL::WL;
  const bool narrow_fits = ceil_div64(ceil_div64(a.n, TL::NC), ctx->sm_count) * TL::NL <= 512;
  if (wide_threads >= (int64_t)ctx->sm_count * 128 || !narrow_fits)
    return launch<T, RP, TL::WL, TL::WC, false, 256>(ctx, a, st);
  return launch<T, RP, TL::NL, TL::NC, false, 512>(ctx, a, st);
}

}  // namespace hals
