// Tensor-core HALS sweep (fp32 path, rank <= 64): blocked Gauss-Seidel that reproduces the exact
// row-by-row recurrence of nn_fac/update_rules/nnls.py:158-170.
//
// For a block of 16 rows the reference's dot products UtU[k,:] @ V split into
//   (a) the part against V as it stood when the block started  -> one [128 cols x 64] x [64 x 16]
//       product per column tile on tcgen05 (bf16 hi/lo split, 3 MMAs per k-step, fp32 in TMEM), and
//   (b) the in-block corrections UtU[k,k'] * dV[k'] for k' < k   -> 120 FMAs per column whose Gram
//       operands come from __constant__ memory (they are the same for every column).
// That removes ~7/8 of the FP32-FMA work of the plain sweep and all of its shared-memory operand
// traffic.  One thread owns one column of V: 64 fp32 master values in registers for the whole
// call, UtM for that column parked in TMEM, the bf16 operand planes of V in shared memory
// (rewritten 16 rows at a time after each block).  A CTA carries up to 4 column tiles whose
// MMA / update phases interleave.  The per-sweep stop test is the same fixed-order grid reduction
// as the FMA kernel (deterministic, identical on every CTA).
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

using bf16 = __nv_bfloat16;

constexpr int RP = 64;            // padded rank (K extent of the MMA, rows of the register tile)
constexpr int BLK = 16;           // rows per Gauss-Seidel block (UMMA N)
constexpr int NBLK = RP / BLK;
constexpr int TILE = 128;         // columns per tile (UMMA M)
constexpr int MAX_TILES = 4;
constexpr int UPD_THREADS = MAX_TILES * TILE;   // 512
constexpr int NTHREADS = UPD_THREADS;
constexpr int TMEM_PER_TILE = 128;              // 64 columns UtM + 4 x 16 columns of block dot products
constexpr uint32_t PLANE_BYTES = TILE * RP * sizeof(bf16);   // 16 KiB
constexpr uint32_t G_PLANE_BYTES = RP * RP * sizeof(bf16);   // 8 KiB
constexpr uint32_t G_BLOCK_BYTES = 3 * BLK * 128;             // per block: 16 rows hi, 16 mid, 16 lo
constexpr int NPLANES = 3;

struct SweepConst {
  float nh[NBLK][BLK][BLK];     // -UtU[k][l] / UtU[k][k] inside the diagonal blocks (row k, column l)
  float invd[RP];               // 1 / UtU[k][k], 0 when the diagonal entry is 0 or k >= r
};
__constant__ SweepConst c_sw;

__global__ void sweep_prep_kernel(const float* __restrict__ G, int64_t ld_g, int r, SweepConst* out) {
  for (int idx = threadIdx.x; idx < NBLK * BLK * BLK; idx += blockDim.x) {
    const int B = idx / (BLK * BLK), e = (idx / BLK) % BLK, e2 = idx % BLK;
    const int k = B * BLK + e, l = B * BLK + e2;
    const float dk = k < r ? G[(int64_t)k * ld_g + k] : 0.f;
    out->nh[B][e][e2] = (k < r && l < r && dk != 0.f) ? -G[(int64_t)k * ld_g + l] * (1.f / dk) : 0.f;
  }
  for (int k = threadIdx.x; k < RP; k += blockDim.x) {
    const float d = k < r ? G[(int64_t)k * ld_g + k] : 0.f;
    out->invd[k] = d != 0.f ? 1.f / d : 0.f;
  }
}

struct TcSweepArgs {
  const float* b;   // UtM r x n
  const float* G;   // UtU r x r
  float* V;         // r x n, in place
  int64_t ld_b, ld_g, ld_v, n;
  int r, maxiter, cols_per_cta;
  double delta;
  float sp;
  double* part;
  unsigned* counter;
  double* result;
};

using tc::tmem_st16;
using tc::tmem_st_wait;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);   // .x = a (low half), .y = b
  return *reinterpret_cast<const uint32_t*>(&p);
}

// Write 8 consecutive K-elements (one 16-byte chunk) of row `row` into the three K-major SW128 planes
// (hi, mid, lo: x = hi + mid + lo exactly for normal fp32 values, i.e. fp32 operands for the tensor core).
// Two values per conversion (cvt.rn.bf16x2.f32); the halves are widened back with a shift / a mask.
__device__ __forceinline__ void store_chunk(uint8_t* plane_hi, int row, int chunk, const float* x, uint32_t plane_stride) {
  uint32_t h[4], m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = x[2 * i], b = x[2 * i + 1];
    h[i] = pack_bf16(a, b);
    const float a1 = a - __uint_as_float(h[i] << 16), b1 = b - __uint_as_float(h[i] & 0xffff0000u);
    m[i] = pack_bf16(a1, b1);
    const float a2 = a1 - __uint_as_float(m[i] << 16), b2 = b1 - __uint_as_float(m[i] & 0xffff0000u);
    l[i] = pack_bf16(a2, b2);
  }
  uint8_t* p = plane_hi + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
  *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(p + plane_stride) = make_uint4(m[0], m[1], m[2], m[3]);
  *reinterpret_cast<uint4*>(p + 2 * plane_stride) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Issue the tensor-core part of block B for one tile.  Operands are 3-term bf16 splits; six of the
// nine partial products are kept (the dropped ones are below 2^-24):
//   cols [0,48)  : V_hi  x [G_hi ; G_mid ; G_lo]   (N = 48)
//   cols [0,32)  : V_mid x [G_hi ; G_mid]          (N = 32, accumulated on top)
//   cols [48,64) : V_lo  x  G_hi                   (N = 16, independent chain)
// The update threads add the four 16-column groups.
__device__ __forceinline__ void issue_block_mma(uint64_t a_hi, uint64_t a_mid, uint64_t a_lo, uint64_t g_blk, uint32_t d,
                                                int nks, uint32_t idesc48, uint32_t idesc32, uint32_t idesc16,
                                                uint64_t* bar) {
#pragma unroll
  for (int ks = 0; ks < NBLK; ++ks) {
    if (ks < nks) {
      const uint64_t koff = (uint64_t)(ks * 2);               // 32 bytes >> 4 inside the 128-byte row
      tc::umma_bf16(d, a_hi + koff, g_blk + koff, idesc48, ks != 0);
      tc::umma_bf16(d + 48, a_lo + koff, g_blk + koff, idesc16, ks != 0);
      tc::umma_bf16(d, a_mid + koff, g_blk + koff, idesc32, true);
    }
  }
  tc::umma_commit(bar);
}

__global__ void __launch_bounds__(NTHREADS, 1) tc_sweep_kernel(const TcSweepArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* g_planes = smem;                                            // per block: 16 rows hi, 16 mid, 16 lo (6 KiB)
  uint8_t* v_planes = smem + NPLANES * G_PLANE_BYTES;                  // [tile][hi|mid|lo][16 KiB]
  uint8_t* tail = v_planes + (size_t)MAX_TILES * NPLANES * PLANE_BYTES;
  uint64_t* s_full = reinterpret_cast<uint64_t*>(tail);                // [MAX_TILES]
  double* red = reinterpret_cast<double*>(s_full + MAX_TILES);         // [24]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(red + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = a.r;
  const int nblk = (r + BLK - 1) / BLK;
  const int64_t cta_col0 = (int64_t)blockIdx.x * a.cols_per_cta;
  int64_t cta_cols = a.n - cta_col0;
  if (cta_cols > a.cols_per_cta) cta_cols = a.cols_per_cta;
  const int ntiles = (int)((cta_cols + TILE - 1) / TILE);

  if (warp == 0) {
    if (lane == 0) {
      for (int t = 0; t < MAX_TILES; ++t) tc::mbar_init(&s_full[t], 1);
      tc::fence_barrier_init();
    }
    __syncwarp();
    tc::tmem_alloc(tmem_slot, 512);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // one thread per column; 4 warps (= 128 TMEM lanes) per tile; the first lane of a tile issues its MMAs
  const int tile = warp >> 2, q = warp & 3;
  const bool active = tile < ntiles;
  const bool issuer = active && q == 0 && lane == 0;
  const int row = q * 32 + lane;                                       // row of the A tile == TMEM lane
  const int64_t col = cta_col0 + (int64_t)tile * TILE + row;
  const bool valid = active && (int64_t)tile * TILE + row < cta_cols;
  uint8_t* vh = v_planes + (size_t)tile * NPLANES * PLANE_BYTES;
  const uint32_t t_b = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tile * TMEM_PER_TILE);
  const uint32_t t_s = t_b + RP;
  const uint32_t d_tile = tmem_base + (uint32_t)(tile * TMEM_PER_TILE + RP);
  const uint32_t idesc48 = tc::umma_idesc_bf16(TILE, 3 * BLK), idesc32 = tc::umma_idesc_bf16(TILE, 2 * BLK),
                 idesc16 = tc::umma_idesc_bf16(TILE, BLK);
  const uint64_t a_hi = tc::umma_desc_k_sw128(tc::smem_u32(vh)), a_mid = tc::umma_desc_k_sw128(tc::smem_u32(vh) + PLANE_BYTES),
                 a_lo = tc::umma_desc_k_sw128(tc::smem_u32(vh) + 2 * PLANE_BYTES);
  const uint64_t g_desc = tc::umma_desc_k_sw128(tc::smem_u32(g_planes));

  // Gram operand planes (K-major, 128B swizzle), block B at byte offset B * 4096: rows 0-15 hi, 16-31 lo
  {
    const int k = threadIdx.x >> 3, c = threadIdx.x & 7;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int l = c * 8 + i;
      x[i] = (k < r && l < r) ? a.G[(int64_t)k * a.ld_g + l] : 0.f;
    }
    uint8_t* blk = g_planes + (size_t)(k / BLK) * G_BLOCK_BYTES;
    store_chunk(blk, k % BLK, c, x, BLK * 128);    // (k % 16) & 7 == k & 7: the swizzle phase is preserved
  }

  float v[RP];
  if (active) {
#pragma unroll
    for (int k = 0; k < RP; ++k) v[k] = (valid && k < r) ? a.V[(int64_t)k * a.ld_v + col] : 0.f;
    // UtM of this column -> TMEM (stays there for the whole call)
#pragma unroll
    for (int c0 = 0; c0 < RP; c0 += 16) {
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        w[j] = __float_as_uint((valid && c0 + j < r) ? a.b[(int64_t)(c0 + j) * a.ld_b + col] : 0.f);
      tmem_st16(t_b + c0, w);
    }
    tmem_st_wait();
#pragma unroll
    for (int c = 0; c < RP / 8; ++c) store_chunk(vh, row, c, &v[c * 8], PLANE_BYTES);
  }
  tc::fence_proxy_async_smem();
  tc::tcgen05_fence_before();
  named_bar_sync(1, UPD_THREADS);                                      // Gram and V planes complete
  if (issuer) {
    tc::tcgen05_fence_after();
    issue_block_mma(a_hi, a_mid, a_lo, g_desc, d_tile, nblk, idesc48, idesc32, idesc16, &s_full[tile]);
  }

  // One Gauss-Seidel block (16 rows) of this thread's column: wait for the tensor-core part, run the
  // in-block recurrence, rewrite the operand planes of the 16 rows and hand the next block to the tensor
  // core.  Returns the squared step of the block (nnls.py:170).
  uint32_t s_phase = 0;
  const float sp_t = valid ? a.sp : 0.f;
#ifdef SWEEP_PROF
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PROF_T(i) { const long long t__ = clock64(); prof[i] += t__ - tp; tp = t__; }
#define PROF_START long long tp = clock64();
#else
#define PROF_T(i)
#define PROF_START
#endif
  auto block_update = [&](auto Bc) -> float {
    constexpr int B = decltype(Bc)::value;
    float nd = 0.f;
    PROF_START
    uint32_t bu[16];
    tc::tmem_ld16(t_b + B * BLK, bu);                               // UtM of the block: does not wait for the MMA
    tc::mbar_wait(&s_full[tile], s_phase);
    s_phase ^= 1;
    tc::tcgen05_fence_after();
    PROF_T(0)
    // u[e] = (UtM[k] - UtU[k,:] V - sp) / UtU[k,k] with V as it stood when the block started (tensor-core part)
    float u[BLK];
    {
      uint32_t s0[16], s1[16];
      tc::tmem_ld16(t_s + 16, s0);
      tc::tmem_ld16(t_s + 48, s1);
      tc::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < BLK; ++e) u[e] = __uint_as_float(s0[e]) + __uint_as_float(s1[e]);
      tc::tmem_ld16(t_s + 32, s0);
      tc::tmem_ld16(t_s, s1);
      tc::tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < BLK; ++e) u[e] = __uint_as_float(s1[e]) + (u[e] + __uint_as_float(s0[e]));
#pragma unroll
      for (int e = 0; e < BLK; ++e) u[e] = (__uint_as_float(bu[e]) - u[e] - sp_t) * c_sw.invd[B * BLK + e];
    }
    tc::tcgen05_fence_before();
    PROF_T(1)
    // In-block Gauss-Seidel recurrence (nnls.py:158-170) on the scaled residuals: the step of row k is
    // max(u, -V[k]) and every later row of the block sees it through -UtU[k2][k] / UtU[k2][k2].
    // Two dependent instructions per row (FMNMX, FFMA) instead of the six of the literal formula.
#pragma unroll
    for (int e = 0; e < BLK; ++e) {
      const int k = B * BLK + e;
      const float cur = v[k];
      const float lb = c_sw.invd[k] != 0.f ? -cur : 0.f;            // zero diagonal: row skipped (nnls.py:160), u = 0
      const float dd = fmaxf(u[e], lb);                             // nnls.py:163/167
      if (e + 1 < BLK) u[e + 1] = fmaf(c_sw.nh[B][e + 1][e], dd, u[e + 1]);
      v[k] = cur + dd;
      nd = fmaf(dd, dd, nd);                                        // nnls.py:170
#pragma unroll
      for (int e2 = e + 2; e2 < BLK; ++e2) u[e2] = fmaf(c_sw.nh[B][e2][e], dd, u[e2]);
    }
    PROF_T(2)
    store_chunk(vh, row, 2 * B, &v[B * BLK], PLANE_BYTES);
    store_chunk(vh, row, 2 * B + 1, &v[B * BLK + 8], PLANE_BYTES);
    PROF_T(3)
    tc::fence_proxy_async_smem();
    tc::tcgen05_fence_before();
    named_bar_sync(2 + tile, TILE);                                 // this tile's planes are rewritten
    PROF_T(4)
    if (issuer) {
      // next block (block 0 of the next sweep after the last one: if the stop test ends the solve its
      // result is simply dropped)
      const int nb = (B + 1 < nblk) ? B + 1 : 0;
      tc::tcgen05_fence_after();
      issue_block_mma(a_hi, a_mid, a_lo, g_desc + (uint64_t)(nb * (int)G_BLOCK_BYTES >> 4), d_tile, nblk, idesc48, idesc32,
                      idesc16, &s_full[tile]);
    }
    PROF_T(5)
    return nd;
  };

  double eps0 = 0.0, eps = 1.0;
  int cnt = 1;
  unsigned epoch = 0;
  // Block 0 of the coming sweep is run SPECULATIVELY while the grid-wide sum of the finished sweep is in
  // flight (the stop test of nnls.py:156 needs that sum): `bk` keeps the 16 values it overwrites.
  bool have_spec = false;
  float nd_spec = 0.f;
  float bk[BLK];
  while (true) {
    float nd = have_spec ? nd_spec : 0.f;
    if (active) {
      if (!have_spec) nd += block_update(std::integral_constant<int, 0>{});
      if (nblk > 1) nd += block_update(std::integral_constant<int, 1>{});
      if (nblk > 2) nd += block_update(std::integral_constant<int, 2>{});
      if (nblk > 3) nd += block_update(std::integral_constant<int, 3>{});
    }
    // ---- sum of squared steps over the whole grid, fixed order: post this CTA's partial ... ----
    double t = warp_sum((double)nd);
    if (lane == 0) red[warp] = t;
    named_bar_sync(1, UPD_THREADS);
    const unsigned nb = gridDim.x;
    if (threadIdx.x == 0) {
      double sum = 0.0;
      for (int w = 0; w < 16; ++w) sum += red[w];
      if (nb > 1) {
        double* slot = a.part + (size_t)(epoch & 1u) * nb;
        slot[blockIdx.x] = sum;
        __threadfence();
        atomicAdd(a.counter, 1u);
      } else {
        red[16] = sum;
      }
    }
    // ---- ... run block 0 of the next sweep while the other CTAs arrive ... ----
    have_spec = false;
    if (cnt + 1 <= a.maxiter) {
      have_spec = true;
      nd_spec = 0.f;
      if (active) {
#pragma unroll
        for (int e = 0; e < BLK; ++e) bk[e] = v[e];
        nd_spec = block_update(std::integral_constant<int, 0>{});
      }
    }
    // ---- ... then collect the total ----
#ifdef SWEEP_PROF
    const long long tg0 = clock64();
#endif
    if (nb > 1) {
      if (warp == 0) {
        const unsigned target = (epoch + 1u) * nb;
        if (lane == 0) {
          while (ld_relaxed_u32(a.counter) < target) {}
          __threadfence();
        }
        __syncwarp();
        const double* slot = a.part + (size_t)(epoch & 1u) * nb;
        double sum = 0.0;
        for (unsigned i = lane; i < nb; i += 32) sum += __ldcg(slot + i);
        sum = warp_sum(sum);
        if (lane == 0) red[16] = sum;
      }
    }
    named_bar_sync(1, UPD_THREADS);
#ifdef SWEEP_PROF
    prof[6] += clock64() - tg0;
#endif
    const double tot = red[16];
    ++epoch;
    if (cnt == 1) eps0 = tot;
    eps = tot;
    ++cnt;
    bool stop = !(eps >= a.delta * eps0 && cnt <= a.maxiter);
    if (tot == 0.0) {                                                   // further sweeps are no-ops (nnls.py:156)
      if (cnt < a.maxiter + 1) cnt = a.maxiter + 1;
      stop = true;
    }
    if (stop) {
      if (have_spec && active) {
#pragma unroll
        for (int e = 0; e < BLK; ++e) v[e] = bk[e];                     // undo the speculative block
      }
      break;
    }
  }
  if (active) {
    tc::mbar_wait(&s_full[tile], s_phase);                              // drain the speculative block
    tc::tcgen05_fence_after();
  }
  if (valid) {
#pragma unroll
    for (int k = 0; k < RP; ++k)
      if (k < r) a.V[(int64_t)k * a.ld_v + col] = v[k];
  }
#ifdef SWEEP_PROF
  if (blockIdx.x == 1 && (threadIdx.x == 0 || threadIdx.x == 32)) {
    printf("sweep prof (cycles/sweep) thr %d: mma_wait %lld tmem_ld %lld chain %lld split_store %lld fence_bar %lld issue %lld grid %lld\n",
           threadIdx.x, prof[0] / (cnt - 1), prof[1] / (cnt - 1), prof[2] / (cnt - 1), prof[3] / (cnt - 1), prof[4] / (cnt - 1),
           prof[5] / (cnt - 1), prof[6] / (cnt - 1));
  }
#endif
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.result[0] = eps;
    a.result[1] = (double)cnt;
    a.result[2] = -1.0;
    a.result[3] = (double)(cnt - 1);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace

// Returns NNFAC_ERR_UNSUPPORTED (without setting an error) when the shape is outside this kernel's
// envelope, so that the caller can use the FMA kernel instead.
int nnfac_tc_sweep_try(nnfac_ctx* ctx, const float* UtM, int64_t ld_utm, const float* UtU, int64_t ld_utu, float* V,
                       int64_t ld_v, int r, int64_t n, int maxiter, double delta, double sparsity, double* result,
                       cudaStream_t st) {
  if (r > RP || maxiter < 1) return NNFAC_ERR_UNSUPPORTED;
  int64_t cols = ceil_div64(n, ctx->sm_count);
  cols = ceil_div64(cols, 32) * 32;
  if (cols > MAX_TILES * TILE) return NNFAC_ERR_UNSUPPORTED;
  const int64_t grid = ceil_div64(n, cols);
  static SweepConst* staging = nullptr;   // one per process is enough: calls are stream-ordered per context
  if (!staging) NNFAC_CUDA(cudaMalloc(&staging, sizeof(SweepConst)));
  sweep_prep_kernel<<<1, 256, 0, st>>>(UtU, ld_utu, r, staging);
  NNFAC_LAUNCH_CHECK(ctx);
  NNFAC_CUDA(cudaMemcpyToSymbolAsync(c_sw, staging, sizeof(SweepConst), 0, cudaMemcpyDeviceToDevice, st));
  TcSweepArgs a;
  a.b = UtM; a.G = UtU; a.V = V; a.ld_b = ld_utm; a.ld_g = ld_utu; a.ld_v = ld_v; a.n = n;
  a.r = r; a.maxiter = maxiter; a.cols_per_cta = (int)cols; a.delta = delta; a.sp = (float)sparsity;
  a.part = ctx->red; a.counter = ctx->sync; a.result = result;
  const size_t smem = NPLANES * G_PLANE_BYTES + (size_t)MAX_TILES * NPLANES * PLANE_BYTES + 512;
  NNFAC_CUDA(cudaFuncSetAttribute(tc_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NNFAC_CUDA(cudaMemsetAsync(ctx->sync, 0, sizeof(unsigned), st));
  void* params[] = {&a};
  NNFAC_CUDA(cudaLaunchCooperativeKernel((const void*)tc_sweep_kernel, dim3((unsigned)grid), dim3(NTHREADS), params, smem, st));
  ctx->launches++;
  return NNFAC_OK;
}
