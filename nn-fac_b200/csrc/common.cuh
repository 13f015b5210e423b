// Shared host/device helpers for the nnfac_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/nnfac_b200.h"

// Scratch of a context that several kernels share (split-K workspace, reduction partials, the sweep's constant bank and
// mailboxes).  A context is used from one stream at a time; when a call arrives on ANOTHER stream than the last user of
// the same scratch, that stream is first made to wait for the last user (nnfac_guard_enter), so calls on different
// streams are ordered instead of racing.
enum { NNFAC_GUARD_WS = 0, NNFAC_GUARD_RED = 1, NNFAC_GUARD_SWEEP = 2, NNFAC_NGUARD = 3 };
struct nnfac_guard {
  cudaStream_t last;
  cudaEvent_t ev;
  int valid;
};

// Ranks of a collective HALS solve (one process per GPU): board[q] is rank q's stop-scalar board as mapped into this
// process (cudaIpcOpenMemHandle; board[rank] is the local allocation).
#define NNFAC_MAX_PEERS 8
struct nnfac_peer_group {
  int world, rank;
  void* board[NNFAC_MAX_PEERS];
};

struct nnfac_ctx {
  int device;
  int sm_count;
  int64_t launches;
  void* ws;         // scratch (split-K partials, reduction partials)
  size_t ws_bytes;
  double* red;      // small scratch for two-stage reductions / sweep barrier partials
  size_t red_count;
  unsigned* sync;   // grid-barrier counters (zeroed before each cooperative launch)
  unsigned long long* mail;   // tagged mailboxes of the tensor-core HALS sweep (zeroed once; tags carry a call generation
  size_t mail_count;          // that lives in device memory, in the word behind the last mailbox: mail[mail_count])
  void* sweep_const[2];       // device address of the sweep's __constant__ bank for padded rank 64 / 128 (per device)
  nnfac_guard guard[NNFAC_NGUARD];
  void* board;                // local stop-scalar board of collective solves (cudaMalloc, IPC-exported)
  nnfac_peer_group peers;
  int collective;             // the next tensor-core sweeps are collective over `peers`
  int64_t collective_n[NNFAC_MAX_PEERS];   // columns of every rank's slice
  unsigned collective_gen;    // collective call counter (advances identically on every rank)
  int sweep_lag;              // variant of the tensor-core sweep's stop test: -1 environment / auto, 0 plain, 1 auto, 2 lagged wherever possible
};

struct nnfac_xchg;      // peer-mapped exchange region (csrc/peer_xchg.cu)

void nnfac_set_error(const char* fmt, ...);

// operand planes the tensor-core HALS sweep can write for the NMF plan (all bf16; see csrc/tc_sweep.cu)
struct nnfac_sweep_planes {
  void *fh, *fl;      // [r_pad x ld_plane] K-major hi / lo
  void *rowh, *rowl;  // [n x row_pitch] rank-contiguous hi / lo (NULL: not wanted)
  int64_t ld_plane;
  int r_pad;
  int row_pitch;      // 64 or 128
};
int nnfac_tc_sweep_run(nnfac_ctx* ctx, const float* UtM, int64_t ld_utm, const float* UtU, int64_t ld_utu, const float* Vin,
                       int64_t ld_vin, float* V, int64_t ld_v, int r, int64_t n, int maxiter, double delta, double sparsity,
                       double* result, const nnfac_sweep_planes* planes, cudaStream_t st, int nsplit, int64_t split_stride);
int nnfac_ws_reserve(nnfac_ctx* ctx, size_t bytes, cudaStream_t st);
int nnfac_guard_enter(nnfac_ctx* ctx, int which, cudaStream_t st);

#define NNFAC_CUDA(call)                                                                   \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      nnfac_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return NNFAC_ERR_CUDA;                                                               \
    }                                                                                      \
  } while (0)

#define NNFAC_ARG(cond, ...)      \
  do {                            \
    if (!(cond)) {                \
      nnfac_set_error(__VA_ARGS__); \
      return NNFAC_ERR_ARG;       \
    }                             \
  } while (0)

#define NNFAC_LAUNCH_CHECK(ctx)                 \
  do {                                          \
    (ctx)->launches++;                          \
    NNFAC_CUDA(cudaGetLastError());             \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device reductions ---------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; result valid in every thread. `sh` must hold >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = lane < nw ? sh[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}
__device__ __forceinline__ double block_max(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = lane < nw ? sh[lane] : -1.0e300;
    t = warp_max(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}
