// Exchange steps of the column-sharded path over peer-mapped memory (NVLink / NVSwitch), one process per GPU.
//
// Per outer iteration the sharded NMF has one exchange on the U side: the partial cross products V_p X_p^T (r x m) and
// Grams of all ranks are summed, rank p needing only ITS m/P columns (a reduce-scatter), and after the slice solves
// every rank needs all slices (an all-gather).  With NCCL each of the two costs 0.1-0.2 ms at C2 on 8 GPUs whatever the
// size (16.8 MB): launch + protocol latency.  Here every rank owns an exchange region in device memory (cudaMalloc,
// exported with CUDA IPC, mapped by every peer):
//     [ flags | stage: r x (len + t) | send: r x (chunk + t) ]
// * the X pass' reduction kernel writes the rank's partial result straight into its own `stage` (tail columns: the partial
//   Gram / row sums), then `post` raises a sequence number in every peer's flag block (one st.release.sys each);
// * `pull_reduce` waits for the flags and sums, in rank order, the columns this rank owns out of all P stages -- loads over
//   NVLink, the reduce-scatter and the reduction in ONE kernel, deterministic;
// * the slice solve writes into `send`; after the next `post`, the kernel that installs the new factor in the plan reads the
//   P slices directly from the peers' `send` regions (nnfac_nmf_plan_set_factor_pulled): the all-gather, the un-permute and
//   the operand-plane construction in ONE kernel.
// Nothing is overwritten early: a rank rewrites its stage only after its own install, which waited for every peer's second
// post (hence for every peer's pull); it rewrites its send only after its own pull of the next iteration, which waited for
// every peer's first post of that iteration (hence for every peer's install).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

struct nnfac_xchg {
  nnfac_ctx* ctx;
  void* region;                 // local allocation
  size_t bytes, stage_off, send_off;
  int world, rank;
  void* peer[NNFAC_MAX_PEERS];  // mapped regions (peer[rank] == region)
  unsigned long long seq[2];    // posts made so far per phase
};

namespace {

constexpr size_t FLAG_BYTES = 256;   // [phase][rank] u64

struct PeerPtrs {
  const float* p[NNFAC_MAX_PEERS];
};
struct FlagPtrs {
  unsigned long long* p[NNFAC_MAX_PEERS];
};

__global__ void xchg_post_kernel(FlagPtrs flags, int world, int phase, int my_rank, unsigned long long seq) {
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flags.p[threadIdx.x] + phase * NNFAC_MAX_PEERS + my_rank), "l"(seq) : "memory");
  }
}

__global__ void xchg_wait_kernel(const unsigned long long* flags, int world, int phase, unsigned long long seq) {
  if ((int)threadIdx.x < world) {
    const unsigned long long* f = flags + phase * NNFAC_MAX_PEERS + threadIdx.x;
    unsigned long long v;
    unsigned long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (++spins > (1ull << 27)) __trap();     // ~1 minute: a protocol bug fails the launch instead of hanging the GPU
    } while (v < seq);
  }
}

// out[k][c] = sum_q src_q[k * pitch + lo + c]  (c < ncols; columns ncols .. out_cols-1-tail are zero padding), and
// out[k][tail_col + j] = sum_q src_q[k * pitch + len + j]  (j < tail); fixed order q = 0 .. world-1.
__global__ void __launch_bounds__(256) xchg_pull_reduce_kernel(PeerPtrs src, int world, int r, int64_t pitch, int64_t lo, int64_t ncols,
                                                               int64_t len, int tail, float* __restrict__ out, int64_t ld_out,
                                                               int64_t tail_col) {
  const int64_t width = tail_col + tail, total = (int64_t)r * width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / width, c = i - k * width;
    float s = 0.f;
    if (c < ncols) {
      for (int q = 0; q < world; ++q) s += __ldcv(src.p[q] + k * pitch + lo + c);
    } else if (c >= tail_col) {
      for (int q = 0; q < world; ++q) s += __ldcv(src.p[q] + k * pitch + len + (c - tail_col));
    }
    out[k * ld_out + c] = s;
  }
}

}  // namespace

extern "C" {

int nnfac_xchg_create(nnfac_ctx* ctx, int64_t stage_floats, int64_t send_floats, nnfac_xchg** out) {
  NNFAC_ARG(ctx && out && stage_floats > 0 && send_floats > 0, "nnfac_xchg_create: bad argument");
  NNFAC_CUDA(cudaSetDevice(ctx->device));
  nnfac_xchg* x = (nnfac_xchg*)calloc(1, sizeof(nnfac_xchg));
  if (!x) return NNFAC_ERR_ALLOC;
  x->ctx = ctx;
  x->stage_off = FLAG_BYTES;
  x->send_off = x->stage_off + (((size_t)stage_floats * sizeof(float) + 255) & ~(size_t)255);
  x->bytes = x->send_off + (((size_t)send_floats * sizeof(float) + 255) & ~(size_t)255);
  if (cudaMalloc(&x->region, x->bytes) != cudaSuccess) {
    cudaGetLastError();
    nnfac_set_error("nnfac_xchg_create: out of device memory (%zu bytes)", x->bytes);
    free(x);
    return NNFAC_ERR_ALLOC;
  }
  NNFAC_CUDA(cudaMemset(x->region, 0, x->bytes));
  NNFAC_CUDA(cudaDeviceSynchronize());
  x->world = 1;
  x->peer[0] = x->region;
  *out = x;
  return NNFAC_OK;
}

int nnfac_xchg_export(nnfac_xchg* x, void* handle_out) {
  NNFAC_ARG(x && handle_out, "nnfac_xchg_export: NULL argument");
  cudaIpcMemHandle_t h;
  NNFAC_CUDA(cudaIpcGetMemHandle(&h, x->region));
  memcpy(handle_out, &h, sizeof(h));
  return NNFAC_OK;
}

int nnfac_xchg_attach(nnfac_xchg* x, int world, int rank, const void* handles) {
  NNFAC_ARG(x && handles && world >= 1 && world <= NNFAC_MAX_PEERS && rank >= 0 && rank < world, "nnfac_xchg_attach: bad argument");
  NNFAC_CUDA(cudaSetDevice(x->ctx->device));
  for (int q = 0; q < world; ++q) {
    if (q == rank) { x->peer[q] = x->region; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)q * sizeof(h), sizeof(h));
    NNFAC_CUDA(cudaIpcOpenMemHandle(&x->peer[q], h, cudaIpcMemLazyEnablePeerAccess));
  }
  x->world = world;
  x->rank = rank;
  return NNFAC_OK;
}

int nnfac_xchg_destroy(nnfac_xchg* x) {
  if (!x) return NNFAC_OK;
  cudaSetDevice(x->ctx->device);
  for (int q = 0; q < x->world; ++q)
    if (q != x->rank && x->peer[q]) cudaIpcCloseMemHandle(x->peer[q]);
  cudaFree(x->region);
  free(x);
  return NNFAC_OK;
}

// local device pointer of the stage (which = 0) or send (which = 1) buffer
void* nnfac_xchg_ptr(nnfac_xchg* x, int which) {
  if (!x) return nullptr;
  return (char*)x->region + (which == 0 ? x->stage_off : x->send_off);
}

// "what this rank wrote into its stage / send buffer so far (on `stream`) is complete": raises the sequence number of
// `phase` (0: stage, 1: send) in every rank's flag block.
int nnfac_xchg_post(nnfac_xchg* x, int phase, void* stream) {
  NNFAC_ARG(x && (phase == 0 || phase == 1), "nnfac_xchg_post: bad argument");
  FlagPtrs f;
  for (int q = 0; q < NNFAC_MAX_PEERS; ++q) f.p[q] = q < x->world ? (unsigned long long*)x->peer[q] : nullptr;
  xchg_post_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, x->world, phase, x->rank, ++x->seq[phase]);
  NNFAC_LAUNCH_CHECK(x->ctx);
  return NNFAC_OK;
}

// every rank has made as many posts on `phase` as this rank (call after the own post)
int nnfac_xchg_wait(nnfac_xchg* x, int phase, void* stream) {
  NNFAC_ARG(x && (phase == 0 || phase == 1), "nnfac_xchg_wait: bad argument");
  xchg_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned long long*)x->region, x->world, phase, x->seq[phase]);
  NNFAC_LAUNCH_CHECK(x->ctx);
  return NNFAC_OK;
}

// Sum over the ranks of the columns [lo, lo + ncols) and of the `tail` columns behind column `len` of every rank's buffer
// `which` ([r x pitch] each): out (r x ld_out) receives the columns at 0 and the tail at tail_col.  Call after nnfac_xchg_wait.
int nnfac_xchg_pull_reduce(nnfac_xchg* x, int which, float* out, int64_t ld_out, int r, int64_t pitch, int64_t lo, int64_t ncols,
                           int64_t len, int tail, int64_t tail_col, void* stream) {
  NNFAC_ARG(x && out && (which == 0 || which == 1) && r > 0 && ncols >= 0 && tail >= 0 && tail_col >= ncols && ld_out >= tail_col + tail &&
                lo >= 0 && lo + ncols <= len && len + tail <= pitch, "nnfac_xchg_pull_reduce: bad argument");
  PeerPtrs src;
  const size_t off = which == 0 ? x->stage_off : x->send_off;
  for (int q = 0; q < NNFAC_MAX_PEERS; ++q) src.p[q] = q < x->world ? (const float*)((const char*)x->peer[q] + off) : nullptr;
  const int64_t total = (int64_t)r * (tail_col + tail);
  const int64_t want = ceil_div64(total, 256);
  const int grid = (int)(want < (int64_t)x->ctx->sm_count * 8 ? (want < 1 ? 1 : want) : (int64_t)x->ctx->sm_count * 8);
  xchg_pull_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, x->world, r, pitch, lo, ncols, len, tail, out, ld_out, tail_col);
  NNFAC_LAUNCH_CHECK(x->ctx);
  return NNFAC_OK;
}

// peer q's send buffer as mapped here (for nnfac_nmf_plan_set_factor_pulled)
const float* nnfac_xchg_peer_send(const nnfac_xchg* x, int q) {
  if (!x || q < 0 || q >= x->world) return nullptr;
  return (const float*)((const char*)x->peer[q] + x->send_off);
}
int nnfac_xchg_world(const nnfac_xchg* x) { return x ? x->world : 0; }

}  // extern "C"
