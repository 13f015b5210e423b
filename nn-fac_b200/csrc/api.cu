// Context management and error reporting for the C ABI (include/nnfac_b200.h).
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

static thread_local char g_err[512] = "";

void nnfac_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int nnfac_ws_reserve(nnfac_ctx* ctx, size_t bytes, cudaStream_t st) {
  if (bytes <= ctx->ws_bytes) return NNFAC_OK;
  // growing the scratch buffer is rare: drain the stream that may still be using the old one
  NNFAC_CUDA(cudaStreamSynchronize(st));
  if (ctx->ws) NNFAC_CUDA(cudaFree(ctx->ws));
  ctx->ws = nullptr;
  ctx->ws_bytes = 0;
  size_t want = bytes + (bytes >> 2);
  if (cudaMalloc(&ctx->ws, want) != cudaSuccess) {
    cudaGetLastError();
    nnfac_set_error("workspace allocation of %zu bytes failed", want);
    return NNFAC_ERR_ALLOC;
  }
  ctx->ws_bytes = want;
  return NNFAC_OK;
}

extern "C" {

int nnfac_abi_version(void) { return NNFAC_ABI_VERSION; }

const char* nnfac_last_error(void) { return g_err; }

int nnfac_ctx_create(int device, nnfac_ctx** out) {
  NNFAC_ARG(out != nullptr, "nnfac_ctx_create: out is NULL");
  int ndev = 0;
  NNFAC_CUDA(cudaGetDeviceCount(&ndev));
  NNFAC_ARG(device >= 0 && device < ndev, "nnfac_ctx_create: device %d out of range (%d visible)", device, ndev);
  NNFAC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  NNFAC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    nnfac_set_error("nnfac_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return NNFAC_ERR_UNSUPPORTED;
  }
  nnfac_ctx* c = (nnfac_ctx*)calloc(1, sizeof(nnfac_ctx));
  if (!c) return NNFAC_ERR_ALLOC;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->red_count = 1 << 16;
  NNFAC_CUDA(cudaMalloc(&c->red, c->red_count * sizeof(double)));
  NNFAC_CUDA(cudaMalloc(&c->sync, 64 * sizeof(unsigned)));
  NNFAC_CUDA(cudaMemset(c->sync, 0, 64 * sizeof(unsigned)));
  c->mail_count = (size_t)1 << 18;
  // one extra word behind the mailboxes holds the call generation (advanced on the device, see csrc/tc_sweep.cu)
  NNFAC_CUDA(cudaMalloc(&c->mail, (c->mail_count + 1) * sizeof(unsigned long long)));
  NNFAC_CUDA(cudaMemset(c->mail, 0, (c->mail_count + 1) * sizeof(unsigned long long)));
  c->ws_bytes = (size_t)64 << 20;
  NNFAC_CUDA(cudaMalloc(&c->ws, c->ws_bytes));
  *out = c;
  return NNFAC_OK;
}

int nnfac_ctx_destroy(nnfac_ctx* ctx) {
  if (!ctx) return NNFAC_OK;
  cudaSetDevice(ctx->device);
  cudaFree(ctx->ws);
  cudaFree(ctx->red);
  cudaFree(ctx->sync);
  cudaFree(ctx->mail);
  free(ctx);
  return NNFAC_OK;
}

int nnfac_ctx_sm_count(const nnfac_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
int64_t nnfac_ctx_launch_count(const nnfac_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
