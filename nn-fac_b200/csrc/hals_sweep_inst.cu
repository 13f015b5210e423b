// One translation unit per (type, padded rank): the sweep kernels are fully unrolled and slow to
// compile, so the build compiles these in parallel (-DSWEEP_T=float|double -DSWEEP_RP=16|32|64|128).
#include "hals_sweep.cuh"

#define SWEEP_CAT2(a, b) a##b
#define SWEEP_CAT(a, b) SWEEP_CAT2(a, b)
#define SWEEP_NAME SWEEP_CAT(SWEEP_CAT(SWEEP_CAT(nnfac_sweep_, SWEEP_TAG), _), SWEEP_RP)

int SWEEP_NAME(nnfac_ctx* ctx, hals::SweepArgs<SWEEP_T> a, cudaStream_t st) {
  return hals::run_rank<SWEEP_T, SWEEP_RP>(ctx, a, st);
}
