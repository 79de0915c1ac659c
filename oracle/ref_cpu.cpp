// Host build of the UNMODIFIED reference translation unit (TEST INFRASTRUCTURE, see oracle/build_ref.sh).
// The reference sources are compiled from where they lie (-I$REF/mpm/csrc); nothing is copied.
// The only intervention: the reference's launch macro (mpm/csrc/common.h:20-25, `kernel<<<...>>> args`) is
// re-pointed at a host loop that calls the same kernel body once per emulated thread.
#include "common.h"  // reference header; #pragma once makes the TU's own include of it a no-op
#include <omp.h>

thread_local shim_dim3 blockIdx, threadIdx;
shim_dim3 blockDim, gridDim;

#undef launch_kernel
#define launch_kernel(kernel, dim, stream, args)                      \
  {                                                                   \
    const int n_threads_ = 256;                                       \
    const int n_blocks_ = ((dim) + n_threads_ - 1) / n_threads_;      \
    blockDim.x = n_threads_;                                          \
    gridDim.x = n_blocks_;                                            \
    _Pragma("omp parallel for schedule(static)")                      \
    for (int b_ = 0; b_ < n_blocks_; ++b_) {                          \
      blockIdx.x = b_;                                                \
      for (int t_ = 0; t_ < n_threads_; ++t_) {                       \
        threadIdx.x = t_;                                             \
        kernel args;                                                  \
      }                                                               \
    }                                                                 \
  }

#include "integrator.cu"

extern "C" int ref_cpu_num_threads() { return omp_get_max_threads(); }
extern "C" void ref_cpu_set_num_threads(int n) { omp_set_num_threads(n); }
