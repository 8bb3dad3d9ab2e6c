// mt_pipe_arz64.cu -- explicit instantiations of the ARZ float64 step kernel
// (see mt_pipe_arz64.inc for why they have a translation unit of their own).
#include "mt_pipe.cuh"

namespace mt {
#define MT_PIPE_ARZ64(K, UNI, U, R) \
  template __global__ void step_pipe_kernel<double, 1, false, K, UNI, U, R>(const PipeArgs<double>);
#include "mt_pipe_arz64.inc"
#undef MT_PIPE_ARZ64
}  // namespace mt
