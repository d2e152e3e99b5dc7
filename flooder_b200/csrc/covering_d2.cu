// cover_eval_kernel for ambient dimension D = 2 (see covering_kernels.cuh)
#include "covering_kernels.cuh"

namespace flood {
template int dispatch_eval<2>(CoverParams &, int64_t, cudaStream_t);
}  // namespace flood
