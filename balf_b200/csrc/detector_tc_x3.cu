// Tensor-core detector path, split precision ("f16x3", precision 2): instantiations of the kernels of detector_tc.cuh with PX = 1
// (a translation unit of its own so that the two precision classes compile in parallel).
#include "detector_tc.cuh"

namespace balf {
template int tc_level_px<1>(int, const float*, const DownW&, const balf_detector_arch&, const float*, int, int, int, float*, float*,
                            float*, float*, float*, cudaStream_t, int);
template int tc_head_px<1>(const float*, const float*, const float*, const balf_detector_arch&, const float*, int, int, int, float*,
                           float*, cudaStream_t, int);
}  // namespace balf
