// Tensor-core detector path, main translation unit: weight packing, the single-rounded (precision 1) kernel instantiations and the
// entry points.  The kernels live in detector_tc.cuh; detector_tc_x3.cu holds the split-precision (precision 2) instantiations.
#define BALF_TC_MAIN
#include "detector_tc.cuh"
