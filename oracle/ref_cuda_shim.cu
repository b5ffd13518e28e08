// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Builds the REFERENCE's own CUDA implementation of the DCNv3 core for sm_100a so that it can be
// (a) a second oracle on the GPU box and (b) the GPU baseline bench.py reports next to our kernels.
// The reference sources are compiled WHERE THEY LIE (/root/reference/models/ops_dcnv3/src/...):
// nothing is copied.  Two things are supplied here because the reference tree lacks them:
//   1. an overload that lets AT_DISPATCH_FLOATING_TYPES_AND_HALF accept `input.type()`
//      (dcnv3_cuda.cu:69,147 pass a DeprecatedTypeProperties; torch >= 2.x only takes a ScalarType —
//      SURVEY §0.3), and
//   2. the pybind entry point: the reference's src/vision.cpp is missing from the tree
//      (.MISSING_LARGE_BLOBS:2); the two exported names are those functions/dcnv3_func.py:39,54 call.
// Output: oracle/_ref/dcnv3_ref_cuda.so (git-ignored, travels to the GPU box).
#include <torch/extension.h>

namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
}  // namespace detail

#include "cuda/dcnv3_cuda.cu"  // resolved through -I /root/reference/models/ops_dcnv3/src

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("dcnv3_forward", &dcnv3_cuda_forward, "reference dcnv3_cuda_forward (dcnv3_cuda.cu:21-85)");
    m.def("dcnv3_backward", &dcnv3_cuda_backward, "reference dcnv3_cuda_backward (dcnv3_cuda.cu:87-173)");
}
