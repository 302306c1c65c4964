// ORACLE (test infrastructure): pybind shim that exposes the REFERENCE's own box-attention CUDA kernels
// (/root/reference/efg/operators/src/box_attn/box_attn.cu + box_attn_kernel.cuh), compiled where they lie for sm_100a
// by oracle/build_ref.py.  It serves as the GPU comparator ("the bar to beat on B200") and as
// an extra parity check; nothing of the reference is copied here: only the two declarations of box_attn.h:8-26.
#include <torch/extension.h>

namespace efg {
at::Tensor box_attn_cuda_forward(const at::Tensor& value, const at::Tensor& spatial_shapes,
                                 const at::Tensor& level_start_index, const at::Tensor& sampling_loc,
                                 const at::Tensor& attn_weight, const int im2col_step);
std::vector<at::Tensor> box_attn_cuda_backward(const at::Tensor& value, const at::Tensor& spatial_shapes,
                                               const at::Tensor& level_start_index, const at::Tensor& sampling_loc,
                                               const at::Tensor& attn_weight, const at::Tensor& grad_output,
                                               const int im2col_step);
}  // namespace efg

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("box_attn_forward", &efg::box_attn_cuda_forward);
  m.def("box_attn_backward", &efg::box_attn_cuda_backward);
}
