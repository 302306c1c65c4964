// ORACLE (test infrastructure): pybind shim exposing the REFERENCE's own CPU BEV-IoU
// (/root/reference/efg/operators/src/iou3d_nms/iou3d_cpu.cpp, compiled where it lies by oracle/build_ref.py).
// Only the declaration of iou3d_cpu.h:12 is repeated here.
#include <torch/extension.h>

namespace efg {
int boxes_iou_bev_cpu(at::Tensor boxes_a_tensor, at::Tensor boxes_b_tensor, at::Tensor ans_iou_tensor);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) { m.def("boxes_iou_bev_cpu", &efg::boxes_iou_bev_cpu); }
