// ORACLE (test infrastructure): pybind shim that exposes the REFERENCE's own CPU voxelizer,
// compiled from /root/reference/efg/operators/src/voxelize/voxelization_cpu.cpp where it lies
// (see oracle/build_ref.py).  Nothing of the reference is copied here: only the two declarations
// needed to bind its functions (voxelization.h:12-18).
#include <torch/extension.h>

namespace efg {
int hard_voxelize_cpu(const at::Tensor& points, at::Tensor& voxels, at::Tensor& coors,
                      at::Tensor& num_points_per_voxel, const std::vector<float> voxel_size,
                      const std::vector<float> coors_range, const int max_points, const int max_voxels,
                      const int NDim);
void dynamic_voxelize_cpu(const at::Tensor& points, at::Tensor& coors, const std::vector<float> voxel_size,
                          const std::vector<float> coors_range, const int NDim);
}  // namespace efg

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("hard_voxelize", &efg::hard_voxelize_cpu);
  m.def("dynamic_voxelize", &efg::dynamic_voxelize_cpu);
}
