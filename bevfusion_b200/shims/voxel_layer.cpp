// Drop-in for the reference's pybind module `voxel_layer` (mmdet3d/ops/voxel/src/voxelization.cpp:7-13,
// declarations voxelization.h:58-140): same names, argument order, defaults and return values, each a thin
// shim over one C-ABI call of libbevfusion_b200.  There is no CPU path: CPU tensors raise.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <string>
#include <vector>

#include "bevfusion_b200.h"

namespace {
void check_cuda(const at::Tensor &t, const char *name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor: bevfusion_b200 has no CPU path");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}
cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
int reduce_code(const std::string &r) {      // scatter_points_cuda.cu:7 / voxelization.h:97-106
  if (r == "sum") return 0;
  if (r == "mean") return 1;
  if (r == "max") return 2;
  TORCH_CHECK(false, "do not support reduce type ", r);
  return -1;
}
}  // namespace

int hard_voxelize(const at::Tensor &points, at::Tensor &voxels, at::Tensor &coors, at::Tensor &num_points_per_voxel,
                  const std::vector<float> voxel_size, const std::vector<float> coors_range, const int max_points,
                  const int max_voxels, const int NDim = 3, const bool deterministic = true) {
  check_cuda(points, "points"); check_cuda(voxels, "voxels"); check_cuda(coors, "coors");
  check_cuda(num_points_per_voxel, "num_points_per_voxel");
  TORCH_CHECK(NDim == 3 && voxel_size.size() == 3 && coors_range.size() == 6, "3-D voxelization only");
  (void)deterministic;                         // the library is always deterministic
  c10::cuda::CUDAGuard guard(points.device());
  auto voxel_num = torch::zeros({1}, coors.options());
  auto ws = torch::empty({(int64_t)bevb200_hard_voxelize_workspace_bytes(points.size(0), max_points) + 256},
                         points.options().dtype(torch::kUInt8));
  const int rc = bevb200_hard_voxelize(points.data_ptr<float>(), points.size(0), points.size(1), voxel_size.data(),
                                       coors_range.data(), max_points, max_voxels, voxels.data_ptr<float>(),
                                       coors.data_ptr<int>(), num_points_per_voxel.data_ptr<int>(),
                                       voxel_num.data_ptr<int>(), ws.data_ptr(), (size_t)ws.numel(), cur_stream());
  TORCH_CHECK(rc == 0, bevb200_last_error());
  return voxel_num.item<int>();                // the one D2H read the reference also does (voxelization_cuda.cu:369-370)
}

void dynamic_voxelize(const at::Tensor &points, at::Tensor &coors, const std::vector<float> voxel_size,
                      const std::vector<float> coors_range, const int NDim = 3) {
  check_cuda(points, "points"); check_cuda(coors, "coors");
  TORCH_CHECK(NDim == 3 && voxel_size.size() == 3 && coors_range.size() == 6, "3-D voxelization only");
  c10::cuda::CUDAGuard guard(points.device());
  const int rc = bevb200_dynamic_voxelize(points.data_ptr<float>(), points.size(0), points.size(1), voxel_size.data(),
                                          coors_range.data(), coors.data_ptr<int>(), cur_stream());
  TORCH_CHECK(rc == 0, bevb200_last_error());
}

std::vector<at::Tensor> dynamic_point_to_voxel_forward(const at::Tensor &feats, const at::Tensor &coors,
                                                       const std::string &reduce_type) {
  check_cuda(feats, "feats"); check_cuda(coors, "coors");
  c10::cuda::CUDAGuard guard(feats.device());
  const int n = feats.size(0), c = feats.size(1), ndim = coors.size(1), red = reduce_code(reduce_type);
  auto reduced = at::empty({n, c}, feats.options());
  auto out_coors = at::empty({n, ndim}, coors.options());
  auto map = at::empty({n}, coors.options()), count = at::empty({n}, coors.options());
  auto meta = at::zeros({2}, coors.options());
  auto ws = at::empty({(int64_t)bevb200_dynamic_scatter_workspace_bytes(n) + 256}, feats.options().dtype(at::kByte));
  const int rc = bevb200_dynamic_scatter(feats.data_ptr<float>(), coors.data_ptr<int>(), n, c, ndim, red,
                                         reduced.data_ptr<float>(), out_coors.data_ptr<int>(), map.data_ptr<int>(),
                                         count.data_ptr<int>(), nullptr, meta.data_ptr<int>(), ws.data_ptr(),
                                         (size_t)ws.numel(), cur_stream());
  TORCH_CHECK(rc == 0, bevb200_last_error());
  auto meta_h = meta.cpu();
  TORCH_CHECK(meta_h[1].item<int>() == 0, "dynamic_point_to_voxel_forward: coordinate too large for the sort key");
  const int m = meta_h[0].item<int>();
  return {reduced.slice(0, 0, m), out_coors.slice(0, 0, m), map, count.slice(0, 0, m)};
}

void dynamic_point_to_voxel_backward(at::Tensor &grad_feats, const at::Tensor &grad_reduced_feats, const at::Tensor &feats,
                                     const at::Tensor &reduced_feats, const at::Tensor &coors_idx,
                                     const at::Tensor &reduce_count, const std::string &reduce_type) {
  check_cuda(grad_feats, "grad_feats"); check_cuda(grad_reduced_feats, "grad_reduced_feats"); check_cuda(feats, "feats");
  check_cuda(reduced_feats, "reduced_feats"); check_cuda(coors_idx, "coors_idx"); check_cuda(reduce_count, "reduce_count");
  c10::cuda::CUDAGuard guard(feats.device());
  const int n = feats.size(0), c = feats.size(1), m = reduced_feats.size(0), red = reduce_code(reduce_type);
  auto from = at::empty({m > 0 ? m : 1, c}, coors_idx.options());
  const int rc = bevb200_dynamic_scatter_backward(grad_reduced_feats.data_ptr<float>(), feats.data_ptr<float>(),
                                                  reduced_feats.data_ptr<float>(), coors_idx.data_ptr<int>(),
                                                  reduce_count.data_ptr<int>(), from.data_ptr<int>(), 0, n, m, c, red,
                                                  grad_feats.data_ptr<float>(), cur_stream());
  TORCH_CHECK(rc == 0, bevb200_last_error());
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("hard_voxelize", &hard_voxelize, "hard voxelize", py::arg("points"), py::arg("voxels"), py::arg("coors"),
        py::arg("num_points_per_voxel"), py::arg("voxel_size"), py::arg("coors_range"), py::arg("max_points"),
        py::arg("max_voxels"), py::arg("NDim") = 3, py::arg("deterministic") = true);
  m.def("dynamic_voxelize", &dynamic_voxelize, "dynamic voxelization", py::arg("points"), py::arg("coors"),
        py::arg("voxel_size"), py::arg("coors_range"), py::arg("NDim") = 3);
  m.def("dynamic_point_to_voxel_forward", &dynamic_point_to_voxel_forward, "dynamic point to voxel forward");
  m.def("dynamic_point_to_voxel_backward", &dynamic_point_to_voxel_backward, "dynamic point to voxel backward");
}
