// Drop-in for the reference's pybind module `bev_pool_ext` (mmdet3d/ops/bev_pool/src/bev_pool_cpu.cpp:89-94):
// same function names, argument order (interval_LENGTHS before interval_STARTS, :22-25) and return
// shapes, each a thin shim over one C-ABI call of libbevfusion_b200 (include/bevfusion_b200.h).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "bevfusion_b200.h"

static void check_cuda(const at::Tensor &t, const char *name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor: bevfusion_b200 has no CPU path");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}

at::Tensor bev_pool_forward(const at::Tensor x, const at::Tensor geom_feats, const at::Tensor interval_lengths,
                            const at::Tensor interval_starts, int b, int d, int h, int w) {
  check_cuda(x, "x"); check_cuda(geom_feats, "geom_feats");
  check_cuda(interval_lengths, "interval_lengths"); check_cuda(interval_starts, "interval_starts");
  const c10::cuda::OptionalCUDAGuard guard(device_of(x));
  const int n = x.size(0), c = x.size(1), n_int = interval_lengths.size(0);
  auto out = torch::empty({b, d, h, w, c}, x.options());          // the library zero-fills
  auto ws = torch::empty({(int64_t)bevb200_bev_pool_workspace_bytes(n, c) + 256}, x.options().dtype(torch::kUInt8));
  const int rc = bevb200_bev_pool(b, d, h, w, n, c, n_int, x.data_ptr<float>(), geom_feats.data_ptr<int>(),
                                  interval_starts.data_ptr<int>(), interval_lengths.data_ptr<int>(),
                                  out.data_ptr<float>(), ws.data_ptr(), (size_t)ws.numel(),
                                  at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, bevb200_last_error());
  return out;
}

at::Tensor bev_pool_backward(const at::Tensor out_grad, const at::Tensor geom_feats, const at::Tensor interval_lengths,
                             const at::Tensor interval_starts, int b, int d, int h, int w) {
  check_cuda(out_grad, "out_grad"); check_cuda(geom_feats, "geom_feats");
  const c10::cuda::OptionalCUDAGuard guard(device_of(out_grad));
  const int n = geom_feats.size(0), c = out_grad.size(4);
  auto x_grad = torch::empty({n, c}, out_grad.options());
  const int rc = bevb200_bev_pool_grad(b, d, h, w, n, c, (int)interval_lengths.size(0), out_grad.data_ptr<float>(),
                                       geom_feats.data_ptr<int>(), interval_starts.data_ptr<int>(),
                                       interval_lengths.data_ptr<int>(), x_grad.data_ptr<float>(),
                                       at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == 0, bevb200_last_error());
  return x_grad;
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("bev_pool_forward", &bev_pool_forward, "bev_pool_forward");
  m.def("bev_pool_backward", &bev_pool_backward, "bev_pool_backward");
}
