"""voxelize -- host-side mirror of mmdet3d/ops/voxel (voxelize.py:1-148, voxelization.h:58-95).

    voxel_layer.hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size,
                              coors_range, max_points, max_voxels, NDim=3, deterministic=True) -> int
    voxel_layer.dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3)
    voxelization(points, voxel_size, coors_range, max_points, max_voxels, deterministic)
    Voxelization(voxel_size, point_cloud_range, max_num_points, max_voxels, deterministic)

GPU only (north star: no CPU fallback): CPU tensors raise."""
import ctypes

import torch
from torch import nn
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from . import _C

__all__ = ["voxel_layer", "voxelization", "Voxelization", "voxelize_mean"]


def _floats(vals, n):
    vals = [float(v) for v in vals]
    if len(vals) != n:
        raise ValueError("expected %d values, got %d" % (n, len(vals)))
    return _C.host_array(ctypes.c_float, vals)


class _VoxelLayer:
    """Stand-in for the reference pybind module `voxel_layer` (voxelization.cpp:7-12)."""

    @staticmethod
    def hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range,
                      max_points, max_voxels, NDim=3, deterministic=True):
        # `deterministic` is accepted for API parity; this implementation is always
        # deterministic (and does not need the O(N^2) scan that made the flag necessary).
        if NDim != 3:
            raise NotImplementedError("only NDim == 3 is supported")
        _C.require_cuda(points, "points", torch.float32)
        _C.require_cuda(voxels, "voxels", torch.float32)
        _C.require_cuda(coors, "coors", torch.int32)
        _C.require_cuda(num_points_per_voxel, "num_points_per_voxel", torch.int32)
        n, f = points.shape
        if voxels.shape[0] < max_voxels or voxels.shape[1] != max_points or voxels.shape[2] != f:
            raise ValueError("voxels must be [>=max_voxels, max_points, num_features]")
        vs, cr = _floats(voxel_size, 3), _floats(coors_range, 6)
        with torch.cuda.device(points.device):
            voxel_num = torch.zeros(1, dtype=torch.int32, device=points.device)
            nbytes = _C.lib().bevb200_hard_voxelize_workspace_bytes(n, int(max_points))
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=points.device)
            rc = _C.lib().bevb200_hard_voxelize(
                _C.ptr(points), n, f, ctypes.cast(vs, ctypes.c_void_p),
                ctypes.cast(cr, ctypes.c_void_p), int(max_points), int(max_voxels),
                _C.ptr(voxels), _C.ptr(coors), _C.ptr(num_points_per_voxel), _C.ptr(voxel_num),
                _C.ptr(ws), ws.numel(), _C.current_stream(points.device))
        _C.check(rc, "hard_voxelize")
        return int(voxel_num.item())  # host int, like voxelization_cuda.cu:369-372

    @staticmethod
    def dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3):
        if NDim != 3:
            raise NotImplementedError("only NDim == 3 is supported")
        _C.require_cuda(points, "points", torch.float32)
        _C.require_cuda(coors, "coors", torch.int32)
        n, f = points.shape
        vs, cr = _floats(voxel_size, 3), _floats(coors_range, 6)
        with torch.cuda.device(points.device):
            rc = _C.lib().bevb200_dynamic_voxelize(
                _C.ptr(points), n, f, ctypes.cast(vs, ctypes.c_void_p),
                ctypes.cast(cr, ctypes.c_void_p), _C.ptr(coors), _C.current_stream(points.device))
        _C.check(rc, "dynamic_voxelize")

    @staticmethod
    def dynamic_point_to_voxel_forward(*args, **kwargs):
        raise NotImplementedError("DynamicScatter is outside the hot path (SURVEY.md section 8f)")

    @staticmethod
    def dynamic_point_to_voxel_backward(*args, **kwargs):
        raise NotImplementedError("DynamicScatter is outside the hot path (SURVEY.md section 8f)")


voxel_layer = _VoxelLayer()
hard_voxelize = voxel_layer.hard_voxelize
dynamic_voxelize = voxel_layer.dynamic_voxelize


class _Voxelization(Function):
    """Same contract as the reference Function (voxelize.py:10-71)."""

    @staticmethod
    def forward(ctx, points, voxel_size, coors_range, max_points=35, max_voxels=20000,
                deterministic=True):
        points = points.contiguous()
        if max_points == -1 or max_voxels == -1:
            coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
            dynamic_voxelize(points, coors, voxel_size, coors_range, 3)
            return coors
        voxels = points.new_zeros(size=(max_voxels, max_points, points.size(1)))
        coors = points.new_zeros(size=(max_voxels, 3), dtype=torch.int)
        num_points_per_voxel = points.new_zeros(size=(max_voxels,), dtype=torch.int)
        voxel_num = hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size,
                                  coors_range, max_points, max_voxels, 3, deterministic)
        return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


voxelization = _Voxelization.apply


class Voxelization(nn.Module):
    """Same constructor and forward as mmdet3d.ops.Voxelization (voxelize.py:77-148)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000,
                 deterministic=True):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else _pair(max_voxels)
        self.deterministic = deterministic
        pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
        vs = torch.tensor(voxel_size, dtype=torch.float32)
        grid_size = torch.round((pcr[3:] - pcr[:3]) / vs).long()
        self.grid_size = grid_size
        self.pcd_shape = [*grid_size[:2], 1]

    def forward(self, input):
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        return voxelization(input, self.voxel_size, self.point_cloud_range, self.max_num_points,
                            max_voxels, self.deterministic)

    def __repr__(self):
        return (self.__class__.__name__ + "(voxel_size=" + str(self.voxel_size)
                + ", point_cloud_range=" + str(self.point_cloud_range) + ", max_num_points="
                + str(self.max_num_points) + ", max_voxels=" + str(self.max_voxels)
                + ", deterministic=" + str(self.deterministic) + ")")


def voxelize_mean(voxels, coors, num_points, batch_idx=0):
    """Fused glue of BEVFusion.voxelize (bevfusion.py:183,191-195): mean over the points of a
    voxel and (batch, x, y, z) coords.  Returns (feats [M, F] fp32, coords [M, 4] int32)."""
    _C.require_cuda(voxels, "voxels", torch.float32)
    _C.require_cuda(coors, "coors", torch.int32)
    _C.require_cuda(num_points, "num_points", torch.int32)
    m, p, f = voxels.shape
    with torch.cuda.device(voxels.device):
        feats = torch.empty((m, f), dtype=torch.float32, device=voxels.device)
        coords4 = torch.empty((m, 4), dtype=torch.int32, device=voxels.device)
        rc = _C.lib().bevb200_voxel_mean(_C.ptr(voxels), _C.ptr(coors), _C.ptr(num_points), m, p, f,
                                         int(batch_idx), _C.ptr(feats), _C.ptr(coords4),
                                         _C.current_stream(voxels.device))
    _C.check(rc, "voxel_mean")
    return feats, coords4


@torch.no_grad()
def voxelize_batch(points, voxelize_module, voxelize_reduce=True):
    """BEVFusion.voxelize (mmdet3d/models/fusion_models/bevfusion.py:169-197): per-sample hard
    voxelization, batch index prepended to the coords (F.pad(c, (1, 0), value=k)), concatenation,
    and -- with voxelize_reduce -- the mean over the points of each voxel.

    points: list of [N_k, F] CUDA tensors.  Returns (feats, coords [M, 4] int32 (b, x, y, z), sizes)."""
    feats, coords, sizes = [], [], []
    for k, res in enumerate(points):
        ret = voxelize_module(res)
        if len(ret) == 3:
            f, c, n = ret
            if voxelize_reduce:
                f, c4 = voxelize_mean(f.contiguous(), c.contiguous(), n.contiguous(), k)
            else:
                c4 = torch.nn.functional.pad(c, (1, 0), mode="constant", value=k)
            sizes.append(n)
        else:                      # dynamic voxelization: coords only
            f, c = res, ret
            c4 = torch.nn.functional.pad(c, (1, 0), mode="constant", value=k)
        feats.append(f)
        coords.append(c4)
    feats = torch.cat(feats, dim=0)
    coords = torch.cat(coords, dim=0)
    sizes = torch.cat(sizes, dim=0) if sizes else sizes
    return feats, coords, sizes
