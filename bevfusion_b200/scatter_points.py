"""scatter_points -- host-side mirror of mmdet3d/ops/voxel/scatter_points.py:1-104.

    dynamic_scatter(feats, coors, reduce_type="max") -> (voxel_feats, voxel_coors)
    DynamicScatter(voxel_size, point_cloud_range, average_points)(points, coors)

GPU only.  Unlike the reference (float atomics) the sum / mean results are bit-reproducible,
and a batched [N, 4] coors tensor is reduced in ONE launch sequence instead of a python loop
over samples (scatter_points.py:84-95): the batch index is simply the leading key column."""
import torch
from torch import nn
from torch.autograd import Function

from .voxelize import _dynamic_scatter_backward, _dynamic_scatter_forward

__all__ = ["dynamic_scatter", "DynamicScatter"]


class _dynamic_scatter(Function):
    @staticmethod
    def forward(ctx, feats, coors, reduce_type="max"):
        feats = feats.contiguous()
        coors = coors.contiguous()
        voxel_feats, voxel_coors, point2voxel_map, count, reduce_from = _dynamic_scatter_forward(
            feats.detach(), coors, reduce_type)
        ctx.reduce_type = reduce_type
        ctx.n_points = feats.shape[0]
        ctx.has_from = reduce_from is not None
        saved = [point2voxel_map, count] + ([reduce_from] if ctx.has_from else [])
        ctx.save_for_backward(*saved)
        ctx.mark_non_differentiable(voxel_coors)
        return voxel_feats, voxel_coors

    @staticmethod
    def backward(ctx, grad_voxel_feats, grad_voxel_coors=None):
        saved = ctx.saved_tensors
        point2voxel_map, count = saved[0], saved[1]
        reduce_from = saved[2] if ctx.has_from else None
        grad_feats = torch.empty((ctx.n_points, grad_voxel_feats.shape[1]), dtype=torch.float32,
                                 device=grad_voxel_feats.device)
        if ctx.n_points:
            _dynamic_scatter_backward(grad_feats, grad_voxel_feats.contiguous().float(), None, None,
                                      point2voxel_map, count, ctx.reduce_type,
                                      reduce_from if ctx.has_from else None)
        return grad_feats, None, None


dynamic_scatter = _dynamic_scatter.apply


class DynamicScatter(nn.Module):
    """Same constructor / forward as the reference module (scatter_points.py:53-104)."""

    def __init__(self, voxel_size, point_cloud_range, average_points: bool):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.average_points = average_points

    def forward_single(self, points, coors):
        reduce = "mean" if self.average_points else "max"
        return dynamic_scatter(points.contiguous(), coors.contiguous(), reduce)

    def forward(self, points, coors):
        # [N, 3] coors: one sample.  [N, 4] (batch, ...) coors: the reference loops over samples
        # and concatenates; sorting on (batch, ...) keys gives the same rows in the same order.
        return self.forward_single(points, coors)

    def __repr__(self):
        return (self.__class__.__name__ + "(voxel_size=" + str(self.voxel_size)
                + ", point_cloud_range=" + str(self.point_cloud_range) + ", average_points="
                + str(self.average_points) + ")")
