"""bevfusion_b200 -- Blackwell (sm_100a) implementation of the BEVFusion view-transform /
LiDAR-voxel hot path: bev_pool, hard voxelization, sparse 3D convolution.

Everything that computes runs in libbevfusion_b200.so (hand-written CUDA behind the C ABI of
include/bevfusion_b200.h).  The Python modules here mirror the reference's op interface
(mmdet3d/ops/{bev_pool,voxel,spconv}) and are thin: argument checks, tensor allocation,
ctypes calls on the current CUDA stream.  There is no CPU fallback: on a machine without the
library or without a GPU the ops raise.
"""
from . import _C  # noqa: F401

__version__ = "0.1.0"
