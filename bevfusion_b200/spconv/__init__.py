"""bevfusion_b200.spconv -- mirror of the mmdet3d.ops.spconv package surface used by the
BEVFusion LiDAR branch (SparseConvTensor, SparseConv3d, SubMConv3d, SparseSequential, ...).
2-D / 4-D, transposed, inverse convs and sparse max-pool are outside the hot path."""
from .conv import CONV_LAYERS, SparseConv3d, SparseConvolution, SubMConv3d
from .modules import SparseModule, SparseSequential, ToDense, RemoveGrid
from .structure import SparseConvTensor
from . import ops
from .ops import sparse_conv_ext

__all__ = ["SparseConv3d", "SubMConv3d", "SparseConvolution", "SparseModule", "SparseSequential",
           "SparseConvTensor", "ToDense", "RemoveGrid", "ops", "sparse_conv_ext", "CONV_LAYERS"]
