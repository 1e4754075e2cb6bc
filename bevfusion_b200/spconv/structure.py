"""SparseConvTensor -- mirror of mmdet3d/ops/spconv/structure.py:21-63."""
import numpy as np
import torch

from . import ops


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = spatial_shape
        self.batch_size = batch_size
        self.indice_dict = {}
        self.grid = grid

    @property
    def spatial_size(self):
        return np.prod(self.spatial_shape)

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key, None)

    def dense(self, channels_first=True):
        """[B, C, X, Y, Z] (channels_first) or [B, X, Y, Z, C]; zero outside the active set."""
        out = ops.sparse_to_dense(self.features.contiguous(), self.indices.int().contiguous(),
                                  int(self.batch_size), self.spatial_shape, z_major=False)
        if channels_first:
            return out
        ndim = len(self.spatial_shape)
        return out.permute(0, *range(2, ndim + 2), 1).contiguous()

    @property
    def sparity(self):
        return self.indices.shape[0] / np.prod(self.spatial_shape) / self.batch_size
