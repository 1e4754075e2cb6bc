"""Containers for sparse modules (the role of mmdet3d/ops/spconv/modules.py:77-139): a
`SparseSequential` is a plain `nn.Sequential` whose dense members (BatchNorm1d, ReLU, ...) are applied to
the `.features` of the SparseConvTensor that flows through it.  Same sub-module naming as
`nn.Sequential` ("0", "1", ... or the given names), so reference state_dicts load unchanged."""
from collections import OrderedDict

from torch import nn

from .structure import SparseConvTensor


class SparseModule(nn.Module):
    """Marker base: a module whose forward takes a SparseConvTensor."""


class SparseSequential(nn.Sequential, SparseModule):
    def __init__(self, *modules, **named):
        if len(modules) == 1 and isinstance(modules[0], OrderedDict):
            super().__init__(modules[0])
        else:
            super().__init__(*modules)
        for name, module in named.items():
            if name in self._modules:
                raise ValueError(f"duplicate sub-module name {name!r}")
            self.add_module(name, module)

    def add(self, module, name=None):
        self.add_module(str(len(self)) if name is None else name, module)

    def forward(self, x):
        for module in self:
            if isinstance(module, SparseModule):
                x = module(x)
            elif isinstance(x, SparseConvTensor):
                if x.features.shape[0]:        # dense layers see the [N, C] feature matrix
                    x.features = module(x.features)
            else:
                x = module(x)
        return x


class ToDense(SparseModule):
    def forward(self, x):
        return x.dense()


class RemoveGrid(SparseModule):
    def forward(self, x):
        x.grid = None
        return x
