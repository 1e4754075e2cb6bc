"""Autograd wrappers -- mirror of mmdet3d/ops/spconv/functional.py:22-123 (forward only for
now; backward raises until indice_conv_backward lands)."""
from torch.autograd import Function

from . import ops


class SparseConvFunction(Function):
    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        ctx.save_for_backward(features, filters)
        return ops.indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out,
                               False)

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError("sparse conv backward is not built yet")


class SubMConvFunction(Function):
    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        ctx.save_for_backward(features, filters)
        return ops.indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out,
                               False, True)

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError("sparse conv backward is not built yet")


def indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out):
    return SparseConvFunction.apply(features, filters, indice_pairs, indice_pair_num,
                                    num_activate_out)


def indice_subm_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out):
    return SubMConvFunction.apply(features, filters, indice_pairs, indice_pair_num,
                                  num_activate_out)
