"""Autograd wrappers -- mirror of mmdet3d/ops/spconv/functional.py:22-123."""
from torch.autograd import Function

from . import ops


class SparseConvFunction(Function):
    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        ctx.save_for_backward(features, filters)
        ctx.rulebook = (indice_pairs, indice_pair_num)
        return ops.indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out,
                               False)

    @staticmethod
    def backward(ctx, grad_output):
        features, filters = ctx.saved_tensors
        indice_pairs, indice_pair_num = ctx.rulebook
        input_bp, filters_bp = ops.indice_conv_backward(features, filters, grad_output.contiguous(),
                                                        indice_pairs, indice_pair_num, False)
        return input_bp, filters_bp, None, None, None


class SubMConvFunction(Function):
    @staticmethod
    def forward(ctx, features, filters, indice_pairs, indice_pair_num, num_activate_out):
        ctx.save_for_backward(features, filters)
        ctx.rulebook = (indice_pairs, indice_pair_num)
        return ops.indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out,
                               False, True)

    @staticmethod
    def backward(ctx, grad_output):
        features, filters = ctx.saved_tensors
        indice_pairs, indice_pair_num = ctx.rulebook
        input_bp, filters_bp = ops.indice_conv_backward(features, filters, grad_output.contiguous(),
                                                        indice_pairs, indice_pair_num, False, True)
        return input_bp, filters_bp, None, None, None


def indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out):
    return SparseConvFunction.apply(features, filters, indice_pairs, indice_pair_num,
                                    num_activate_out)


def indice_subm_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out):
    return SubMConvFunction.apply(features, filters, indice_pairs, indice_pair_num,
                                  num_activate_out)
