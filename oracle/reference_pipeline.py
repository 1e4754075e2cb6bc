"""TEST / BENCH INFRASTRUCTURE -- the reference's hot path executed on the CPU (or, for the
encoder, on whatever device the tensors live on) with the reference's own extension modules
(oracle/_ref) where they run, and restatements where they cannot:

  * bev_pool: the reference has no C++ CPU kernel (bev_pool_cpu.cpp is the CUDA launcher);
    its only CPU-capable path is the pure-torch QuickCumsum (bev_pool.py:9-35), restated here
    together with the index glue of base.py:149-169 / bev_pool.py:87-94 and a dense scatter.
  * hard_voxelize: the reference CPU kernel indexes out of bounds on the non-cubic 1440x1440x40
    grid (voxelization_cpu.cpp:75 vs :129-130), so the C port in oracle.c stands in.
  * SparseEncoder: reference rulebook + indice_conv ops from sparse_conv_ext_ref (CPU or CUDA
    tensors), composed exactly as sparse_encoder.py:99-132 / sparse_block.py:94-110 do.

Imported only by tests/ and bench.py (cpu_baseline / --impl reference)."""
import numpy as np
import torch

from . import conv_output_size, hard_voxelize


def bev_pool_cpu_quickcumsum(x, geom, dx, bx, nx):
    """BaseTransform.bev_pool (base.py:141-176) + bev_pool() (bev_pool.py:84-98) with the
    QuickCumsum CPU path (bev_pool.py:9-35) and a dense scatter into [B, C*nz, nx, ny]."""
    B, N, D, H, W, C = x.shape
    Nprime = B * N * D * H * W
    x = x.reshape(Nprime, C)
    geom_feats = ((geom - (bx - dx / 2.0)) / dx).long().view(Nprime, 3)
    batch_ix = torch.cat([torch.full([Nprime // B, 1], ix, dtype=torch.long) for ix in range(B)])
    geom_feats = torch.cat((geom_feats, batch_ix), 1)
    kept = ((geom_feats[:, 0] >= 0) & (geom_feats[:, 0] < nx[0]) & (geom_feats[:, 1] >= 0)
            & (geom_feats[:, 1] < nx[1]) & (geom_feats[:, 2] >= 0) & (geom_feats[:, 2] < nx[2]))
    x, geom_feats = x[kept], geom_feats[kept]
    nz, nxx, nyy = int(nx[2]), int(nx[0]), int(nx[1])
    ranks = (geom_feats[:, 0] * (nyy * nz * B) + geom_feats[:, 1] * (nz * B)
             + geom_feats[:, 2] * B + geom_feats[:, 3])
    indices = ranks.argsort()
    x, geom_feats, ranks = x[indices], geom_feats[indices], ranks[indices]
    # QuickCumsum.forward
    x = x.cumsum(0)
    k = torch.ones(x.shape[0], dtype=torch.bool)
    k[:-1] = ranks[1:] != ranks[:-1]
    x, geom_feats = x[k], geom_feats[k]
    x = torch.cat((x[:1], x[1:] - x[:-1]))
    final = torch.zeros((B, nz, nxx, nyy, C), dtype=x.dtype)
    final[geom_feats[:, 3], geom_feats[:, 2], geom_feats[:, 0], geom_feats[:, 1]] = x
    final = final.permute(0, 4, 1, 2, 3)
    return torch.cat(final.unbind(dim=2), 1)


def voxelize_cpu(points_np, cfg, max_voxels):
    """Voxelization.forward + BEVFusion.voxelize glue (voxelize.py:121-138, bevfusion.py:169-197)."""
    v, c, n, m = hard_voxelize(points_np, cfg["voxel_size"], cfg["point_cloud_range"],
                               cfg["max_num_points"], max_voxels)
    v, c, n = torch.from_numpy(v), torch.from_numpy(c), torch.from_numpy(n)
    coords = torch.nn.functional.pad(c, (1, 0), mode="constant", value=0)
    feats = v.sum(dim=1) / n.type_as(v).view(-1, 1)
    return feats.contiguous(), coords.contiguous()


def reference_encoder_forward(ref, model, feats, coors, batch_size):
    """SparseEncoder.forward (sparse_encoder.py:99-132) executed with the REFERENCE extension's
    rulebook + conv ops and torch BN / ReLU, using `model`'s weights (any device)."""
    from bevfusion_b200.sparse_block import SparseBasicBlock

    def conv(module, f, idx, shape):
        subm = module.subm
        ks, st, pd, dl = module.kernel_size, module.stride, module.padding, module.dilation
        out_shape = list(shape) if subm else conv_output_size(shape, ks, st, pd, dl)
        outids, pairs, num = ref.get_indice_pairs_3d(idx, batch_size, out_shape, list(shape), ks, st,
                                                     pd, dl, [0, 0, 0], int(subm), 0)
        out = ref.indice_conv_fp32(f, module.weight.detach(), pairs, num, outids.shape[0], 0, int(subm))
        return out, outids, out_shape

    def seq(s, f, idx, shape):
        f, idx, shape = conv(s[0], f, idx, shape)
        return torch.relu(s[1](f)), idx, shape

    shape = list(model.sparse_shape)
    f, idx, shape = seq(model.conv_input, feats, coors, shape)
    for stage in model.encoder_layers:
        for block in stage:
            if isinstance(block, SparseBasicBlock):
                identity = f
                o, _, _ = conv(block.conv1, f, idx, shape)
                o = torch.relu(block.norm1(o))
                o, _, _ = conv(block.conv2, o, idx, shape)
                f = torch.relu(block.norm2(o) + identity)
            else:
                f, idx, shape = seq(block, f, idx, shape)
    f, idx, shape = seq(model.conv_out, f, idx, shape)
    dense = torch.zeros(batch_size, *shape, f.shape[1], device=f.device)
    li = idx.long()
    dense[li[:, 0], li[:, 1], li[:, 2], li[:, 3]] = f
    dense = dense.permute(0, 4, 1, 2, 3).contiguous()
    N, C, H, W, D = dense.shape
    return dense.permute(0, 1, 4, 2, 3).contiguous().view(N, C * D, H, W)
