"""The compiled drop-in pybind modules (bevfusion_b200/shims/*.cpp -> bev_pool_ext, voxel_layer,
sparse_conv_ext) called exactly as the reference's python wrappers call theirs
(mmdet3d/ops/bev_pool/bev_pool.py:41-81, ops/voxel/voxelize.py:43-70, ops/spconv/ops.py:45-189), checked
against the oracle.  INTEGRATION.md section B describes these bindings; here they are built and run."""
import numpy as np
import pytest
import torch

import oracle
from bevfusion_b200.shims import build as shim_build

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shims():
    return {n: shim_build.load_module(n) for n in shim_build.MODULES}


def test_bev_pool_ext_like_quickcumsumcuda(cuda, shims):
    """QuickCumsumCuda.forward / backward (bev_pool.py:38-81) on top of the compiled bev_pool_ext."""
    ext = shims["bev_pool_ext"]
    rng = np.random.default_rng(0)
    B, D, H, W, C, n = 2, 1, 40, 36, 80, 60000
    coords = np.stack([rng.integers(0, H, n), rng.integers(0, W, n), rng.integers(0, D, n), rng.integers(0, B, n)], 1)
    feats = rng.standard_normal((n, C)).astype(np.float32)
    ranks = coords[:, 0] * (W * D * B) + coords[:, 1] * (D * B) + coords[:, 2] * B + coords[:, 3]      # bev_pool.py:87-92
    order = np.argsort(ranks, kind="stable")
    x = torch.from_numpy(feats[order]).to(cuda)
    geom = torch.from_numpy(coords[order].astype(np.int32)).to(cuda)
    r = torch.from_numpy(ranks[order]).to(cuda)
    kept = torch.ones(n, device=cuda, dtype=torch.bool)                                                 # bev_pool.py:41-46
    kept[1:] = r[1:] != r[:-1]
    interval_starts = torch.where(kept)[0].int()
    interval_lengths = torch.zeros_like(interval_starts)
    interval_lengths[:-1] = interval_starts[1:] - interval_starts[:-1]
    interval_lengths[-1] = n - interval_starts[-1]
    out = ext.bev_pool_forward(x, geom, interval_lengths, interval_starts, B, D, H, W)                   # bev_pool.py:49-58
    assert tuple(out.shape) == (B, D, H, W, C)
    gold = oracle.bev_pool(feats[order], coords[order], B, D, H, W).transpose(0, 2, 3, 4, 1)             # [B, D, H, W, C]
    assert np.abs(out.cpu().numpy() - gold).max() <= 1e-4 * np.abs(gold).max()
    og = torch.from_numpy(rng.standard_normal((B, D, H, W, C)).astype(np.float32)).to(cuda)
    xg = ext.bev_pool_backward(og, geom, interval_lengths, interval_starts, B, D, H, W)                  # bev_pool.py:70-79
    g = geom.long()
    assert bool(torch.equal(xg, og[g[:, 3], g[:, 2], g[:, 0], g[:, 1]]))
    with pytest.raises(RuntimeError):
        ext.bev_pool_forward(x.cpu(), geom, interval_lengths, interval_starts, B, D, H, W)


def test_voxel_layer_like_voxelization(cuda, shims):
    """_Voxelization.forward (voxelize.py:43-70): zero-filled cap-size outputs, slice by the returned count."""
    vl = shims["voxel_layer"]
    vs, cr, max_points, max_voxels = [0.25, 0.25, 0.5], [-8.0, -8.0, -2.0, 8.0, 8.0, 2.0], 5, 3000
    from bevfusion_b200 import synthetic as S
    pts_np = S.uniform_cloud(20000, seed=3, margin=1.0, rng_range=cr)
    points = torch.from_numpy(pts_np).to(cuda)
    voxels = points.new_zeros(size=(max_voxels, max_points, points.size(1)))
    coors = points.new_zeros(size=(max_voxels, 3), dtype=torch.int)
    num = points.new_zeros(size=(max_voxels,), dtype=torch.int)
    voxel_num = vl.hard_voxelize(points, voxels, coors, num, vs, cr, max_points, max_voxels, 3, True)
    gv, gc, gn, gm = oracle.hard_voxelize(pts_np, vs, cr, max_points, max_voxels)
    assert voxel_num == gm
    assert np.array_equal(coors[:voxel_num].cpu().numpy(), gc) and np.array_equal(num[:voxel_num].cpu().numpy(), gn)
    assert np.array_equal(voxels[:voxel_num].cpu().numpy(), gv)
    assert not bool(voxels[voxel_num:].any())                                  # untouched rows stay zero
    dcoors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
    vl.dynamic_voxelize(points, dcoors, vs, cr, 3)
    assert np.array_equal(dcoors.cpu().numpy(), oracle.dynamic_voxelize(pts_np, vs, cr))
    ok = (dcoors >= 0).all(1)
    red, oc, cmap, cnt = vl.dynamic_point_to_voxel_forward(points[ok], dcoors[ok], "mean")
    gred, goc, gmap, gcnt = oracle.dynamic_scatter(pts_np[ok.cpu().numpy()], dcoors[ok].cpu().numpy(), "mean")
    assert np.array_equal(oc.cpu().numpy(), goc) and np.array_equal(cnt.cpu().numpy(), gcnt)
    assert np.abs(red.cpu().numpy() - gred).max() <= 1e-5 * max(np.abs(gred).max(), 1.0)
    with pytest.raises(RuntimeError):
        vl.hard_voxelize(points.cpu(), voxels, coors, num, vs, cr, max_points, max_voxels, 3, True)


def test_sparse_conv_ext_like_ops(cuda, shims):
    """ops.get_indice_pairs + indice_conv + indice_conv_backward (spconv/ops.py:45-189) on the compiled module."""
    sp = shims["sparse_conv_ext"]
    rng = np.random.default_rng(4)
    shape, B, n, cin, cout = [30, 28, 9], 2, 3000, 16, 32
    vol = B * shape[0] * shape[1] * shape[2]
    flat = rng.choice(vol, size=n, replace=False)
    idx = np.stack([flat // (shape[0] * shape[1] * shape[2]), (flat // (shape[1] * shape[2])) % shape[0],
                    (flat // shape[2]) % shape[1], flat % shape[2]], 1).astype(np.int32)
    feat = rng.standard_normal((n, cin)).astype(np.float32)
    W = (rng.standard_normal((3, 3, 3, cin, cout)) / np.sqrt(cin * 9)).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    for subm, stride, pad in ((1, [1, 1, 1], [1, 1, 1]), (0, [2, 2, 2], [1, 1, 1])):
        out_shape = shape if subm else [(s + 2 * p - 3) // st + 1 for s, p, st in zip(shape, pad, stride)]
        outids, pairs, num = sp.get_indice_pairs_3d(t(idx), B, out_shape, shape, [3, 3, 3], stride, pad, [1, 1, 1],
                                                    [0, 0, 0], subm, 0)
        gold, gids, _ = oracle.sparse_conv(feat, idx, B, shape, W, [3] * 3, stride, pad, [1] * 3, bool(subm), acc64=True)
        assert np.array_equal(outids.cpu().numpy(), gids)
        out = sp.indice_conv_fp32(t(feat), t(W), pairs, num, outids.shape[0], 0, subm)
        assert np.abs(out.cpu().numpy() - gold).max() <= 1e-4 * np.abs(gold).max()
        bias = rng.standard_normal(cout).astype(np.float32)
        outb = sp.fused_indice_conv_fp32(t(feat), t(W), t(bias), pairs, num, outids.shape[0], 0, subm)
        assert np.abs(outb.cpu().numpy() - (gold + bias)).max() <= 1e-4 * np.abs(gold).max()
        outh = sp.indice_conv_half(t(feat).half(), t(W).half(), pairs, num, outids.shape[0], 0, subm)
        assert outh.dtype == torch.half and np.abs(outh.float().cpu().numpy() - gold).max() <= 2e-2 * np.abs(gold).max()
        g = rng.standard_normal((outids.shape[0], cout)).astype(np.float32)
        din, dw = sp.indice_conv_backward_fp32(t(feat), t(W), t(g), pairs, num, 0, subm)
        # gradient check against autograd of the dense formulation is done in test_spconv_gpu; here: the
        # weight gradient of this binding equals the python mirror's (same library call underneath)
        from bevfusion_b200.spconv import ops
        rb, _ = ops.get_rulebook(t(idx), B, shape, 3, stride, pad, 1, 0, bool(subm))
        din2, dw2 = ops.sparse_conv_backward(t(feat), t(W), t(g), rb.nbr)
        assert float((din - din2).abs().max()) <= 1e-5 * float(din2.abs().max())
        assert float((dw - dw2).abs().max()) <= 1e-4 * float(dw2.abs().max())
