"""GPU parity for the LiDAR depth-image kernel (BaseDepthTransform.forward, base.py:279-329)."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def run_ours(cuda, clouds, M, image_size, **kw):
    from bevfusion_b200.vtransform import points_to_depth
    pts = [torch.from_numpy(c).to(cuda) for c in clouds]
    keep = [p.clone() for p in pts]
    d = points_to_depth(pts, M["lidar2image"].to(cuda), M["img_aug_matrix"].to(cuda),
                        M["lidar_aug_matrix"].to(cuda), image_size, **kw)
    for p, k in zip(pts, keep):
        assert torch.equal(p, k)                       # the caller's points are not modified
    return d.cpu().numpy()


def oracle_kw(kw):
    return {k: v for k, v in kw.items() if k != "height_expand"}


@pytest.mark.parametrize("kw", [dict(depth_input="scalar"), dict(depth_input="scalar", add_depth_features=True),
                                dict(depth_input="one-hot", depth_bins=20),
                                dict(depth_input="one-hot", depth_bins=20, add_depth_features=True)])
def test_golden_inputs_bit_exact_vs_oracle(cuda, golden_dir, kw):
    g = np.load(os.path.join(golden_dir, "depth_tiny.npz"))
    M = {k: torch.from_numpy(g[k]) for k in ("lidar2image", "img_aug_matrix", "lidar_aug_matrix")}
    clouds = [g["points0"], g["points1"]]
    ours = run_ours(cuda, clouds, M, g["image_size"], **kw)
    for b in range(2):
        gold = oracle.points_to_depth(clouds[b], g["lidar2image"][b], g["img_aug_matrix"][b],
                                      g["lidar_aug_matrix"][b], g["image_size"], **oracle_kw(kw))
        assert np.array_equal(ours[b], gold)
    # and against the tensor the reference source produced (same tolerance as the oracle's pinning test)
    if kw == dict(depth_input="scalar"):
        ref = g["depth_scalar"]
        assert np.array_equal(ours != 0, ref != 0) and np.allclose(ours, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("kw", [dict(depth_input="scalar"), dict(depth_input="one-hot", depth_bins=118, add_depth_features=True)])
def test_full_size_six_cameras(cuda, kw):
    """config C2/C3 sizes: ~295 k points into 6 cameras of 256 x 704, thousands of pixel collisions."""
    from bevfusion_b200 import synthetic as S
    M = S.lidar_camera_matrices(6, (256, 704), batch=1)
    cloud = S.lidar_cloud(seed=0)
    ours = run_ours(cuda, [cloud], M, (256, 704), **kw)[0]
    gold = oracle.points_to_depth(cloud, M["lidar2image"][0].numpy(), M["img_aug_matrix"][0].numpy(),
                                  M["lidar_aug_matrix"][0].numpy(), (256, 704), **oracle_kw(kw))
    assert (gold != 0).sum() > 20000
    assert np.array_equal(ours, gold)


def test_edge_cases(cuda):
    from bevfusion_b200 import synthetic as S
    from bevfusion_b200.vtransform import points_to_depth
    M = S.lidar_camera_matrices(2, (64, 176), batch=1)
    args = (M["lidar2image"].to(cuda), M["img_aug_matrix"].to(cuda), M["lidar_aug_matrix"].to(cuda), (64, 176))
    empty = torch.zeros(0, 5, device=cuda)
    d = points_to_depth([empty], *args)
    assert d.shape == (1, 2, 1, 64, 176) and float(d.abs().sum()) == 0.0
    # NaN / inf points and points behind every camera leave no trace
    bad = torch.tensor([[float("nan"), 0, 0, 0, 0], [float("inf"), 1, 1, 0, 0], [0, 0, 0, 0, 0]], device=cuda)
    d = points_to_depth([bad], *args, add_depth_features=True)
    gold = oracle.points_to_depth(bad.cpu().numpy(), M["lidar2image"][0].numpy(), M["img_aug_matrix"][0].numpy(),
                                  M["lidar_aug_matrix"][0].numpy(), (64, 176), add_depth_features=True)
    assert np.array_equal(d[0].cpu().numpy(), gold)
    # radar-style height expansion (base.py:266-270): 8 copies at z = 0.25 .. 2.0
    cloud = torch.from_numpy(S.lidar_cloud(seed=5, sweeps=1)[::9].copy()).to(cuda)
    d = points_to_depth([cloud], *args, height_expand=True)
    rep = cloud.repeat_interleave(8, dim=0)
    rep[:, 2] = torch.arange(0.25, 2.25, 0.25, device=cuda).repeat(cloud.shape[0])
    gold = oracle.points_to_depth(rep.cpu().numpy(), M["lidar2image"][0].numpy(), M["img_aug_matrix"][0].numpy(),
                                  M["lidar_aug_matrix"][0].numpy(), (64, 176))
    assert np.array_equal(d[0].cpu().numpy(), gold)
    with pytest.raises(ValueError):
        points_to_depth([cloud], *args, depth_input="one-hot")
    with pytest.raises(Exception):
        points_to_depth([cloud.cpu()], *args)
