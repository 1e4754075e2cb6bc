"""CPU: pins the oracle (oracle/) against fixtures produced by the reference's own code
(tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

import oracle


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_hard_voxelize_matches_reference_cpu(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "voxelize_%s.npz" % name))
    v, c, n, m = oracle.hard_voxelize(g["points"], g["voxel_size"], g["coors_range"],
                                      int(g["max_points"]), int(g["max_voxels"]))
    assert m == int(g["voxel_num"])
    assert np.array_equal(c, g["coors"])          # bit-exact coords and order
    assert np.array_equal(n, g["num_points"])     # bit-exact counts
    assert np.array_equal(v, g["voxels"])         # bit-exact point payloads / slots
    dyn = oracle.dynamic_voxelize(g["points"], g["voxel_size"], g["coors_range"])
    ref = g["dyn_coors"]
    assert np.array_equal(dyn[:, 0] == -1, ref[:, 0] == -1)
    ok = ref[:, 0] != -1
    assert np.array_equal(dyn[ok], ref[ok])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden",
                                                                "spconv_*.npz"))))
def test_spconv_matches_reference_cpu(path):
    g = np.load(path)
    outids, pairs, num, out_shape = oracle.get_indice_pairs(
        g["indices"], int(g["batch_size"]), list(g["spatial_shape"]), list(g["ksize"]),
        list(g["stride"]), list(g["padding"]), [1, 1, 1], bool(g["subm"]))
    assert out_shape == list(g["out_shape"])
    assert np.array_equal(outids, g["outids"])           # bit-exact indices / order
    assert np.array_equal(pairs, g["indice_pairs"])      # bit-exact rulebook
    assert np.array_equal(num, g["indice_num"])
    w = g["weight"].reshape(-1, g["weight"].shape[-2], g["weight"].shape[-1])
    out32 = oracle.indice_conv(g["features"], w, pairs, num, outids.shape[0], False,
                               bool(g["subm"]), acc64=False)
    out64 = oracle.indice_conv(g["features"], w, pairs, num, outids.shape[0], False,
                               bool(g["subm"]), acc64=True)
    scale = np.abs(g["out"]).max()
    assert np.abs(out32 - g["out"]).max() <= 1e-5 * scale
    assert np.abs(out64 - g["out"]).max() <= 1e-5 * scale


def test_bev_pool_matches_reference_quickcumsum(golden_dir):
    g = np.load(os.path.join(golden_dir, "bev_pool_quickcumsum.npz"))
    B, D, H, W = (int(v) for v in g["dims"])
    feats, coords = g["feats"], g["coords"]
    ranks = oracle.ranks_of(coords, B, D, H, W)
    order, rs, starts, lengths = oracle.sort_and_intervals(ranks)
    # the restated QuickCumsum reproduces the reference's own CPU path up to the fp32
    # cumsum association order (torch scans blockwise, numpy serially; cancellation ~1e-4)
    qc = oracle.quick_cumsum(feats[order], rs)
    assert qc.shape == g["pooled"].shape
    assert np.abs(qc - g["pooled"]).max() < 5e-4
    assert np.array_equal(coords[order][starts + lengths - 1], g["pooled_geom"])
    # gold (float64 interval sums) agrees with the reference within QuickCumsum's own
    # cancellation error
    out = oracle.bev_pool_forward(feats[order], coords[order].astype(np.int32), lengths, starts,
                                  B, D, H, W, acc64=True)
    cs = coords[order][starts]
    got = out[cs[:, 3], cs[:, 2], cs[:, 0], cs[:, 1]]
    assert np.abs(got - g["pooled"]).max() < 5e-4
    # and the full op output has zeros exactly where no interval lands
    mask = np.zeros((B, D, H, W), dtype=bool)
    mask[cs[:, 3], cs[:, 2], cs[:, 0], cs[:, 1]] = True
    assert np.all(out[~mask] == 0)


def test_quantize_matches_torch_expression():
    """base.py:149 in torch (fp32 sub, div, .long()) == oracle.quantize_filter."""
    import torch
    rng = np.random.default_rng(5)
    geom = rng.uniform(-60, 60, size=(20000, 3)).astype(np.float32)
    geom[:50] = np.array([[-54.0, 54.0, -10.0]], dtype=np.float32)    # on the boundaries
    geom[50:100] = np.array([[-54.0 - 1e-4, 1.0, 0.0]], dtype=np.float32)  # x in (-1, 0) -> truncates to 0: kept
    dx, bx, nx = oracle.gen_dx_bx([-54.0, 54.0, 0.3], [-54.0, 54.0, 0.3], [-10.0, 10.0, 20.0])
    coords, kept = oracle.quantize_filter(geom, dx, bx, nx, 1)
    tdx, tbx = torch.from_numpy(dx), torch.from_numpy(bx)
    t = ((torch.from_numpy(geom) - (tbx - tdx / 2.0)) / tdx).long().numpy()
    assert np.array_equal(coords[:, :3], t)
    assert kept[50:100].all() and (coords[50:100, 0] == 0).all()


def test_hard_voxelize_edge_cases():
    vs, cr = [0.5, 0.5, 0.5], [0, 0, 0, 4, 4, 2]          # non-cubic grid (8, 8, 4)
    # empty cloud
    v, c, n, m = oracle.hard_voxelize(np.zeros((0, 4), np.float32), vs, cr, 3, 10)
    assert m == 0 and v.shape == (0, 3, 4)
    # all points out of range
    pts = np.full((17, 4), 100.0, np.float32)
    assert oracle.hard_voxelize(pts, vs, cr, 3, 10)[3] == 0
    # all points in one voxel: one voxel, max_points kept, in index order
    pts = np.tile(np.array([[0.1, 0.1, 0.1, 0.0]], np.float32), (9, 1))
    pts[:, 3] = np.arange(9)
    v, c, n, m = oracle.hard_voxelize(pts, vs, cr, 3, 10)
    assert m == 1 and n[0] == 3 and list(v[0, :, 3]) == [0.0, 1.0, 2.0]
    # max_voxels cap: later voxels dropped, later points of kept voxels still land
    pts = np.array([[0.1, 0.1, 0.1, 0], [0.6, 0.1, 0.1, 1], [1.1, 0.1, 0.1, 2], [0.1, 0.1, 0.1, 3]],
                   np.float32)
    v, c, n, m = oracle.hard_voxelize(pts, vs, cr, 3, 2)
    assert m == 2 and list(n) == [2, 1] and c.tolist() == [[0, 0, 0], [1, 0, 0]]


def test_vtransform_matches_reference_source(golden_dir):
    """get_geometry + the bev_pool index glue, pinned to a fixture produced by exec'ing the
    reference's own vtransforms/base.py on the CPU (tests/golden/make_golden.py::gen_vtransform)."""
    import torch
    from bevfusion_b200 import synthetic as S
    from bevfusion_b200.vtransform import create_frustum, gen_dx_bx, get_geometry
    g = np.load(os.path.join(golden_dir, "vtransform_tiny.npz"))
    cfg = S.CONFIGS["tiny"]
    frustum = create_frustum(cfg["image_size"], cfg["feature_size"], cfg["dbound"])
    assert np.array_equal(frustum.numpy(), g["frustum"])                       # bit-exact frustum
    dx, bx, nx = gen_dx_bx(cfg["xbound"], cfg["ybound"], cfg["zbound"])
    assert np.array_equal(dx.numpy(), g["dx"]) and np.array_equal(bx.numpy(), g["bx"])
    assert nx.tolist() == g["nx"].tolist()
    rig = S.camera_rig(cfg["n_cam"], cfg["image_size"], batch=2)
    geom = get_geometry(frustum, rig["camera2lidar_rots"], rig["camera2lidar_trans"], rig["intrins"],
                        rig["post_rots"], rig["post_trans"], torch.from_numpy(g["extra_rots"]),
                        torch.from_numpy(g["extra_trans"]))
    assert np.array_equal(geom.numpy(), g["geom"])                             # same fp32 op chain
    # oracle quantise + filter reproduces the coords the reference hands to the op, row for row
    odx, obx, onx = oracle.gen_dx_bx(cfg["xbound"], cfg["ybound"], cfg["zbound"])
    coords, kept = oracle.quantize_filter(g["geom"], odx, obx, onx, 2)
    assert int(kept.sum()) == int(g["x_rows"])
    assert np.array_equal(coords[kept], g["coords"])
    B, D, H, W = (int(v) for v in g["dims"])
    assert (B, D, H, W) == (2, int(onx[2]), int(onx[0]), int(onx[1]))


def test_depth_images_vs_reference_source(golden_dir):
    """oracle.points_to_depth against BaseDepthTransform.forward's depth tensor produced by the
    reference source itself (tests/golden/make_golden.py::gen_depth): identical pixel sets, one-hot
    bins and feature channels bit-equal, scalar distances within 1e-5 (the reference's matmul /
    torch.inverse round differently in the last bit)."""
    g = np.load(os.path.join(golden_dir, "depth_tiny.npz"))
    cases = (("scalar", dict(depth_input="scalar")),
             ("onehot_feats", dict(depth_input="one-hot", depth_bins=int(g["bins_onehot_feats"]),
                                   add_depth_features=True)))
    for tag, kw in cases:
        gold = g["depth_" + tag]
        for b in range(2):
            d = oracle.points_to_depth(g["points%d" % b], g["lidar2image"][b], g["img_aug_matrix"][b],
                                       g["lidar_aug_matrix"][b], g["image_size"], **kw)
            assert d.shape == gold[b].shape
            assert (gold[b] != 0).sum() > 500
            assert np.array_equal(d != 0, gold[b] != 0)
            if tag == "scalar":
                assert np.allclose(d, gold[b], rtol=1e-5, atol=1e-5)
            else:
                assert np.array_equal(d, gold[b])


def test_depth_oracle_last_point_wins():
    """colliding points: the fancy-index assignment in the oracle must mean 'largest index wins'."""
    eye = np.eye(4, dtype=np.float32)
    l2i = eye[None].copy()
    pts = np.array([[2.0, 3.0, 1.0, 0.1, 0.0], [2.2, 3.4, 1.0, 0.2, 0.0], [4.0, 1.0, 2.0, 0.3, 0.0],
                    [2.1, 3.9, 1.0, 0.4, 0.0]], np.float32)
    d = oracle.points_to_depth(pts, l2i, eye[None], eye, (8, 8), add_depth_features=True)
    assert d[0, 0, 3, 2] == 1.0 and d[0, 4, 3, 2] == np.float32(0.4)     # point 3 beats 0 and 1
    assert d[0, 0, 0, 2] == 2.0 and d[0, 4, 0, 2] == np.float32(0.3)     # (4,1)/2 -> col 2,row 0
    assert (d[0, 0] != 0).sum() == 2


def test_dynamic_scatter_oracle_properties():
    """the restatement of dynamic_point_to_voxel_forward / _backward (scatter_points_cuda.cu:187-315):
    lexicographic unique rows, negative rows dropped, permutation invariance, gradient routing."""
    rng = np.random.default_rng(5)
    n, c = 400, 3
    coors = rng.integers(0, 5, (n, 3)).astype(np.int32)
    coors[rng.random(n) < 0.1, 1] = -1
    feats = rng.standard_normal((n, c)).astype(np.float32)
    for red in ("sum", "mean", "max"):
        r, oc, cmap, cnt = oracle.dynamic_scatter(feats, coors, red)
        valid = (coors >= 0).all(1)
        assert (cmap[~valid] == -1).all() and (cmap[valid] >= 0).all()
        assert cnt.sum() == valid.sum() and (cnt > 0).all()
        assert np.array_equal(oc, np.unique(coors[valid], axis=0))           # sorted unique rows
        assert np.array_equal(oc[cmap[valid]], coors[valid])                 # the map points at the right voxel
        perm = rng.permutation(n)
        r2, oc2, cmap2, cnt2 = oracle.dynamic_scatter(feats[perm], coors[perm], red)
        assert np.array_equal(oc2, oc) and np.array_equal(cnt2, cnt) and np.array_equal(cmap2, cmap[perm])
        assert np.allclose(r2, r, rtol=1e-6, atol=1e-6)
        g = rng.standard_normal(r.shape).astype(np.float32)
        d = oracle.dynamic_scatter_backward(g, feats, r, cmap, cnt, red)
        assert (d[~valid] == 0).all()
        if red == "sum":
            assert np.array_equal(d[valid], g[cmap[valid]])
        elif red == "mean":
            assert np.allclose(d[valid], g[cmap[valid]] / cnt[cmap[valid]][:, None])
        else:
            # every (voxel, channel) gradient lands on exactly one point: the first that attains the maximum
            for v in range(r.shape[0]):
                rows = np.nonzero(cmap == v)[0]
                for ch in range(c):
                    hit = rows[feats[rows, ch] == r[v, ch]]
                    assert d[hit[0], ch] == g[v, ch] and (d[rows, ch] != 0).sum() <= 1
    e = oracle.dynamic_scatter(feats[:0], coors[:0], "max")
    assert e[0].shape == (0, c) and e[2].shape == (0,)
