"""Generates tests/golden/*.npz from the REFERENCE's own code, run in the build container.

  voxelize_*.npz   reference CPU hard_voxelize / dynamic_voxelize (oracle/_ref/voxel_layer_ref.so,
                   built unmodified from /root/reference/mmdet3d/ops/voxel/src).  Cubic grids only:
                   the reference CPU path indexes its lookup tensor out of bounds on non-cubic
                   grids (voxelization_cpu.cpp:75 vs :129-130).
  spconv_*.npz     reference CPU get_indice_pairs_3d + indice_conv_fp32
                   (oracle/_ref/sparse_conv_ext_ref.so) for the four conv geometries of the
                   VoxelNet encoder (SubM k3; k3 s2 p1; k3 s2 p(1,1,0); k(1,1,3) s(1,1,2)).
  bev_pool_quickcumsum.npz  the reference's pure-torch QuickCumsum (bev_pool.py:9-35), executed
                   from the reference source file with a stub for the compiled extension import.

Run:  python tests/golden/make_golden.py      (needs /root/reference and oracle/_ref)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from oracle.build_ref import load_ref  # noqa: E402

REF = os.environ.get("BEVFUSION_REFERENCE", "/root/reference")


def gen_voxelize():
    vl = load_ref("voxel_layer_ref")
    rng = np.random.default_rng(1)
    cases = {
        # name: (N, voxel_size, range, max_points, max_voxels)
        "a": (3000, [0.5, 0.5, 0.5], [0, 0, 0, 8, 8, 8], 5, 300),       # both caps bind
        "b": (2000, [1.0, 1.0, 1.0], [-4, -4, -4, 4, 4, 4], 10, 2000),   # no cap binds
        "c": (500, [2.0, 2.0, 2.0], [0, 0, 0, 8, 8, 8], 3, 64),          # dense voxels
    }
    for name, (n, vs, cr, mp, mv) in cases.items():
        lo, hi = np.array(cr[:3], dtype=np.float32), np.array(cr[3:], dtype=np.float32)
        pts = rng.uniform(lo - 1.0, hi + 1.0, size=(n, 3)).astype(np.float32)
        pts = np.concatenate([pts, rng.uniform(0, 1, size=(n, 2)).astype(np.float32)], axis=1)
        pts[::97, 0] = lo[0]            # exactly on the lower bound
        pts[5::101, 1] = hi[1]          # exactly on the upper bound (out)
        p = torch.from_numpy(pts)
        voxels = torch.zeros(mv, mp, 5)
        coors = torch.zeros(mv, 3, dtype=torch.int32)
        num = torch.zeros(mv, dtype=torch.int32)
        m = vl.hard_voxelize(p, voxels, coors, num, [float(v) for v in vs], [float(v) for v in cr],
                             mp, mv, 3, True)
        dyn = torch.zeros(n, 3, dtype=torch.int32)
        vl.dynamic_voxelize(p, dyn, [float(v) for v in vs], [float(v) for v in cr], 3)
        np.savez_compressed(os.path.join(HERE, "voxelize_%s.npz" % name), points=pts,
                            voxel_size=np.array(vs, np.float32), coors_range=np.array(cr, np.float32),
                            max_points=mp, max_voxels=mv, voxel_num=m, voxels=voxels[:m].numpy(),
                            coors=coors[:m].numpy(), num_points=num[:m].numpy(), dyn_coors=dyn.numpy())


def gen_spconv():
    sp = load_ref("sparse_conv_ext_ref")
    rng = np.random.default_rng(2)
    shape, B = [14, 12, 9], 2
    allc = np.stack(np.meshgrid(np.arange(B), *[np.arange(s) for s in shape], indexing="ij"), -1).reshape(-1, 4)
    idx = allc[rng.permutation(len(allc))[:260]].astype(np.int32)
    geoms = {
        "subm_k3": ([3, 3, 3], [1, 1, 1], [1, 1, 1], True),
        "conv_k3s2p1": ([3, 3, 3], [2, 2, 2], [1, 1, 1], False),
        "conv_k3s2p110": ([3, 3, 3], [2, 2, 2], [1, 1, 0], False),
        "conv_k113s112": ([1, 1, 3], [1, 1, 2], [0, 0, 0], False),
    }
    for name, (ks, st, pd, subm) in geoms.items():
        cin, cout = (5, 16) if subm else (16, 32)
        feat = rng.standard_normal((idx.shape[0], cin)).astype(np.float32)
        W = (rng.standard_normal((*ks, cin, cout)) / np.sqrt(cin * 9)).astype(np.float32)
        if subm:
            out_shape = shape
        else:
            out_shape = [(shape[i] + 2 * pd[i] - (ks[i] - 1) - 1) // st[i] + 1 for i in range(3)]
        r = sp.get_indice_pairs_3d(torch.from_numpy(idx), B, out_shape, shape, ks, st, pd, [1, 1, 1],
                                   [0, 0, 0], int(subm), 0)
        out = sp.indice_conv_fp32(torch.from_numpy(feat), torch.from_numpy(W), r[1], r[2],
                                  r[0].shape[0], 0, int(subm))
        np.savez_compressed(os.path.join(HERE, "spconv_%s.npz" % name), indices=idx, features=feat,
                            weight=W, batch_size=B, spatial_shape=np.array(shape),
                            out_shape=np.array(out_shape), ksize=np.array(ks), stride=np.array(st),
                            padding=np.array(pd), subm=int(subm), outids=r[0].numpy(),
                            indice_pairs=r[1].numpy(), indice_num=r[2].numpy(), out=out.numpy())


def load_reference_quickcumsum():
    """exec mmdet3d/ops/bev_pool/bev_pool.py from the reference tree with the compiled-extension
    import stubbed (only QuickCumsum, which is pure torch, is used)."""
    pkg = types.ModuleType("_refpkg")
    pkg.__path__ = []
    pkg.bev_pool_ext = types.SimpleNamespace()
    sys.modules["_refpkg"] = pkg
    src = open(os.path.join(REF, "mmdet3d/ops/bev_pool/bev_pool.py")).read()
    mod = types.ModuleType("_refpkg.bev_pool")
    mod.__package__ = "_refpkg"
    exec(compile(src, "reference:bev_pool.py", "exec"), mod.__dict__)
    return mod


def gen_bev_pool():
    mod = load_reference_quickcumsum()
    g = torch.Generator().manual_seed(3)
    n, c, B, D, H, W = 6000, 16, 2, 2, 12, 10
    feats = torch.randn(n, c, generator=g)
    coords = torch.stack([torch.randint(0, H, (n,), generator=g), torch.randint(0, W, (n,), generator=g),
                          torch.randint(0, D, (n,), generator=g), torch.randint(0, B, (n,), generator=g)], 1)
    ranks = coords[:, 0] * (W * D * B) + coords[:, 1] * (D * B) + coords[:, 2] * B + coords[:, 3]
    indices = ranks.argsort(stable=True)
    fs, cs, rs = feats[indices], coords[indices], ranks[indices]
    pooled, geom = mod.QuickCumsum.apply(fs, cs, rs)     # reference CPU path
    np.savez_compressed(os.path.join(HERE, "bev_pool_quickcumsum.npz"), feats=feats.numpy(),
                        coords=coords.numpy(), dims=np.array([B, D, H, W]), pooled=pooled.numpy(),
                        pooled_geom=geom.numpy())


def load_reference_base_transform(captured):
    """exec mmdet3d/models/vtransforms/base.py from the reference tree with stubs for the two
    imports it cannot satisfy here: mmcv.runner.force_fp32 (identity decorator) and
    mmdet3d.ops.bev_pool (records its arguments and returns zeros of the op's output shape)."""
    mmcv = types.ModuleType("mmcv"); runner = types.ModuleType("mmcv.runner")
    runner.force_fp32 = lambda *a, **k: (lambda f: f)
    mmcv.runner = runner
    mm = types.ModuleType("mmdet3d"); ops = types.ModuleType("mmdet3d.ops")

    def bev_pool_stub(x, geom_feats, B, D, H, W):
        captured["x_rows"] = x.shape[0]
        captured["coords"] = geom_feats.clone()
        captured["dims"] = (int(B), int(D), int(H), int(W))
        return torch.zeros(int(B), x.shape[1], int(D), int(H), int(W))
    ops.bev_pool = bev_pool_stub
    mm.ops = ops
    saved = {k: sys.modules.get(k) for k in ("mmcv", "mmcv.runner", "mmdet3d", "mmdet3d.ops")}
    sys.modules.update({"mmcv": mmcv, "mmcv.runner": runner, "mmdet3d": mm, "mmdet3d.ops": ops})
    try:
        src = open(os.path.join(REF, "mmdet3d/models/vtransforms/base.py")).read()
        mod = types.ModuleType("_ref_vtransform_base")
        exec(compile(src, "reference:vtransforms/base.py", "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def gen_vtransform():
    """get_geometry (base.py:92-135) and the index glue of BaseTransform.bev_pool (:141-169) run
    from the REFERENCE source on the CPU for a 2-camera rig with image / lidar augmentation."""
    from bevfusion_b200 import synthetic as S
    captured = {}
    mod = load_reference_base_transform(captured)
    cfg = S.CONFIGS["tiny"]
    t = mod.BaseTransform(in_channels=8, out_channels=cfg["C"], image_size=cfg["image_size"],
                          feature_size=cfg["feature_size"], xbound=cfg["xbound"], ybound=cfg["ybound"],
                          zbound=cfg["zbound"], dbound=cfg["dbound"])
    B = 2
    rig = S.camera_rig(cfg["n_cam"], cfg["image_size"], batch=B)
    g = torch.Generator().manual_seed(4)
    # a lidar augmentation (rotation about z + translation), different per sample
    ang = torch.tensor([0.1, -0.25])
    extra_rots = torch.stack([torch.tensor([[torch.cos(a), -torch.sin(a), 0.0], [torch.sin(a), torch.cos(a), 0.0],
                                            [0.0, 0.0, 1.0]]) for a in ang])
    extra_trans = torch.tensor([[0.5, -0.25, 0.1], [-1.0, 0.75, 0.0]])
    geom = t.get_geometry(rig["camera2lidar_rots"], rig["camera2lidar_trans"], rig["intrins"], rig["post_rots"],
                          rig["post_trans"], extra_rots=extra_rots, extra_trans=extra_trans)
    D, fH, fW = geom.shape[2:5]
    x = torch.randn(B, cfg["n_cam"], D, fH, fW, cfg["C"], generator=g)
    t.bev_pool(geom, x)                         # records the op's inputs
    np.savez_compressed(os.path.join(HERE, "vtransform_tiny.npz"), frustum=t.frustum.detach().numpy(),
                        dx=t.dx.detach().numpy(), bx=t.bx.detach().numpy(), nx=t.nx.detach().numpy(),
                        extra_rots=extra_rots.numpy(), extra_trans=extra_trans.numpy(),
                        geom=geom.detach().numpy(), coords=captured["coords"].numpy(),
                        dims=np.array(captured["dims"]), x_rows=captured["x_rows"])


def gen_depth():
    """BaseDepthTransform.forward's LiDAR depth images (base.py:279-329) run from the REFERENCE source
    on the CPU (single thread, so the index_put at :319 is sequential: last point wins) for a
    2-camera rig with image + lidar augmentation: 'scalar' depth and 'one-hot' + point features."""
    from bevfusion_b200 import synthetic as S
    captured = {}
    mod = load_reference_base_transform(captured)
    cfg = S.CONFIGS["tiny"]

    class Stop(Exception):
        pass

    class Probe(mod.BaseDepthTransform):
        def get_cam_feats(self, img, depth, mats_dict):
            captured["depth"] = depth.clone()
            raise Stop()

    torch.set_num_threads(1)
    B = 2
    M = S.lidar_camera_matrices(cfg["n_cam"], cfg["image_size"], B)
    clouds = [S.lidar_cloud(seed=11 + b, sweeps=1)[::7].copy() for b in range(B)]
    out = {}
    for tag, kw in (("scalar", dict(depth_input="scalar", add_depth_features=False)),
                    ("onehot_feats", dict(depth_input="one-hot", add_depth_features=True))):
        t = Probe(in_channels=8, out_channels=cfg["C"], image_size=cfg["image_size"],
                  feature_size=cfg["feature_size"], xbound=cfg["xbound"], ybound=cfg["ybound"],
                  zbound=cfg["zbound"], dbound=cfg["dbound"], use_points="lidar", height_expand=False, **kw)
        pts = [torch.from_numpy(c.copy()) for c in clouds]      # the reference shifts these in place
        img = torch.zeros(B, cfg["n_cam"], 3, *cfg["image_size"])
        try:
            t(img, pts, None, M["camera2lidar"], torch.eye(4).repeat(B, 1, 1), M["lidar2camera"], M["lidar2image"],
              M["cam_intrinsic"], M["camera2lidar"], M["img_aug_matrix"], M["lidar_aug_matrix"], None)
        except Stop:
            pass
        out["depth_" + tag] = captured["depth"].numpy()
        out["bins_" + tag] = t.D
    np.savez_compressed(os.path.join(HERE, "depth_tiny.npz"), points0=clouds[0], points1=clouds[1],
                        lidar2image=M["lidar2image"].numpy(), img_aug_matrix=M["img_aug_matrix"].numpy(),
                        lidar_aug_matrix=M["lidar_aug_matrix"].numpy(),
                        image_size=np.array(cfg["image_size"]), **out)


if __name__ == "__main__":
    gen_vtransform()
    gen_depth()
    gen_voxelize()
    gen_spconv()
    gen_bev_pool()
    print(sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
