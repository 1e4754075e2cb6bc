"""Deterministic synthetic inputs with the shapes / statistics of the reference's nuScenes
configs.  numpy / torch on the host; no dataset, no network."""
import math

import numpy as np
import torch

from .vtransform import create_frustum, get_geometry

# BASELINE.json configs, made concrete
CONFIGS = {
    # C1: bev_pool correctness case (1 cam, 64x176 features, D=60, 128x128 grid, C=64)
    "C1": dict(n_cam=1, image_size=(512, 1408), feature_size=(64, 176), dbound=(1.0, 61.0, 1.0),
               xbound=(-51.2, 51.2, 0.8), ybound=(-51.2, 51.2, 0.8), zbound=(-10.0, 10.0, 20.0), C=64),
    # C2: camera+lidar/swint_v0p075 view transform (6 cam, 256x704 -> 32x88, D=118, 360x360, C=80)
    "C2": dict(n_cam=6, image_size=(256, 704), feature_size=(32, 88), dbound=(1.0, 60.0, 0.5),
               xbound=(-54.0, 54.0, 0.3), ybound=(-54.0, 54.0, 0.3), zbound=(-10.0, 10.0, 20.0), C=80),
    # C2 literal: 180x180 BEV grid as BASELINE.json words it
    "C2_180": dict(n_cam=6, image_size=(256, 704), feature_size=(32, 88), dbound=(1.0, 60.0, 0.5),
                   xbound=(-54.0, 54.0, 0.6), ybound=(-54.0, 54.0, 0.6), zbound=(-10.0, 10.0, 20.0), C=80),
    # C5: stress (6 cam 512x1408 -> 64x176, D=200, 256x256, C=80)
    "C5": dict(n_cam=6, image_size=(512, 1408), feature_size=(64, 176), dbound=(1.0, 61.0, 0.3),
               xbound=(-51.2, 51.2, 0.4), ybound=(-51.2, 51.2, 0.4), zbound=(-10.0, 10.0, 20.0), C=80),
    # tiny case for smoke / fast tests
    "tiny": dict(n_cam=2, image_size=(64, 176), feature_size=(8, 22), dbound=(1.0, 21.0, 1.0),
                 xbound=(-16.0, 16.0, 0.5), ybound=(-16.0, 16.0, 0.5), zbound=(-10.0, 10.0, 20.0), C=80),
}

LIDAR_C3 = dict(voxel_size=[0.075, 0.075, 0.2], point_cloud_range=[-54.0, -54.0, -5.0, 54.0, 54.0, 3.0],
                max_num_points=10, max_voxels=(120000, 160000), sparse_shape=[1440, 1440, 41])


def camera_rig(n_cam=6, image_size=(256, 704), batch=1):
    """nuScenes-shaped 6-camera rig: yaw {0,-55,+55,180,+110,-110} deg,
    fx=fy=1266 (rear 809), cx=816, cy=491 on 1600x900, image aug = resize then crop to
    image_size (0.48 / (32,176) for 704x256).  Returns a dict of [B, N, ...] fp32 tensors."""
    yaws = [0.0, -55.0, 55.0, 180.0, 110.0, -110.0][:n_cam]
    iH, iW = image_size
    scale = 0.48 * iW / 704.0
    crop_w = (1600 * scale - iW) / 2.0
    crop_h = 900 * scale - iH
    rots, trans, intr, prot, ptr = [], [], [], [], []
    R0 = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    for i, yaw in enumerate(yaws):
        t = math.radians(yaw)
        Rz = torch.tensor([[math.cos(t), -math.sin(t), 0.0], [math.sin(t), math.cos(t), 0.0],
                           [0.0, 0.0, 1.0]])
        rots.append(Rz @ R0)
        trans.append(Rz @ torch.tensor([0.6, 0.0, 0.0]) + torch.tensor([0.0, 0.0, -0.3]))
        f = 809.0 if abs(yaw) == 180.0 else 1266.0
        intr.append(torch.tensor([[f, 0.0, 816.0], [0.0, f, 491.0], [0.0, 0.0, 1.0]]))
        prot.append(torch.diag(torch.tensor([scale, scale, 1.0])))
        ptr.append(torch.tensor([-crop_w, -crop_h, 0.0]))

    def st(v):
        return torch.stack(v).unsqueeze(0).repeat(batch, *([1] * (v[0].dim() + 1))).float().contiguous()

    return dict(camera2lidar_rots=st(rots), camera2lidar_trans=st(trans), intrins=st(intr),
                post_rots=st(prot), post_trans=st(ptr))


def lidar_camera_matrices(n_cam=6, image_size=(256, 704), batch=1, augment=True):
    """4x4 matrices the depth-aware lift consumes (base.py:237-262), consistent with camera_rig:
    lidar2image = K @ inverse(camera2lidar), img_aug_matrix from the resize / crop, and a
    per-sample lidar augmentation (z rotation, scale, translation).  [B, N, 4, 4] / [B, 4, 4] fp32."""
    rig = camera_rig(n_cam, image_size, batch)
    B, N = batch, n_cam
    eye = torch.eye(4).view(1, 1, 4, 4).repeat(B, N, 1, 1)
    cam2lidar = eye.clone()
    cam2lidar[..., :3, :3] = rig["camera2lidar_rots"]
    cam2lidar[..., :3, 3] = rig["camera2lidar_trans"]
    lidar2cam = torch.inverse(cam2lidar)
    K = eye.clone()
    K[..., :3, :3] = rig["intrins"]
    lidar2image = K.matmul(lidar2cam)
    img_aug = eye.clone()
    img_aug[..., :3, :3] = rig["post_rots"]
    img_aug[..., :3, 3] = rig["post_trans"]
    lidar_aug = torch.eye(4).view(1, 4, 4).repeat(B, 1, 1)
    if augment:
        for b in range(B):
            a = 0.1 - 0.17 * b
            sc = 1.0 + 0.05 * (b + 1)
            lidar_aug[b, :3, :3] = sc * torch.tensor([[math.cos(a), -math.sin(a), 0.0],
                                                      [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]])
            lidar_aug[b, :3, 3] = torch.tensor([0.5 - b, -0.25 + 0.5 * b, 0.1 * b])
    out = dict(rig)
    out.update(lidar2image=lidar2image.float().contiguous(), img_aug_matrix=img_aug.float().contiguous(),
               lidar_aug_matrix=lidar_aug.float().contiguous(), camera2lidar=cam2lidar.float().contiguous(),
               lidar2camera=lidar2cam.float().contiguous(), cam_intrinsic=K.float().contiguous())
    return out


def camera_geometry(cfg_name="C2", batch=1, device="cpu"):
    """(geom [B, N, D, fH, fW, 3] fp32, cfg dict) for one of CONFIGS."""
    cfg = CONFIGS[cfg_name]
    rig = camera_rig(cfg["n_cam"], cfg["image_size"], batch)
    frustum = create_frustum(cfg["image_size"], cfg["feature_size"], cfg["dbound"])
    geom = get_geometry(frustum, rig["camera2lidar_rots"], rig["camera2lidar_trans"],
                        rig["intrins"], rig["post_rots"], rig["post_trans"])
    return geom.contiguous().to(device), cfg


def lifted_features(cfg_name, batch=1, device="cpu", seed=0, dtype=torch.float32):
    """x [B, N, D, fH, fW, C] standard-normal lifted camera features."""
    cfg = CONFIGS[cfg_name]
    D = len(np.arange(*cfg["dbound"]))
    fH, fW = cfg["feature_size"]
    g = torch.Generator(device="cpu").manual_seed(seed)
    shape = (batch, cfg["n_cam"], D, fH, fW, cfg["C"])
    if device == "cpu":
        return torch.randn(shape, generator=g, dtype=dtype)
    gd = torch.Generator(device=device).manual_seed(seed)
    return torch.randn(shape, generator=gd, dtype=dtype, device=device)


def lidar_cloud(seed=0, sweeps=10, beams=32, az_steps=1090, shuffle=True):
    """10-sweep, 32-beam synthetic LiDAR cloud: ground plane at z=-1.84 plus
    per-sweep random walls, ~295 k points of (x, y, z, intensity, dt) fp32."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(-30.67, 10.67, beams))
    az = np.linspace(-np.pi, np.pi, az_steps, endpoint=False)
    pts = []
    for s in range(sweeps):
        ox = 0.5 * s  # ego motion between sweeps
        sector_r = rng.uniform(6.0, 60.0, size=64)
        E, A = np.meshgrid(elev, az, indexing="ij")
        sec = ((A + np.pi) / (2 * np.pi) * 64).astype(int) % 64
        wall_r = sector_r[sec] * (1.0 + 0.02 * rng.standard_normal(E.shape))
        with np.errstate(divide="ignore", invalid="ignore"):
            ground_r = np.where(E < 0, 1.84 / np.sin(-E), np.inf)
        hits_wall = (wall_r * np.sin(E) + 1.84) < 4.0
        r = np.minimum(ground_r, np.where(hits_wall, wall_r / np.maximum(np.cos(E), 1e-3), np.inf))
        r = r * (1.0 + 0.003 * rng.standard_normal(E.shape))
        ok = np.isfinite(r) & (r > 1.0) & (r < 75.0)
        r, E_, A_ = r[ok], E[ok], A[ok]
        x = r * np.cos(E_) * np.cos(A_) - ox
        y = r * np.cos(E_) * np.sin(A_)
        z = r * np.sin(E_)
        inten = rng.uniform(0.0, 1.0, size=x.shape)
        dt = np.full(x.shape, 0.05 * s)
        pts.append(np.stack([x, y, z, inten, dt], axis=1))
    pts = np.concatenate(pts, axis=0).astype(np.float32)
    if shuffle:
        pts = pts[rng.permutation(pts.shape[0])]
    return pts


def uniform_cloud(n, seed=0, margin=2.0, rng_range=(-54.0, -54.0, -5.0, 54.0, 54.0, 3.0), nf=5):
    """uniform-random cloud, some points outside the range on every side."""
    rng = np.random.default_rng(seed)
    lo = np.array(rng_range[:3]) - margin
    hi = np.array(rng_range[3:]) + margin
    xyz = rng.uniform(lo, hi, size=(n, 3))
    extra = rng.uniform(0, 1, size=(n, nf - 3))
    return np.concatenate([xyz, extra], axis=1).astype(np.float32)
