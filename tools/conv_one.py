"""Dev tool: run ONE conv layer config (for ncu): python tools/conv_one.py <layer index>"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bevfusion_b200 import synthetic as S
from bevfusion_b200.spconv import ops
from bevfusion_b200.voxelize import Voxelization, voxelize_mean
dev = torch.device("cuda:0")
L = S.LIDAR_C3
pts = torch.from_numpy(S.lidar_cloud(seed=0)).to(dev)
v, c, n = Voxelization(L["voxel_size"], L["point_cloud_range"], L["max_num_points"], L["max_voxels"]).eval()(pts)
feats, idx = voxelize_mean(v, c, n, 0)
shape = L["sparse_shape"]
want = [int(a) for a in sys.argv[1:]] or [2]
layers = [(16, 16, True, 3, 1, 1), (16, 32, False, 3, 2, 1), (32, 32, True, 3, 1, 1), (32, 64, False, 3, 2, 1),
          (64, 64, True, 3, 1, 1), (64, 128, False, 3, 2, [1, 1, 0]), (128, 128, True, 3, 1, 1)]
for li, (cin, cout, subm, ks, st, pd) in enumerate(layers):
    rb, oshape = ops.get_rulebook(idx, 1, shape, ks, st, pd, 1, 0, subm)
    if li in want:
        f = torch.randn(idx.shape[0], cin, device=dev)
        w = torch.randn(rb.nbr.shape[0], cin, cout, device=dev) / (cin * 5)
        prec = int(os.environ.get("PREC", "3"))
        packed = ops.pack_weights(w, prec)
        for _ in range(3):
            ops.sparse_conv(f, w, rb.nbr, rb.n_out, precision=prec, packed=packed)
        torch.cuda.synchronize()
    if not subm:
        idx, shape = rb.outids, oshape
