"""Dev tool: one sparse-conv backward of one C3 encoder layer shape (for ncu captures).
    python tools/wgrad_one.py <layer 0..8> [reps]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from bevfusion_b200 import synthetic as S
from bevfusion_b200.spconv import ops
from bevfusion_b200.voxelize import Voxelization, voxelize_mean

which = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
L = S.LIDAR_C3
pts = torch.from_numpy(S.lidar_cloud(seed=0)).to(dev)
vox = Voxelization(L["voxel_size"], L["point_cloud_range"], L["max_num_points"], L["max_voxels"]).eval()
v, c, n = vox(pts)
_, idx = voxelize_mean(v, c, n, 0)
shape = L["sparse_shape"]
layers = [("in", 5, 16, True, 3, 1, 1), ("s1 subm", 16, 16, True, 3, 1, 1), ("s1 down", 16, 32, False, 3, 2, 1),
          ("s2 subm", 32, 32, True, 3, 1, 1), ("s2 down", 32, 64, False, 3, 2, 1),
          ("s3 subm", 64, 64, True, 3, 1, 1), ("s3 down", 64, 128, False, 3, 2, [1, 1, 0]),
          ("s4 subm", 128, 128, True, 3, 1, 1), ("out", 128, 128, False, [1, 1, 3], [1, 1, 2], 0)]
for i, (name, cin, cout, subm, ks, st, pd) in enumerate(layers):
    rb, oshape = ops.get_rulebook(idx, 1, shape, ks, st, pd, 1, 0, subm)
    if i == which:
        n_in = idx.shape[0]
        f = torch.randn(n_in, cin, device=dev)
        w = torch.randn(rb.nbr.shape[0], cin, cout, device=dev) / (cin * 5)
        g = torch.randn(rb.n_out, cout, device=dev)
        nbr_t = ops.transpose_nbr(rb.nbr, n_in)
        for _ in range(reps):
            ops.sparse_conv_backward(f, w, g, rb.nbr, nbr_t=nbr_t, precision=3)
        torch.cuda.synchronize()
        print("DONE", name)
        break
    if not subm:
        idx, shape = rb.outids, oshape
