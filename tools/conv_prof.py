"""Dev tool: per-role cycle counters of CTA 0 of the v4 sparse-conv kernel.
Needs a profiling build (BEVB200_BUILD_PROFILE=1 python bevfusion_b200/build.py --force) and
BEVB200_SPCONV_TC_VARIANT=4; the default v5 kernel is analysed with BEVB200_TC_DBG (tools/conv_bench.py)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
dev = torch.device("cuda:0")
prof = torch.zeros(8, dtype=torch.int64, device=dev)
os.environ["BEVB200_TC_PROF"] = hex(prof.data_ptr())
from bevfusion_b200 import synthetic as S
from bevfusion_b200.spconv import ops
from bevfusion_b200.voxelize import Voxelization, voxelize_mean
L = S.LIDAR_C3
pts = torch.from_numpy(S.lidar_cloud(seed=0)).to(dev)
v, c, n = Voxelization(L["voxel_size"], L["point_cloud_range"], L["max_num_points"], L["max_voxels"]).eval()(pts)
feats, idx = voxelize_mean(v, c, n, 0)
shape = L["sparse_shape"]
layers = [(16, 16, True, 3, 1, 1), (16, 32, False, 3, 2, 1), (32, 32, True, 3, 1, 1), (32, 64, False, 3, 2, 1),
          (64, 64, True, 3, 1, 1), (64, 128, False, 3, 2, [1, 1, 0]), (128, 128, True, 3, 1, 1)]
names = ["prod: transpose+wait a_empty", "prod: split+st+wait::st+arrive", "prod: issue loads", "prod: drain wait",
         "mma: wait a_full", "mma: wait b_full", "mma: issue+commit", "CTA total"]
for li, (cin, cout, subm, ks, st, pd) in enumerate(layers):
    rb, oshape = ops.get_rulebook(idx, 1, shape, ks, st, pd, 1, 0, subm)
    if subm:
        f = torch.randn(idx.shape[0], cin, device=dev)
        w = torch.randn(rb.nbr.shape[0], cin, cout, device=dev) / (cin * 5)
        packed = ops.pack_weights(w, 1)
        ops.sparse_conv(f, w, rb.nbr, rb.n_out, precision=1, packed=packed)
        torch.cuda.synchronize(); prof.zero_()
        ops.sparse_conv(f, w, rb.nbr, rb.n_out, precision=1, packed=packed)
        torch.cuda.synchronize()
        nkb = 27 * cin // 32
        vals = prof.tolist()
        print(f"C={cin}: K blocks {nkb}; per-K-block cycles:", {nm: round(v / nkb) for nm, v in zip(names[:7], vals[:7])}, "CTA total cycles", vals[7])
    else:
        idx, shape = rb.outids, oshape
