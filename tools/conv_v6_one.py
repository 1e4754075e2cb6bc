"""Dev tool: run generation-6 sparse conv on ONE encoder layer (for ncu): python tools/conv_v6_one.py <layer index>
layers: 0 in, 1 s1 subm, 2 s1 down, 3 s2 subm, 4 s2 down, 5 s3 subm, 6 s3 down, 7 s4 subm, 8 out"""
import os, sys
os.environ.setdefault("BEVB200_SPCONV_TC_VARIANT", "5")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bevfusion_b200 import _C, synthetic as S
from bevfusion_b200.spconv import ops
from bevfusion_b200.voxelize import Voxelization, voxelize_mean
dev = torch.device("cuda:0")
L = S.LIDAR_C3
pts = torch.from_numpy(S.lidar_cloud(seed=0)).to(dev)
v, c, n = Voxelization(L["voxel_size"], L["point_cloud_range"], L["max_num_points"], L["max_voxels"]).eval()(pts)
_, idx = voxelize_mean(v, c, n, 0)
shape = L["sparse_shape"]
want = [int(a) for a in sys.argv[1:]] or [3]
lib = _C.lib()
layers = [(5, 16, True, 3, 1, 1), (16, 16, True, 3, 1, 1), (16, 32, False, 3, 2, 1), (32, 32, True, 3, 1, 1),
          (32, 64, False, 3, 2, 1), (64, 64, True, 3, 1, 1), (64, 128, False, 3, 2, [1, 1, 0]), (128, 128, True, 3, 1, 1),
          (128, 128, False, [1, 1, 3], [1, 1, 2], 0)]
for li, (cin, cout, subm, ks, st, pd) in enumerate(layers):
    rb, oshape = ops.get_rulebook(idx, 1, shape, ks, st, pd, 1, 0, subm)
    if li in want:
        n_in = idx.shape[0]
        f = torch.randn(n_in, cin, device=dev)
        kv = rb.nbr.shape[0]
        w = torch.randn(kv, cin, cout, device=dev) / (cin * 5)
        ce = lib.bevb200_spconv_split_channels(cin)
        fs = torch.empty((n_in, ce * 4), dtype=torch.uint8, device=dev)
        _C.check(lib.bevb200_spconv_split_rows(_C.ptr(f), n_in, 0, cin, _C.ptr(fs), _C.current_stream(dev)), "split")
        pk = torch.empty(lib.bevb200_spconv_split_weight_bytes(cin, cout, kv), dtype=torch.uint8, device=dev)
        _C.check(lib.bevb200_spconv_pack_split_weights(_C.ptr(w), cin, cout, kv, _C.ptr(pk), _C.current_stream(dev)), "pack")
        out = torch.empty((rb.n_out, cout), dtype=torch.float32, device=dev)
        osp = torch.empty((rb.n_out, cout * 4), dtype=torch.uint8, device=dev)
        for _ in range(3):
            _C.check(lib.bevb200_spconv_forward_split(_C.ptr(fs), _C.ptr(pk), _C.ptr(rb.nbr), rb.n_out, n_in, rb.n_out, 0, ce,
                                                      cout, kv, 0, 0, 0, 1, _C.ptr(out), _C.ptr(osp), _C.current_stream(dev)), "fwd")
        torch.cuda.synchronize()
    if not subm:
        idx, shape = rb.outids, oshape
