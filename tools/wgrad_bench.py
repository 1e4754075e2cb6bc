"""Dev tool: sparse-conv backward (input gradient + filter gradient) per C3 encoder layer, CUDA-event times.
    python tools/wgrad_bench.py            # tensor-core filter gradient (default for 32 / 64 / 128 channels)
    BEVB200_WGRAD_TC=0 python tools/wgrad_bench.py   # SIMT filter gradient"""
import os, sys, statistics
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from bevfusion_b200 import synthetic as S
from bevfusion_b200.spconv import ops
from bevfusion_b200.voxelize import Voxelization, voxelize_mean

dev = torch.device("cuda:0")
L = S.LIDAR_C3
pts = torch.from_numpy(S.lidar_cloud(seed=0)).to(dev)
vox = Voxelization(L["voxel_size"], L["point_cloud_range"], L["max_num_points"], L["max_voxels"]).eval()
v, c, n = vox(pts)
_, idx = voxelize_mean(v, c, n, 0)
shape = L["sparse_shape"]


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ev = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)


layers = [("in", 5, 16, True, 3, 1, 1), ("s1 subm", 16, 16, True, 3, 1, 1), ("s1 down", 16, 32, False, 3, 2, 1),
          ("s2 subm", 32, 32, True, 3, 1, 1), ("s2 down", 32, 64, False, 3, 2, 1),
          ("s3 subm", 64, 64, True, 3, 1, 1), ("s3 down", 64, 128, False, 3, 2, [1, 1, 0]),
          ("s4 subm", 128, 128, True, 3, 1, 1), ("out", 128, 128, False, [1, 1, 3], [1, 1, 2], 0)]
tot = 0.0
for name, cin, cout, subm, ks, st, pd in layers:
    rb, oshape = ops.get_rulebook(idx, 1, shape, ks, st, pd, 1, 0, subm)
    n_in = idx.shape[0]
    f = torch.randn(n_in, cin, device=dev)
    kv = rb.nbr.shape[0]
    w = torch.randn(kv, cin, cout, device=dev) / (cin * 5)
    g = torch.randn(rb.n_out, cout, device=dev)
    nbr_t = ops.transpose_nbr(rb.nbr, n_in)
    din, dw = ops.sparse_conv_backward(f, w, g, rb.nbr, nbr_t=nbr_t, precision=3)
    # check two offsets against fp64
    err = 0.0
    for k in (0, kv // 2):
        valid = rb.nbr[k] >= 0
        want = f[rb.nbr[k][valid].long()].double().t() @ g[valid].double()
        err = max(err, float((dw[k].double() - want).abs().max() / (want.abs().max() + 1e-30)))
    t = timeit(lambda: ops.sparse_conv_backward(f, w, g, rb.nbr, nbr_t=nbr_t, precision=3))
    mult = 4 if (subm and cin > 5) else 1
    tot += t * mult
    pairs = int((rb.nbr >= 0).sum())
    print(f"{name:8s} n_out {rb.n_out:7d} {cin:4d}->{cout:4d} pairs {pairs:8d}  backward {t*1e3:8.1f} us  dW rel err {err:.2e}  (x{mult})", flush=True)
    if not subm:
        idx, shape = rb.outids, oshape
print("sum over the 21 convs: backward %.3f ms" % tot)
