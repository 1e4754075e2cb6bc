"""Dev tool: per-layer sparse-conv kernel times on the real C3 rulebooks (CUDA events)."""
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
feats, coords = voxelize_mean(v, c, n, 0)

def timeit(fn, n=10):
    for _ in range(3): fn()
    ev = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)

prec = int(os.environ.get("PREC", "1"))
shape = L["sparse_shape"]
idx = coords
layers = [("s1 subm", 16, 16, True, 3, 1, 1), ("s1 down", 16, 32, False, 3, 2, 1),
          ("s2 subm", 32, 32, True, 3, 1, 1), ("s2 down", 32, 64, False, 3, 2, 1),
          ("s3 subm", 64, 64, True, 3, 1, 1), ("s3 down", 64, 128, False, 3, 2, [1, 1, 0]),
          ("s4 subm", 128, 128, True, 3, 1, 1), ("out", 128, 128, False, [1, 1, 3], [1, 1, 2], 0)]
total = 0.0
for name, cin, cout, subm, ks, st, pd in layers:
    rb, oshape = ops.get_rulebook(idx, 1, shape, ks, st, pd, 1, 0, subm)
    f = torch.randn(idx.shape[0], cin, device=dev)
    kv = rb.nbr.shape[0]
    w = torch.randn(kv, cin, cout, device=dev) / (cin * 5)
    packed = ops.pack_weights(w, prec)
    ms = timeit(lambda: ops.sparse_conv(f, w, rb.nbr, rb.n_out, precision=prec, packed=packed))
    pairs = int((rb.nbr >= 0).sum())
    gf = 2.0 * pairs * cin * cout / 1e9
    mult = 4 if subm else 1
    total += ms * mult
    print(f"{name:8s} n_in {idx.shape[0]:7d} n_out {rb.n_out:7d} {cin:4d}->{cout:4d} pairs {pairs:8d}  {ms*1e3:8.1f} us  {gf/ms:8.1f} TFLOP/s useful  (x{mult})", flush=True)
    if not subm:
        idx, shape = rb.outids, oshape
print("sum over the 20 tensor-core convs: %.2f ms" % total)
