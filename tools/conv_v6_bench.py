"""Dev tool: generation-6 sparse-conv kernel (split operand images) against generation 5 on the real C3
rulebooks: max relative difference, split-image consistency, CUDA-event times per layer.
    BEVB200_SPCONV_TC_VARIANT=5 python tools/conv_v6_bench.py [small]
(the env makes ops.sparse_conv run generation 5; generation 6 is called through its own entry point)."""
import os, sys, statistics
os.environ.setdefault("BEVB200_SPCONV_TC_VARIANT", "5")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from bevfusion_b200 import _C, synthetic as S
from bevfusion_b200.spconv import ops
from bevfusion_b200.voxelize import Voxelization, voxelize_mean

dev = torch.device("cuda:0")
small = "small" in sys.argv
L = S.LIDAR_C3
if small:
    rng = np.random.default_rng(0)
    shape = [48, 40, 41]
    flat = rng.choice(shape[0] * shape[1] * shape[2], size=5000, replace=False)
    coords = torch.from_numpy(np.stack([np.zeros_like(flat), flat // (shape[1] * shape[2]),
                                        (flat // shape[2]) % shape[1], flat % shape[2]], 1).astype(np.int32)).to(dev)
else:
    pts = torch.from_numpy(S.lidar_cloud(seed=0)).to(dev)
    vox = Voxelization(L["voxel_size"], L["point_cloud_range"], L["max_num_points"], L["max_voxels"]).eval()
    v, c, n = vox(pts)
    _, coords = voxelize_mean(v, c, n, 0)
    shape = L["sparse_shape"]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ev = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)


lib = _C.lib()


def split_rows(f):
    n, c = f.shape
    ce = lib.bevb200_spconv_split_channels(c)
    out = torch.empty((n, ce * 4), dtype=torch.uint8, device=dev)
    _C.check(lib.bevb200_spconv_split_rows(_C.ptr(f), n, 0, c, _C.ptr(out), _C.current_stream(dev)), "split_rows")
    return out, ce


def pack6(w, cin, cout, kv):
    nb = lib.bevb200_spconv_split_weight_bytes(cin, cout, kv)
    pk = torch.empty(nb, dtype=torch.uint8, device=dev)
    _C.check(lib.bevb200_spconv_pack_split_weights(_C.ptr(w), cin, cout, kv, _C.ptr(pk), _C.current_stream(dev)), "pack")
    return pk


def conv6(fs, ce, pk, rb, n_in, cout, kv, scale, shift, res, relu, n_dev=None, want_split=True):
    out = torch.empty((rb.n_out, cout), dtype=torch.float32, device=dev)
    osp = torch.empty((rb.n_out, cout * 4), dtype=torch.uint8, device=dev) if want_split else None
    rc = lib.bevb200_spconv_forward_split(_C.ptr(fs), _C.ptr(pk), _C.ptr(rb.nbr), rb.n_out, n_in, rb.n_out,
                                          _C.ptr(n_dev), ce, cout, kv, _C.ptr(scale), _C.ptr(shift), _C.ptr(res),
                                          int(relu), _C.ptr(out), _C.ptr(osp), _C.current_stream(dev))
    _C.check(rc, "forward_split")
    return out, osp


def decode_split(osp, c):
    """split image -> fp32 (hi + lo)"""
    n = osp.shape[0]
    w = osp.view(torch.int16).view(n, c // 16, 2, 16)          # [row, group, hi|lo, 16 bf16]
    f = (w.to(torch.int32) << 16).view(torch.float32)
    return (f[:, :, 0, :] + f[:, :, 1, :]).reshape(n, c)


idx = coords
layers = [("in", 5, 16, True, 3, 1, 1), ("s1 subm", 16, 16, True, 3, 1, 1), ("s1 down", 16, 32, False, 3, 2, 1),
          ("s2 subm", 32, 32, True, 3, 1, 1), ("s2 down", 32, 64, False, 3, 2, 1),
          ("s3 subm", 64, 64, True, 3, 1, 1), ("s3 down", 64, 128, False, 3, 2, [1, 1, 0]),
          ("s4 subm", 128, 128, True, 3, 1, 1), ("out", 128, 128, False, [1, 1, 3], [1, 1, 2], 0)]
tot5 = tot6 = 0.0
for name, cin, cout, subm, ks, st, pd in layers:
    rb, oshape = ops.get_rulebook(idx, 1, shape, ks, st, pd, 1, 0, subm)
    n_in = idx.shape[0]
    f = torch.randn(n_in, cin, device=dev)
    kv = rb.nbr.shape[0]
    w = torch.randn(kv, cin, cout, device=dev) / (cin * 5)
    scale = torch.rand(cout, device=dev) + 0.5
    shift = torch.randn(cout, device=dev) * 0.1
    res = torch.randn(rb.n_out, cout, device=dev) * 0.1
    packed5 = ops.pack_weights(w, 3)
    o5 = ops.sparse_conv(f, w, rb.nbr, rb.n_out, scale, shift, res, True, precision=3, packed=packed5)
    fs, ce = split_rows(f)
    pk = pack6(w, cin, cout, kv)
    o6, osp = conv6(fs, ce, pk, rb, n_in, cout, kv, scale, shift, res, True)
    torch.cuda.synchronize()
    den = o5.abs().max().item() + 1e-30
    err = (o6 - o5).abs().max().item() / den
    dec = decode_split(osp, cout)
    err_s = (dec - o6).abs().max().item() / den
    # device-side row count: only the first 3/4 of the rows are produced
    n_part = (rb.n_out * 3) // 4
    nd = torch.tensor([n_part], dtype=torch.int32, device=dev)
    o6p, _ = conv6(fs, ce, pk, rb, n_in, cout, kv, scale, shift, res, True, n_dev=nd, want_split=False)
    torch.cuda.synchronize()
    err_p = (o6p[:n_part] - o5[:n_part]).abs().max().item() / den if n_part else 0.0
    t5 = timeit(lambda: ops.sparse_conv(f, w, rb.nbr, rb.n_out, scale, shift, res, True, precision=3, packed=packed5))
    t6 = timeit(lambda: conv6(fs, ce, pk, rb, n_in, cout, kv, scale, shift, res, True))
    t6n = timeit(lambda: conv6(fs, ce, pk, rb, n_in, cout, kv, scale, shift, res, True, want_split=False))
    pairs = int((rb.nbr >= 0).sum())
    gf = 2.0 * pairs * cin * cout / 1e9
    mult = 4 if (subm and cin > 5) else 1
    tot5 += t5 * mult
    tot6 += t6 * mult
    print(f"{name:8s} n_out {rb.n_out:7d} {cin:4d}->{cout:4d} pairs {pairs:8d}  v5 {t5*1e3:7.1f} us  v6 {t6*1e3:7.1f} us "
          f"(fp32 only {t6n*1e3:7.1f})  {gf/t6:7.1f} TF/s  err {err:.2e} split {err_s:.2e} part {err_p:.2e}  (x{mult})",
          flush=True)
    if not subm:
        idx, shape = rb.outids, oshape
print("sum over the 21 convs: v5 %.3f ms   v6 %.3f ms" % (tot5, tot6))
