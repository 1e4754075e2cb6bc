"""Dev tool: bev_pool forward bandwidth, sorted (contiguous rows) vs perm (gathered rows)."""
import os, sys, statistics
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bevfusion_b200 import synthetic as S, _C
from bevfusion_b200.bev_pool import BEVPoolPlan, bev_pool_ext

dev = torch.device("cuda:0")
geom, cfg = S.camera_geometry("C2", device=dev)
plan = BEVPoolPlan(geom, cfg["xbound"], cfg["ybound"], cfg["zbound"])
t = plan.tables
x = S.lifted_features("C2", device=dev).reshape(-1, 80)
xs = x[t.perm[:t.n_kept].long()].contiguous()
B, D, H, W = t.dims
bytes_alg = 4 * 80 * t.n_kept + 4 * 80 * H * W + 4 * t.n_kept + 12 * t.n_intervals

def timeit(fn, n=20):
    for _ in range(3): fn()
    ev = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)

ms = timeit(lambda: plan.pool(x))
print("perm  : %.1f us  %.0f GB/s" % (ms * 1e3, bytes_alg / ms / 1e6))
ms = timeit(lambda: bev_pool_ext.bev_pool_forward(xs, t.geom, t.lengths, t.starts, B, D, H, W))
print("sorted: %.1f us  %.0f GB/s" % (ms * 1e3, bytes_alg / ms / 1e6))
# same instruction stream as the perm path (one bulk copy per row) but contiguous addresses
import ctypes
ident = torch.arange(t.n_kept, dtype=torch.int32, device=dev)
out = torch.empty((B, D, H, W, 80), device=dev)
ws = torch.empty(_C.lib().bevb200_bev_pool_workspace_bytes(t.n_kept, 80), dtype=torch.uint8, device=dev)
def ident_call():
    rc = _C.lib().bevb200_bev_pool_perm(B, D, H, W, t.n_kept, 80, t.n_intervals, _C.ptr(xs), _C.ptr(ident),
                                        _C.ptr(t.geom), _C.ptr(t.starts), _C.ptr(t.lengths), _C.ptr(out),
                                        _C.ptr(ws), ws.numel(), _C.current_stream(dev))
    assert rc == 0
ms = timeit(ident_call)
print("perm=identity (row copies, contiguous): %.1f us  %.0f GB/s" % (ms * 1e3, bytes_alg / ms / 1e6))
y = torch.empty_like(xs)
ms = timeit(lambda: y.copy_(xs))
print("copy 588MB r+w: %.1f us  %.0f GB/s" % (ms * 1e3, 2 * xs.numel() * 4 / ms / 1e6))
ms = timeit(lambda: xs.sum())
print("torch sum (read only): %.1f us  %.0f GB/s" % (ms * 1e3, xs.numel() * 4 / ms / 1e6))
g = torch.empty(B, D, H, W, 80, device=dev)
ms = timeit(lambda: g.zero_())
print("zero 41MB: %.1f us" % (ms * 1e3))
