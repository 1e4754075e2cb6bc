"""Dev tool: main-stream timeline of one SparseEncoder pass (gap before / duration of every conv launch)."""
import os, sys, statistics
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from bevfusion_b200.spconv import ops as sp_ops
dev = torch.device("cuda:0")
hp = bench.HotPath(dev)
x, pts = hp.device_inputs()
from bevfusion_b200.voxelize import voxelize_mean
v, c, n = hp.voxelize(pts)
feats, coords = voxelize_mean(v, c, n, 0)
real_conv, real_dense = sp_ops.sparse_conv, sp_ops.sparse_to_dense
rows = []
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
def conv(*a, **k):
    e0 = ev(); out = real_conv(*a, **k); e1 = ev()
    cur.append(("conv %dx%d n=%d" % (a[0].shape[1], a[1].shape[-1], out.shape[0]), e0, e1)); return out
def dense(*a, **k):
    e0 = ev(); out = real_dense(*a, **k); e1 = ev(); cur.append(("dense", e0, e1)); return out
with torch.no_grad():
    for _ in range(5):
        hp.encoder(feats, coords, 1)
    sp_ops.sparse_conv, sp_ops.sparse_to_dense = conv, dense
    runs = []
    for _ in range(8):
        cur = []
        torch.cuda.synchronize()
        s = ev(); hp.encoder(feats, coords, 1); e = ev()
        torch.cuda.synchronize()
        runs.append((s, e, cur))
names = [r[0] for r in runs[0][2]]
tot = statistics.median(s.elapsed_time(e) for s, e, _ in runs)
print("encoder total %.3f ms" % tot)
gsum = dsum = 0.0
for i, name in enumerate(names):
    gaps, durs = [], []
    for s, e, cur in runs:
        prev_end = s if i == 0 else cur[i - 1][2]
        gaps.append(prev_end.elapsed_time(cur[i][1]) * 1e3); durs.append(cur[i][1].elapsed_time(cur[i][2]) * 1e3)
    g, d = statistics.median(gaps), statistics.median(durs)
    gsum += g; dsum += d
    print("%2d %-24s gap before %7.1f us   duration %7.1f us" % (i, name, g, d))
tail = statistics.median(cur[-1][2].elapsed_time(e) * 1e3 for s, e, cur in runs)
print("sum of gaps %.1f us, sum of durations %.1f us, tail %.1f us" % (gsum, dsum, tail))
