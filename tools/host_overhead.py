"""Dev tool: host-side (Python + launch) time of one SparseEncoder pass vs its GPU time."""
import os, sys, time, statistics, cProfile, pstats, io
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
dev = torch.device("cuda:0")
hp = bench.HotPath(dev)
x, pts = hp.device_inputs()
from bevfusion_b200.voxelize import voxelize_mean
v, c, n = hp.voxelize(pts)
feats, coords = voxelize_mean(v, c, n, 0)
with torch.no_grad():
    for _ in range(5):
        hp.encoder(feats, coords, 1)
    torch.cuda.synchronize()
    host, gpu = [], []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); e0.record()
        hp.encoder(feats, coords, 1)
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        host.append((t1 - t0) * 1e3); gpu.append(e0.elapsed_time(e1))
    print("encoder: host %.3f ms (call returns), gpu %.3f ms" % (statistics.median(host), statistics.median(gpu)))
    pr = cProfile.Profile(); pr.enable()
    for _ in range(10):
        hp.encoder(feats, coords, 1)
    pr.disable(); torch.cuda.synchronize()
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(28)
    print(st.getvalue()[:6000])
