"""Dev tool: fused lift + pool, column kernels vs the round-1 row kernel vs pooling the materialised volume
(CUDA-graph replays, CUDA events).  python tools/lift_bench.py [C2|C5]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
from bevfusion_b200 import synthetic as S
from bevfusion_b200.bev_pool import BEVPoolPlan
geom, c = S.camera_geometry(cfg, device=dev)
plan = BEVPoolPlan(geom, c["xbound"], c["ybound"], c["zbound"])
B, N, D, fH, fW, _ = geom.shape
del geom
g = torch.Generator(device=dev).manual_seed(0)
depth = torch.softmax(torch.randn(B, N, D, fH, fW, generator=g, device=dev), dim=2).contiguous()
ctx = torch.randn(B, N, fH, fW, c["C"], generator=g, device=dev)
out = plan.lift_pool(depth, ctx)
tabs = plan._lift_cache[1]
print(cfg, "kept", plan.tables.n_kept, "intervals", plan.tables.n_intervals, "segments", tabs[5],
      "mean pixels per segment %.1f" % (plan.tables.n_kept / max(tabs[5], 1)))
print("columns (graph) %.1f us" % (1e3 * bench.graph_time_ms(dev, lambda: plan.lift_pool(depth, ctx))))
os.environ["BEVB200_LIFT_VARIANT"] = "rows"
print("rows    (graph) %.1f us" % (1e3 * bench.graph_time_ms(dev, lambda: plan.lift_pool(depth, ctx))))
del os.environ["BEVB200_LIFT_VARIANT"]
if cfg == "C2":
    x = depth.unsqueeze(-1) * ctx.unsqueeze(2)
    print("pool of the materialised volume (graph) %.1f us" % (1e3 * bench.graph_time_ms(dev, lambda: plan.pool(x))))
for _ in range(2):
    plan.lift_pool(depth, ctx)
torch.cuda.synchronize()
