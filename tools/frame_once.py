"""Dev tool for ncu launch lists: N eager hot-path frames (bench.HotPath.frame), nothing else.
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/launches_r2.csv python tools/frame_once.py 3
The LAST frame's launches are the ones to read (tools/traffic_digest.py does)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
hp = bench.HotPath(dev, seed=0)
x, pts = hp.device_inputs(seed=0)
for i in range(n):
    torch.cuda.synchronize()
    print("FRAME", i, flush=True)
    hp.frame(x, pts)
torch.cuda.synchronize()
