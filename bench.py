#!/usr/bin/env python
"""bench.py -- frames/sec of the BEVFusion C+L hot path on B200 (one JSON line on stdout).

A "step" is one frame of the hot path named by BASELINE.json's north_star, on synthetic
nuScenes-shaped inputs (bevfusion_b200/synthetic.py):

    bev_pool forward   6 cam x 118 depth x 32 x 88 frustum, C=80 -> 360x360 BEV   (config C2)
    hard_voxelize      ~296 k points x 5, 0.075 m voxels, grid 1440x1440x40        (config C3)
    voxel mean + SparseEncoder (VoxelNet 0.075: 17 SubM + 4 strided sparse convs) -> [256,180,180]

The dense glue networks of the full model (SwinT, FPNs, fuser, SECOND, TransFusion head) are not
part of the hot path and are not run; `config.workload` says so.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); every rank runs the same per-frame work on its
own synthetic sample (weak scaling, no data-path collective -- the path shards by sample), the
timed region is bracketed by barrier + cuda synchronize, time = max over ranks.

--impl reference times the reference's CPU implementation of the path on the host cores
(oracle/_ref extension for the sparse encoder, the restated QuickCumsum for bev_pool, the C port
for hard_voxelize), each step a bounded sample of the frame (see `sample`).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "frames/sec C+L BEVFusion hot path (bev_pool + hard_voxelize + SparseEncoder)"
WORKLOAD = ("C2+C3 hot path per frame: bev_pool fwd 6-cam 256x704 D=118 C=80 -> 360x360 (N'=1,993,728 rows, "
            "638 MB fp32) + hard_voxelize ~296k pts 0.075 m (1440x1440x40, cap 160k x 10) + voxel mean + "
            "SparseEncoder VoxelNet-0.075 (17 SubM + 4 strided convs) -> [1,256,180,180]; "
            "dense glue nets (SwinT/FPN/fuser/SECOND/TransFusion) not included")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# dram__bytes_read.sum + dram__bytes_write.sum summed over the 21 sparse-conv launches of one frame, from
# the committed ncu pass (profiles/r1_launches_v5.md); not measured in this process
ENC_TRAFFIC_BYTES = 1.4428e9

# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
class HotPath:
    """One frame of the hot path through the repo's public API (bevfusion_b200.*)."""

    def __init__(self, device, seed=0, precision=None):
        from bevfusion_b200 import synthetic as S
        from bevfusion_b200.bev_pool import BEVPoolPlan
        from bevfusion_b200.sparse_encoder import voxelnet_0p075_encoder
        from bevfusion_b200.voxelize import Voxelization
        self.S, self.device = S, device
        geom, cfg = S.camera_geometry("C2", device=device)
        self.cfg = cfg
        self.plan = BEVPoolPlan(geom, cfg["xbound"], cfg["ybound"], cfg["zbound"])   # static per calibration
        del geom
        L = S.LIDAR_C3
        self.voxelize = Voxelization(L["voxel_size"], L["point_cloud_range"], L["max_num_points"],
                                     L["max_voxels"]).eval()
        torch.manual_seed(seed)
        self.encoder = voxelnet_0p075_encoder().to(device).eval()
        self.precision = precision
        self.points_host = torch.from_numpy(S.lidar_cloud(seed=seed)).pin_memory()
        self.x_host = None
        t = self.plan.tables
        self.n_kept, self.n_intervals, self.n_total = t.n_kept, t.n_intervals, t.n_total
        self.ev = {}

    def device_inputs(self, seed=0):
        x = self.S.lifted_features("C2", device=self.device, seed=seed)          # 638 MB, > L2
        return x, self.points_host.to(self.device)

    def host_inputs(self, seed=0):
        if self.x_host is None:
            g = torch.Generator().manual_seed(seed)
            shape = (1, 6, 118, 32, 88, 80)
            self.x_host = torch.empty(shape, dtype=torch.float32).pin_memory()
            # fill blockwise (cheap): one camera of randn repeated with a per-camera offset
            block = torch.randn(shape[2:], generator=g)
            for cam in range(6):
                self.x_host[0, cam].copy_(block + 0.01 * cam)
        return self.x_host, self.points_host

    def frame(self, x, points, timers=None):
        """x [1,6,118,32,88,80] and points [N,5] on the device -> (bev [1,80,360,360], lidar [1,256,180,180])."""
        from bevfusion_b200.voxelize import voxelize_mean

        def mark(name):
            if timers is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                timers.setdefault(name, []).append(e)

        mark("t0")
        bev = self.plan(x)
        mark("bev_pool")
        v, c, n = self.voxelize(points)
        feats, coords = voxelize_mean(v, c, n, 0)
        mark("voxelize")
        with torch.no_grad():
            lidar = self.encoder(feats, coords, 1, precision=self.precision)
        mark("encoder")
        return bev, lidar

    # algorithmic work per frame (DESIGN.md section "roofline accounting")
    def bev_pool_bytes(self):
        C = 80
        return 4 * C * self.n_kept + 4 * C * 360 * 360 + 4 * self.n_kept + 12 * self.n_intervals

    def encoder_flops(self, feats, coords):
        """sum over convs of 2 * pairs * Cin * Cout, pairs counted from the rulebooks."""
        from bevfusion_b200 import spconv
        from bevfusion_b200.spconv.conv import SparseConvolution
        total, pairs_total = 0, 0
        x = spconv.SparseConvTensor(feats, coords.int(), self.encoder.sparse_shape, 1)
        convs = []
        hooks = [m.register_forward_pre_hook(lambda mod, inp: convs.append((mod, inp[0])))
                 for m in self.encoder.modules() if isinstance(m, SparseConvolution)]
        with torch.no_grad():
            self.encoder(feats, coords, 1, precision=self.precision)
        for h in hooks:
            h.remove()
        for mod, inp in convs:
            rb, _ = mod._rulebook(inp)
            pairs = int((rb.nbr >= 0).sum())
            pairs_total += pairs
            total += 2 * pairs * mod.in_channels * mod.out_channels
        return total, pairs_total


def conv_launch_times(hp, feats, coords, frames=5):
    """CUDA-event time of every sparse-conv kernel launch of `frames` encoder passes (the events
    bracket the library call on the launching stream).  Returns (ms per frame summed over the conv
    launches, launches per frame)."""
    from bevfusion_b200.spconv import ops as sp_ops
    real = sp_ops.sparse_conv
    evs = []

    def timed(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = real(*a, **k)
        e1.record()
        evs.append((e0, e1))
        return out

    with torch.no_grad():
        hp.encoder(feats, coords, 1, precision=hp.precision)          # warm
        sp_ops.sparse_conv = timed
        try:
            for _ in range(frames):
                hp.encoder(feats, coords, 1, precision=hp.precision)
        finally:
            sp_ops.sparse_conv = real
    torch.cuda.synchronize()
    total = sum(a.elapsed_time(b) for a, b in evs)
    return total / frames, len(evs) // frames


def run_ours(args, rank, world, local_rank):
    from bevfusion_b200 import _C
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    _C.lib()
    peaks = load_peaks()
    hp = HotPath(device, seed=rank, precision=args.precision)
    x, pts = hp.device_inputs(seed=rank)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = hp.frame(x, pts)
    del out
    barrier()
    # --- device-resident throughput ------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:          # one nvidia-smi poller per job (its driver queries can stall CUDA calls)
        sampler.start()
    timers = {}
    _C.reset_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = hp.frame(x, pts, timers)
    e1.record()
    barrier()
    launches = _C.launch_count()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps

    def stage_ms(a, b):
        return statistics.mean(ea.elapsed_time(eb) for ea, eb in zip(timers[a], timers[b]))

    stages = dict(bev_pool_ms=stage_ms("t0", "bev_pool"), voxelize_ms=stage_ms("bev_pool", "voxelize"),
                  encoder_ms=stage_ms("voxelize", "encoder"))

    # --- end to end through the public API with HOST buffers ------------------------------
    xh, ph = hp.host_inputs(seed=rank)
    bev_h = torch.empty((1, 80, 360, 360), dtype=torch.float32).pin_memory()
    lid_h = torch.empty((1, 256, 180, 180), dtype=torch.float32).pin_memory()
    del x
    torch.cuda.empty_cache()

    # double-buffered: the H2D copy of frame i+1 (copy stream) overlaps the compute of frame i
    copy_stream = torch.cuda.Stream(device=device)
    main_stream = torch.cuda.current_stream(device)
    bufs = [(torch.empty(xh.shape, dtype=xh.dtype, device=device), torch.empty(ph.shape, dtype=ph.dtype, device=device))
            for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def stage_in(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i % 2])
            bufs[i % 2][0].copy_(xh, non_blocking=True)
            bufs[i % 2][1].copy_(ph, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_run(nframes):
        for f in freed:
            f.record(main_stream)
        stage_in(0)
        for i in range(nframes):
            if i + 1 < nframes:
                stage_in(i + 1)
            main_stream.wait_event(ready[i % 2])
            bev, lidar = hp.frame(bufs[i % 2][0], bufs[i % 2][1])
            freed[i % 2].record(main_stream)
            bev_h.copy_(bev, non_blocking=True)       # D2H of the step's results
            lid_h.copy_(lidar, non_blocking=True)

    e2e_run(2)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    e0.record()
    e2e_run(e2e_steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    h2d = xh.numel() * 4 + ph.numel() * 4
    d2h = bev_h.numel() * 4 + lid_h.numel() * 4

    if rank != 0:
        return
    # --- roofline of the dominant kernel + the north star's named kernel (bev_pool) --------
    x, pts = hp.device_inputs(seed=0)
    v, c, n = hp.voxelize(pts)
    from bevfusion_b200.voxelize import voxelize_mean
    feats, coords = voxelize_mean(v, c, n, 0)
    flops, pairs = hp.encoder_flops(feats, coords)
    pool_bytes = hp.bev_pool_bytes()
    # bev_pool alone, CUDA events, inputs (638 / 588 MB) larger than L2:
    #   plan path   = interval-cell kernel + bevpool_fwd_tma_kernel<20,PERM> (gather + zero-fill fused) + fix-up
    #   drop-in op  = memset + bevpool_fwd_tma_kernel<20,SORTED> + fix-up on already sorted rows (the reference contract)
    def time_us(fn, n=20):
        for _ in range(3):
            fn()
        evs = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return statistics.median(a.elapsed_time(b) for a, b in evs)

    from bevfusion_b200.bev_pool import bev_pool_ext
    t = hp.plan.tables
    pool_ms = time_us(lambda: hp.plan.pool(x))
    xs = x.reshape(-1, 80)[t.perm[:t.n_kept].long()].contiguous()
    Bq, Dq, Hq, Wq = t.dims
    op_ms = time_us(lambda: bev_pool_ext.bev_pool_forward(xs, t.geom, t.lengths, t.starts, Bq, Dq, Hq, Wq))
    del xs
    # bev_pool backward (plan path: grads written in the caller's row order, dropped rows zeroed)
    from bevfusion_b200.bev_pool import _PoolPerm
    og = torch.randn(Bq, Dq, Hq, Wq, 80, device=device)

    class _Ctx:
        tables, c = t, 80
    bwd_ms = time_us(lambda: _PoolPerm.backward(_Ctx, og))
    bwd_bytes = 4 * 80 * t.n_intervals + 4 * 80 * t.n_total + 4 * t.n_total
    # fused LSS lift + pool: depth [1,6,118,32,88] (x) ctx [1,6,32,88,80], no 638 MB volume
    depth = torch.softmax(torch.randn(1, 6, 118, 32, 88, device=device), dim=2).contiguous()
    ctx = torch.randn(1, 6, 32, 88, 80, device=device)
    lift_ms = time_us(lambda: hp.plan.lift_pool(depth, ctx))
    del og, depth, ctx
    # SURVEY.md section 8(f) rows built after the path itself (timed alone, CUDA events, median of 20):
    from bevfusion_b200 import synthetic as S_
    from bevfusion_b200.scatter_points import dynamic_scatter
    from bevfusion_b200.voxelize import voxel_layer, voxelize_mean_fused
    from bevfusion_b200.vtransform import points_to_depth
    L_ = S_.LIDAR_C3
    fused_vox_ms = time_us(lambda: voxelize_mean_fused(pts, L_["voxel_size"], L_["point_cloud_range"], 10, 160000, 0))
    unfused_vox_ms = time_us(lambda: voxelize_mean(*hp.voxelize(pts), 0))
    dcoors = torch.zeros(pts.shape[0], 3, dtype=torch.int32, device=device)
    voxel_layer.dynamic_voxelize(pts, dcoors, L_["voxel_size"], L_["point_cloud_range"], 3)
    scatter_ms = time_us(lambda: dynamic_scatter(pts, dcoors, "mean"))
    M_ = S_.lidar_camera_matrices(6, (256, 704), batch=1)
    margs = (M_["lidar2image"].to(device), M_["img_aug_matrix"].to(device), M_["lidar_aug_matrix"].to(device), (256, 704))
    depth_ms = time_us(lambda: points_to_depth([pts], *margs))
    next_rows = {"voxelize_mean_fused_ms": round(fused_vox_ms, 4), "voxelize_then_mean_ms": round(unfused_vox_ms, 4),
                 "dynamic_scatter_mean_ms": round(scatter_ms, 4), "lidar_depth_images_6x256x704_ms": round(depth_ms, 4),
                 "points": int(pts.shape[0]),
                 "note": "each includes its host-side result-size readback (.item()) where the API returns sized tensors"}
    pool_gbs = pool_bytes / (pool_ms * 1e-3) / 1e9
    op_gbs = pool_bytes / (op_ms * 1e-3) / 1e9
    enc_tflops = flops / (stages["encoder_ms"] * 1e-3) / 1e12
    # `traffic`: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed
    # ncu pass (profiles/r1_launches_v5.md), not measured in this process
    roof_pool = dict(kernel="bevpool_fwd_tma_kernel<20,PERM> (plan API: gather through perm + zero-fill fused; + cells, fix-up)",
                     bound="hbm", achieved=round(pool_gbs, 1), peak=peaks["hbm_gbs"], unit="GB/s",
                     frac=round(pool_gbs / peaks["hbm_gbs"], 4), traffic=677.3e6, ms=round(pool_ms, 4),
                     algorithmic_bytes=pool_bytes, peak_source=peaks["source"])
    roof_pool_op = dict(kernel="bevpool_fwd_tma_kernel<20,SORTED> (drop-in bev_pool_forward on sorted rows; + memset, fix-up)",
                        bound="hbm", achieved=round(op_gbs, 1), peak=peaks["hbm_gbs"], unit="GB/s",
                        frac=round(op_gbs / peaks["hbm_gbs"], 4), traffic=None, ms=round(op_ms, 4),
                        algorithmic_bytes=pool_bytes, peak_source=peaks["source"])
    roof_enc = dict(kernel="spconv_tc_kernel_v5<3> x21 (whole SparseEncoder incl. rulebooks on the side stream, dense)",
                    bound="tensor", achieved=round(enc_tflops, 3), peak=peaks["bf16_tflops_sustained"], unit="TFLOP/s",
                    frac=round(enc_tflops / peaks["bf16_tflops_sustained"], 5), traffic=ENC_TRAFFIC_BYTES,
                    ms=round(stages["encoder_ms"], 4), algorithmic_flops=flops, pairs=pairs,
                    peak_source=peaks["source"],
                    note="useful FLOPs = sum 2*pairs*Cin*Cout over real (non-missing) neighbour pairs; the kernel issues "
                         "3 (2 when the hi|lo weight images are merged) bf16 MMAs per fp32 product (BF16x3 split) and "
                         "also multiplies the zero rows of missing neighbours; traffic = sum over the 21 tensor-core "
                         "convs of dram__bytes_read+write in profiles/r1_launches_v5.md")
    # hard_voxelize + mean: algorithmic bytes 4*F*N (points) + M*(4*P*F + 16) (voxels, coors, num) -- latency bound
    n_pts, m_vox = int(pts.shape[0]), int(v.shape[0])
    vox_bytes = 4 * 5 * n_pts + m_vox * (4 * 10 * 5 + 16)
    vox_gbs = vox_bytes / (stages["voxelize_ms"] * 1e-3) / 1e9
    roof_vox = dict(kernel="hard_voxelize (5 kernels + scan) + voxel_mean", bound="hbm", achieved=round(vox_gbs, 1),
                    peak=peaks["hbm_gbs"], unit="GB/s", frac=round(vox_gbs / peaks["hbm_gbs"], 4), traffic=None,
                    ms=round(stages["voxelize_ms"], 4), algorithmic_bytes=vox_bytes,
                    points_per_s=round(n_pts / (stages["voxelize_ms"] * 1e-3)), peak_source=peaks["source"],
                    note="latency bound: 40 MB of algorithmic traffic in ~10 dependent launches")
    # dominant kernel of the step: the tcgen05 sparse conv (21 launches per frame, ~3/4 of the step);
    # its launches are timed one by one with CUDA events, achieved = useful FLOPs of those launches / that time
    conv_ms, conv_launches = conv_launch_times(hp, feats, coords)
    conv_tflops = flops / (conv_ms * 1e-3) / 1e12
    roof_conv = dict(kernel="spconv_tc_kernel_v5<3> (tcgen05 implicit-GEMM sparse conv, BF16x3; %d launches per frame)" % conv_launches,
                     bound="tensor", achieved=round(conv_tflops, 3), peak=peaks["bf16_tflops_sustained"], unit="TFLOP/s",
                     frac=round(conv_tflops / peaks["bf16_tflops_sustained"], 5), traffic=ENC_TRAFFIC_BYTES,
                     ms=round(conv_ms, 4), avg_launch_us=round(1e3 * conv_ms / max(conv_launches, 1), 2),
                     launches_per_frame=conv_launches, algorithmic_flops=flops, pairs=pairs, peak_source=peaks["source"],
                     note="achieved = sum over the frame's conv launches of 2*pairs*Cin*Cout (real neighbour pairs only) / "
                          "their summed CUDA-event time; the kernel issues 2-3 bf16 MMAs per fp32 product (hi/lo split) and "
                          "also multiplies the zero rows of missing neighbours, so the tensor pipe is busier than this "
                          "fraction says (ncu: 36 % active at C=64, 64 % at C=128; profiles/r1_ncu_full_v5.md); traffic = "
                          "dram bytes of those launches (profiles/r1_launches_v5.md)")
    dominant = roof_conv if stages["encoder_ms"] >= stages["bev_pool_ms"] else roof_pool
    # the CPU baseline is timed on rank 0 at N = 1 only
    cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline(n_steps=1)
    line = {
        "metric": METRIC, "value": round(world * 1000.0 / ms_per_step, 3), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": 1, "parallelism": "sample-parallel x%d" % world,
                   "spconv_precision": {None: "bf16x3 (tcgen05 kind::f16, bf16 hi/lo split of fp32 operands, fp32 accumulate; default)",
                                        0: "fp32 (SIMT)", 1: "tf32x3", 2: "tf32", 3: "bf16x3"}[args.precision],
                   "l2": "inputs larger than L2: the 638 MB feature volume streams through L2 every step",
                   "bev_pool_plan": "rank/sort/interval tables cached per calibration (static geometry)",
                   "kept_rows": hp.n_kept, "intervals": hp.n_intervals},
        "stages_ms": {k: round(v, 4) for k, v in stages.items()},
        "e2e": {"value": round(world * 1000.0 / e2e_ms, 3), "unit": "frames/s", "ms_per_step": round(e2e_ms, 3),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "bev_pool_extra": {"backward_ms": round(bwd_ms, 4), "backward_GBs": round(bwd_bytes / (bwd_ms * 1e-3) / 1e9, 1),
                           "backward_frac_of_hbm_peak": round(bwd_bytes / (bwd_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                           "fused_lift_pool_ms": round(lift_ms, 4),
                           "note": "backward = bevpool_bwd_kernel through perm (660 MB algorithmic); fused lift+pool reads "
                                   "depth (8 MB) + L2-resident ctx (5.4 MB) instead of the 638 MB lifted volume"},
        "next_rows": next_rows,
        "roofline": dominant, "roofline_bev_pool": roof_pool, "roofline_bev_pool_op": roof_pool_op,
        "roofline_encoder": roof_enc, "roofline_voxelize": roof_vox,
        "cpu_baseline": cpu, "clocks": clocks,
    }
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Route fd 1 to stderr while the job runs (NCCL prints its version banner to stdout) so that
    stdout carries exactly ONE line: the JSON emitted by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's CPU path on the host cores
# ---------------------------------------------------------------------------------------------
CPU_SAMPLE = ("per step: bev_pool CPU path (torch QuickCumsum restatement, bev_pool.py:9-35 + base.py:149-169) on "
              "camera 0 of 6 (x6), hard_voxelize C port on the full cloud, reference CPU spconv extension "
              "(oracle/_ref) SparseEncoder on a 45-degree azimuth wedge of the cloud scaled by the voxel ratio")


class CpuFrame:
    def __init__(self, seed=0):
        import oracle
        from oracle import reference_pipeline as RP
        from oracle.build_ref import built, load_ref
        from bevfusion_b200 import synthetic as S
        from bevfusion_b200.bev_pool import gen_dx_bx
        from bevfusion_b200.sparse_encoder import voxelnet_0p075_encoder
        self.oracle, self.RP, self.S = oracle, RP, S
        # the reference CPU path is small GEMMs + serial gather/scatter: it stops scaling (and then
        # regresses) beyond a few tens of threads, so use at most 16 of the host cores
        self.threads = min(os.cpu_count() or 1, 16)
        torch.set_num_threads(self.threads)
        self.kind = "reference" if built("sparse_conv_ext_ref") else "port"
        self.ref = load_ref("sparse_conv_ext_ref") if self.kind == "reference" else None
        geom, cfg = S.camera_geometry("C2")
        self.geom0 = geom[:, :1].contiguous()
        self.dx, self.bx, self.nx = gen_dx_bx(cfg["xbound"], cfg["ybound"], cfg["zbound"])
        g = torch.Generator().manual_seed(seed)
        self.x0 = torch.randn((1, 1, 118, 32, 88, 80), generator=g)
        self.points = S.lidar_cloud(seed=seed)
        az = np.arctan2(self.points[:, 1], self.points[:, 0])
        self.wedge = self.points[(az >= 0) & (az < np.pi / 4)]
        torch.manual_seed(seed)
        self.encoder = voxelnet_0p075_encoder().eval()

    def step(self):
        """returns the extrapolated CPU seconds for one full frame"""
        RP, L = self.RP, self.S.LIDAR_C3
        t0 = time.perf_counter()
        RP.bev_pool_cpu_quickcumsum(self.x0, self.geom0, self.dx, self.bx, self.nx)
        t_pool = (time.perf_counter() - t0) * 6.0
        t0 = time.perf_counter()
        feats, coords = RP.voxelize_cpu(self.points, L, 160000)
        t_vox = time.perf_counter() - t0
        wf, wc = RP.voxelize_cpu(self.wedge, L, 160000)
        t0 = time.perf_counter()
        with torch.no_grad():
            if self.ref is not None:
                RP.reference_encoder_forward(self.ref, self.encoder, wf, wc, 1)
            else:
                self._port_encoder(wf, wc)
        t_enc = (time.perf_counter() - t0) * (feats.shape[0] / max(wf.shape[0], 1))
        return t_pool + t_vox + t_enc

    def _port_encoder(self, feats, coords):
        # oracle port of the first stage only, scaled by the FLOP share (used when oracle/_ref is absent)
        o = self.oracle
        w = self.encoder.conv_input[0].weight.detach().numpy()
        o.sparse_conv(feats.numpy(), coords.numpy(), 1, [1440, 1440, 41], w, [3, 3, 3], [1, 1, 1], [1, 1, 1],
                      [1, 1, 1], True, acc64=False)


def cpu_baseline(n_steps=1):
    cf = CpuFrame()
    secs = [cf.step() for _ in range(n_steps)]
    s = statistics.median(secs)
    return {"value": round(1.0 / s, 5), "unit": "frames/s", "cores": cf.threads, "kind": cf.kind,
            "sample": CPU_SAMPLE, "seconds_per_frame": round(s, 3)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cf = CpuFrame()
    for _ in range(min(args.warmup, 1)):
        cf.step()
    t0 = time.perf_counter()
    secs = [cf.step() for _ in range(args.steps)]
    wall = (time.perf_counter() - t0) / max(args.steps, 1)
    s = sum(secs) / len(secs)
    value = round(1.0 / s, 5)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": round(s * 1000.0, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU host cores only; one process regardless of n_gpus; "
                       "ms_per_step is the frame time extrapolated from the per-step sample",
                       "sample_wall_ms_per_step": round(wall * 1000.0, 1)},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cf.threads, "kind": cf.kind,
                             "sample": CPU_SAMPLE},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", type=int, default=None, help="spconv precision: 0 fp32, 1 tf32x3, 2 tf32, 3 bf16x3 (default)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
