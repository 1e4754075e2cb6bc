"""CPU, 2 gloo ranks: the multi-GPU part of bench.py is sample-parallel with no data-path
collective -- each rank runs its own frames, the timed region is bracketed by a barrier and the
reported time is the MAX over ranks.  This test drives that protocol (rank -> sample seed
assignment, barrier, max-reduce, whole-job throughput) on the CPU with the gloo backend."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, steps, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bevfusion_b200 import synthetic as S
        # every rank generates ITS OWN sample (seed = rank): no scatter of inputs is needed
        pts = S.lidar_cloud(seed=rank, sweeps=1)
        assert pts.shape[1] == 5
        dist.barrier()
        ms = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)   # rank-dependent "device time"
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        frames_per_s = world * steps * 1000.0 / float(ms.item())
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([pts.shape[0]], dtype=torch.int64))
        if rank == 0:
            out.put((float(ms.item()), frames_per_s, [int(s.item()) for s in sizes]))
    finally:
        dist.destroy_process_group()


def test_two_rank_sample_parallel_protocol():
    world, steps = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    ms, fps, sizes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 20.0                          # max over ranks, not mean / rank-0
    assert abs(fps - world * steps * 1000.0 / 20.0) < 1e-9
    assert len(sizes) == 2 and sizes[0] != sizes[1]   # different seeds -> different samples
