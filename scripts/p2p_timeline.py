#!/usr/bin/env python
"""Kernel timeline of one replayed step of the peer-memory sharded DLRM on rank 0 (torchrun, >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/p2p_timeline.py [--replicate-small 4096]
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicate-small", type=int, default=0)
    ap.add_argument("--batch", type=int, default=65536)
    a = ap.parse_args()
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from recommender_b200.graph import GraphedTrainStep
    from recommender_b200.model import bce_clipped
    from recommender_b200.optimizers import Adam
    from recommender_b200.p2p import P2PShardedDLRM
    D, B = 64, a.batch
    gen = torch.Generator(device=dev).manual_seed(4)
    model = P2PShardedDLRM(bench.BOTTOM[:-1] + [D], bench.TOP, D, 1_000_000, 26, 13, num_tables=26, device=dev, compute_dtype=torch.bfloat16,
                           generator=gen, table_rows=bench.CRITEO_TB_ROWS, capacity_factor=2.0, replicate_rows_upto=a.replicate_small)
    host = bench.synth_batches(4, B, 1_000_000, "uniform", seed=4 + rank, pin=False, table_rows=bench.CRITEO_TB_ROWS)
    res = [tuple(t.to(dev) for t in b) for b in host]
    gs = GraphedTrainStep(model, Adam(), bce_clipped, res[0], warmup=3)
    for i in range(5):
        gs.step(res[i % 4])
    torch.cuda.synchronize()
    dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(3):
            gs.step(res[i % 4])
        torch.cuda.synchronize()
    if rank == 0:
        evs = []
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                evs.append((e.time_range.start, e.time_range.end - e.time_range.start, e.name[:95]))
        evs.sort()
        # the last replay: cut at the largest gaps
        starts = [i for i in range(1, len(evs)) if evs[i][0] - (evs[i - 1][0] + evs[i - 1][1]) > 150]
        last = evs[starts[-1]:] if starts else evs
        t0 = last[0][0]
        print(f"{'start':>8} {'dur':>8}  name")
        for s, d, n in last:
            print(f"{s - t0:8.1f} {d:8.1f}  {n}")
        print(f"span {last[-1][0] + last[-1][1] - t0:.1f} us, {len(last)} kernels")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
