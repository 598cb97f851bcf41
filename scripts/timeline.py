#!/usr/bin/env python
"""Kernel timeline of one replayed DLRM training step (BASELINE config 2) from CUPTI activity records (torch.profiler):
which kernel ran when, on which stream, and where the step's critical path has gaps.  ncu serialises kernels and
therefore cannot show overlap; nsys is not installed.

    python scripts/timeline.py [--out gpurun_out/timeline.json] [--mlp-backend tcgen05|cublas] [--dist uniform|zipf]
"""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/timeline.json")
    ap.add_argument("--mlp-backend", default="tcgen05")
    ap.add_argument("--dist", default="uniform")
    ap.add_argument("--tables", type=int, default=26)
    ap.add_argument("--batch", type=int, default=65536)
    a = ap.parse_args()
    from recommender_b200.graph import GraphedTrainStep
    from recommender_b200.model import DLRM, bce_clipped
    from recommender_b200.optimizers import Adam
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    D, V, T, B = 64, 1_000_000, a.tables, a.batch
    gen = torch.Generator(device=dev).manual_seed(4)
    model = DLRM(bench.BOTTOM[:-1] + [D], bench.TOP, D, V, 26, 13, num_tables=T, device=dev, compute_dtype=torch.bfloat16, generator=gen)
    for m in model.modules():
        if hasattr(m, "backend"):
            m.backend = a.mlp_backend
    opt = Adam()
    host = bench.synth_batches(4, B, V, a.dist, seed=4, pin=False)
    res = [tuple(t.to(dev) for t in b) for b in host]
    gs = GraphedTrainStep(model, opt, bce_clipped, res[0], warmup=3)
    for i in range(5):
        gs.step(res[i % 4])
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(3):
            gs.step(res[i % 4])
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            evs.append(dict(name=e.name[:90], start=e.time_range.start, dur=e.time_range.end - e.time_range.start,
                            stream=getattr(e, "stream", None) if hasattr(e, "stream") else None))
    evs.sort(key=lambda r: r["start"])
    # keep the last replay: split on the largest gaps
    if not evs:
        print("no CUDA events recorded")
        return
    n = len(evs) // 3
    step = evs[-n:]
    t0 = step[0]["start"]
    for r in step:
        r["start"] = round(r["start"] - t0, 2)
        r["dur"] = round(r["dur"], 2)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(step, open(a.out, "w"), indent=0)
    end = 0.0
    print(f"{'start':>9} {'dur':>8} {'gap':>7}  stream  name")
    for r in step:
        gap = r["start"] - end
        print(f"{r['start']:9.1f} {r['dur']:8.1f} {gap:7.1f}  {str(r['stream']):>6}  {r['name']}")
        end = max(end, r["start"] + r["dur"])
    print(f"step span {end:.1f} us, {len(step)} kernels")


if __name__ == "__main__":
    main()
