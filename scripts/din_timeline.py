#!/usr/bin/env python
"""Per-kernel times of one DIN attention step at BASELINE config 4's shape (CUPTI records through torch.profiler).

    python scripts/din_timeline.py [--batch 65536]
"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    a = ap.parse_args()
    from recommender_b200.din import DIN
    from recommender_b200.optimizers import Adam
    dev = torch.device("cuda", 0)
    B, L, Vi, Vc, D = a.batch, 100, 400_000, 2_000, 32
    g = torch.Generator(device=dev).manual_seed(4)
    din = DIN(Vi, D, Vc, D, device=dev, generator=g)
    lens = torch.randint(1, L + 1, (B, 1), device=dev, generator=g)
    hi = torch.randint(1, Vi, (B, L), device=dev, generator=g)
    hi = torch.where(torch.arange(L, device=dev)[None] < lens, hi, torch.zeros_like(hi))
    hc = torch.where(hi != 0, hi % (Vc - 1) + 1, torch.zeros_like(hi))
    ti = torch.randint(1, Vi, (B, 1), device=dev, generator=g)
    inp = dict(target_item=ti, target_cat=ti % (Vc - 1) + 1, pos_his_item=hi, pos_his_cat=hc)
    d_out = torch.randn(B, 4 * D, device=dev, generator=g) * 1e-3
    opt = Adam()

    def step():
        o = din(inp)
        o.backward(d_out)
        opt.apply_gradients(din)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    tot = collections.OrderedDict()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = e.name[:100]
            t = tot.setdefault(k, [0, 0.0])
            t[0] += 1
            t[1] += e.time_range.end - e.time_range.start
    s = sum(v[1] for v in tot.values())
    print(f"{'us':>10}  {'n':>3}  kernel      (sum {s:.0f} us)")
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{us:10.1f}  {n:3d}  {k}")


if __name__ == "__main__":
    main()
