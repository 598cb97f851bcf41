#!/usr/bin/env python
"""DeepFM training step (ctr/model.py:6-31 + Adam) on one GPU: BASELINE config 1 shape (B = 1024, D = 16, one shared
1M-row table, MLP 512-256-1) and the same model at B = 65536.  Eager and as one CUDA graph.

    python scripts/deepfm_bench.py [--batches 1024,65536] [--steps 30]
"""
import argparse
import json
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommender_b200.graph import GraphedTrainStep  # noqa: E402
from recommender_b200.model import DeepFM, bce_logits  # noqa: E402
from recommender_b200.optimizers import Adam  # noqa: E402


class Logits(nn.Module):
    """Keras evaluates DeepFM's loss on the logits (its last op is Sigmoid)."""

    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, inputs):
        return self.m.logits(inputs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1024,65536")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--mlp-dtype", default="bf16")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    V, D = 1_000_000, 16
    out = {}
    for B in [int(x) for x in a.batches.split(",")]:
        g = torch.Generator(device=dev).manual_seed(4)
        model = Logits(DeepFM(D, V, 13, 26, [512, 256, 1], device=dev, generator=g,
                              compute_dtype=torch.bfloat16 if a.mlp_dtype == "bf16" else None))
        opt = Adam()
        ring = [(torch.randint(0, V, (B, 26), device=dev, generator=g),
                 torch.log1p(torch.randint(0, 1000, (B, 13), device=dev, generator=g).float()),
                 (torch.rand(B, device=dev, generator=g) < 0.25).float()) for _ in range(4)]

        def eager(b):
            loss = bce_logits(model({"cat_features": b[0], "int_features": b[1]}), b[2])
            loss.backward()
            opt.apply_gradients(model)
            return loss.detach()

        def timed(fn):
            for i in range(5):
                fn(ring[i % 4])
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for i in range(a.steps):
                loss = fn(ring[i % 4])
            e.record()
            torch.cuda.synchronize()
            return s.elapsed_time(e) / a.steps, float(loss)

        ms_e, _ = timed(eager)
        graphed = GraphedTrainStep(model, opt, bce_logits, ring[0])
        ms_g, loss = timed(graphed.step)
        out[f"B={B}"] = dict(eager_ms=round(ms_e, 4), graph_ms=round(ms_g, 4), samples_per_s=round(B / ms_g * 1e3), loss=round(loss, 5))
    print(json.dumps(dict(model="DeepFM", emb_dim=D, rows=V, mlp=[512, 256, 1], mlp_dtype=a.mlp_dtype, **out)))


if __name__ == "__main__":
    main()
