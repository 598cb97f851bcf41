#!/usr/bin/env python
"""Kernel micro-benchmark at BASELINE config 2 shapes (B=65536, F=26, D=64, 26 x 1M rows): times each
C-ABI call with CUDA events on the launching stream, inputs larger than L2 (6.7 GB tables, a ring of
batches).  Used for tuning; `RB_LIB_PATH` selects a variant build (recommender_b200/build.py --variant).

    python scripts/kbench.py [--ops fwd,bwd,update,gather] [--iters 20] [--tables 26] [--dist uniform|zipf]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommender_b200 import ops  # noqa: E402
from recommender_b200.ops import GradSource, LookupGroup  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="fwd,bwd,update")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--tables", type=int, default=26)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--dist", default="uniform")
    ap.add_argument("--optimizer", default="adam_lazy")
    ap.add_argument("--row-cache", default="l2", choices=["l2", "l1"], help="how the fused lookups copy table rows (rb_row_cache)")
    ap.add_argument("--tag", default=os.environ.get("RB_LIB_PATH", "default"))
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, F, D, V, T = a.batch, 26, a.dim, a.rows, a.tables
    g = torch.Generator(device=dev).manual_seed(4)
    table = torch.empty(V * T, D, device=dev).uniform_(-0.05, 0.05, generator=g)
    m, v = torch.zeros_like(table), torch.zeros_like(table)
    off = torch.arange(T, device=dev, dtype=torch.int64) * V if T > 1 else None
    ring = []
    for i in range(4):
        if a.dist == "uniform":
            cat = torch.randint(0, V, (B, F), device=dev, generator=g)
        else:
            u = torch.rand(B, F, device=dev, dtype=torch.float64, generator=g).clamp_(min=1e-12)
            cat = (u.pow(-1.0 / 0.05).clamp_(max=2.0 ** 62).to(torch.int64) % V)
            cat[torch.rand(B, F, device=dev, generator=g) < 0.02] = 0
        ring.append(cat)
    dense = torch.randn(B, D, device=dev, generator=g) * 0.1
    width = 27 * 27 + D
    stride = (width + 7) // 8 * 8
    dout = (torch.randn(B, stride, device=dev, generator=g) * 1e-3).to(torch.bfloat16)
    dE = torch.randn(B, F, D, device=dev, generator=g) * 1e-3
    out = torch.empty(B, stride, device=dev, dtype=torch.bfloat16)
    res = {}

    def timeit(name, fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        evs = []
        for i in range(a.iters):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(i)
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        ts = sorted(s.elapsed_time(e) * 1e3 for s, e in evs)
        res[name] = dict(us_median=ts[len(ts) // 2], us_min=ts[0], us_mean=sum(ts) / len(ts))

    which = a.ops.split(",")
    rc = a.row_cache == "l1"
    if "fwd" in which:
        timeit("fwd", lambda i: ops.dot_interaction_fwd(table=table, idx=ring[i % 4], field_row_offset=off, dense_vec=dense, tail=True,
                                                        out=out, out_stride=stride, out_dtype=torch.bfloat16, row_cache=rc))
    if "bwd" in which:
        timeit("bwd", lambda i: ops.dot_interaction_bwd(dout, table=table, idx=ring[i % 4], field_row_offset=off, dense_vec=dense, tail=True,
                                                        row_cache=rc))
    if "gather" in which:
        timeit("gather", lambda i: ops.gather_fwd(table, ring[i % 4], L=F, field_row_offset=off))
    if "pool" in which:      # BASELINE config 4: masked mean over a behaviour history, L = 100, D = 32, item table of 400k rows
        Lh, Dh, Vh = 100, 32, 400_000
        wt = torch.empty(Vh, Dh, device=dev).uniform_(-0.05, 0.05, generator=g)
        lens = torch.randint(1, Lh + 1, (B, 1), device=dev, generator=g)
        hist = torch.randint(1, Vh, (B, Lh), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
        hist = torch.where(torch.arange(Lh, device=dev)[None] < lens, hist, torch.zeros_like(hist))     # trailing pads (dien/data_loader.py:44)
        pout = torch.empty(B, Dh, device=dev)
        timeit("pool_masked_mean", lambda i: ops.bag_pool_fwd(wt, hist, "masked_mean", out=pout, out_stride=Dh))
        valid = int((hist != 0).sum())
        res["pool_masked_mean"]["algorithmic_bytes"] = valid * Dh * 4 + B * Lh * 4 + B * Dh * 4
        dpool = torch.randn(B, Dh, device=dev, generator=g) * 1e-3
        cnt = (hist != 0).sum(1).float()
        mh, vh = torch.zeros_like(wt), torch.zeros_like(wt)
        hstep = [0]

        def hupd(i):
            hstep[0] += 1
            grp = LookupGroup(hist, Lh, GradSource.per_bag([dpool], scale="masked_mean", mask_idx=hist, count=cnt))
            ops.sparse_bwd_update(wt, mh, vh, [grp], optimizer="adam_lazy", step=hstep[0])
        timeit("pool_masked_mean_bwd_update", hupd)
        res["pool_masked_mean_bwd_update"]["pairs"] = int(valid + (hist[:, :1] == 0).sum()
                                                          + ((hist[:, 1:] == 0) & (hist[:, :-1] != 0)).sum())   # one pad per run survives
        hws = ops.sparse_workspace(B * Lh, Dh, Vh, dev)

        def hupd_all(i):       # the two-phase form sorts ids only and keeps every pad pair: the "before" of the collapse
            hstep[0] += 1
            grp = LookupGroup(hist, Lh, GradSource.per_bag([dpool], scale="masked_mean", mask_idx=hist, count=cnt))
            sel = ops.sparse_bwd_prepare(Vh, Dh, [grp], hws)
            ops.sparse_bwd_apply(wt, mh, vh, [grp], hws, sel, optimizer="adam_lazy", step=hstep[0])
        timeit("pool_masked_mean_bwd_update_all_pairs", hupd_all)
        res["pool_masked_mean_bwd_update_all_pairs"]["pairs"] = B * Lh
    if "fm" in which:        # DeepFM front end (ctr/model.py:19-23) at D = 16, one shared 1M-row table
        wf = torch.empty(1_000_000, 16, device=dev).uniform_(-0.05, 0.05, generator=g)
        catf = torch.randint(0, 1_000_000, (B, F), device=dev, generator=g)
        timeit("gather_fm_fwd", lambda i: ops.gather_fm_fwd(wf, catf))
        res["gather_fm_fwd"]["algorithmic_bytes"] = B * F * 16 * 4 * 2 + B * F * 8 + B * 16 * 4 + B * 4
    if "update" in which:
        step = [0]

        def upd(i):
            step[0] += 1
            grp = LookupGroup(ring[i % 4], F, GradSource.per_position(dE, F), field_row_offset=off)
            ops.sparse_bwd_update(table, m, v if a.optimizer.startswith("adam") else None, [grp], optimizer=a.optimizer, step=step[0])
        timeit("update", upd)
    if "bwd_update" in which:     # the fused backward + row update (rows touched once) and the reduction over the remaining rows
        ws = ops.sparse_workspace(B * F, D, V * T, dev)
        single = torch.empty(B * F, dtype=torch.uint8, device=dev)
        dEb = torch.empty(B, F, D, device=dev)
        step = [0]
        evs, evs2 = [], []
        for i in range(a.iters + 3):
            step[0] += 1
            idx = ring[i % 4]
            sel = ops.sparse_bwd_prepare(V * T, D, [LookupGroup(idx, F, None, field_row_offset=off)], ws)
            sel = ops.sparse_bwd_mark_singletons(V * T, D, B * F, ws, sel, single)
            s_, e_, e2_ = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            s_.record()
            ops.dot_interaction_bwd_update(dout, table=table, idx=idx, single=single, state0=m, state1=v, field_row_offset=off,
                                           dense_vec=dense, tail=True, optimizer=a.optimizer, step=step[0], dE=dEb)
            e_.record()
            grp = LookupGroup(idx, F, GradSource.per_position(dEb, F), field_row_offset=off)
            ops.sparse_bwd_apply(table, m, v, [grp], ws, sel, optimizer=a.optimizer, step=step[0], skip_singletons=True)
            e2_.record()
            if i >= 3:
                evs.append((s_, e_))
                evs2.append((e_, e2_))
        torch.cuda.synchronize()
        for name, ev in (("bwd_update", evs), ("apply_rest", evs2)):
            ts = sorted(x.elapsed_time(y) * 1e3 for x, y in ev)
            res[name] = dict(us_median=ts[len(ts) // 2], us_min=ts[0], us_mean=sum(ts) / len(ts))
        res["bwd_update"]["singletons"] = int(single.sum())
    if "apply" in which:     # the segmented reduction + row update alone (keys + sort outside the timed region, as in the step)
        ws = ops.sparse_workspace(B * F, D, V * T, dev)
        step = [0]
        evs = []
        for i in range(a.iters + 3):
            step[0] += 1
            grp = LookupGroup(ring[i % 4], F, GradSource.per_position(dE, F), field_row_offset=off)
            sel = ops.sparse_bwd_prepare(V * T, D, [grp], ws)
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            ops.sparse_bwd_apply(table, m, v, [grp], ws, sel, optimizer=a.optimizer, step=step[0])
            e_.record()
            if i >= 3:
                evs.append((s_, e_))
        torch.cuda.synchronize()
        ts = sorted(s_.elapsed_time(e_) * 1e3 for s_, e_ in evs)
        res["apply"] = dict(us_median=ts[len(ts) // 2], us_min=ts[0], us_mean=sum(ts) / len(ts))
    for v_ in res.values():
        if "algorithmic_bytes" in v_:
            v_["gbs"] = round(v_["algorithmic_bytes"] / v_["us_median"] / 1e3, 1)
    print(json.dumps(dict(tag=a.tag, tables=T, dist=a.dist, row_cache=a.row_cache, **res)))


if __name__ == "__main__":
    main()
