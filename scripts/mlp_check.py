#!/usr/bin/env python
"""Bring-up and timing of the tcgen05 Dense kernels (csrc/mlp.cu) against torch on the same bf16 operands.

    python scripts/mlp_check.py [--time] [--only fwd,bwd_input,bwd_weight,head]

Prints one line per case: max |err| relative to max |ref|, and where the worst element sits.  Exit code 1 if any case
is off by more than the bf16 output rounding."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommender_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, generator=g, device=dev) * scale).to(torch.bfloat16)


def report(name, got, ref, tol):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-30
    worst = err.max().item() / scale
    idx = int(err.argmax().item())
    r, c = divmod(idx, ref.shape[-1]) if ref.dim() == 2 else (idx, 0)
    bad = int((err > tol * scale).sum().item())
    ok = worst <= tol and not torch.isnan(got).any().item()
    print(f"{'ok ' if ok else 'BAD'} {name}: max rel err {worst:.3e} at ({r},{c}) got {got.reshape(-1)[idx].item():.5f} ref {ref.reshape(-1)[idx].item():.5f}; "
          f"{bad} of {err.numel()} beyond tol", flush=True)
    if not ok and ref.dim() == 2:
        rows_bad = (err > tol * scale).any(1).nonzero().reshape(-1)[:8].tolist()
        cols_bad = (err > tol * scale).any(0).nonzero().reshape(-1)[:16].tolist()
        print(f"    first bad rows {rows_bad} cols {cols_bad}", flush=True)
    return ok


def check_fwd(rows, in_dim, units, act=None, out_dtype=torch.bfloat16, bias=True):
    x, w = rnd(rows, in_dim), rnd(in_dim, units, scale=in_dim ** -0.5)
    b = torch.randn(units, generator=g, device=dev) if bias else None
    y = ops.dense_fwd(x, w, b, act, out_dtype)
    ref = x.float() @ w.float() + (b if bias else 0)
    if act == "relu":
        ref = ref.relu()
    elif act == "sigmoid":
        ref = ref.sigmoid()
    return report(f"fwd rows={rows} in={in_dim} units={units} act={act} out={str(out_dtype)[6:]}", y, ref, 1e-2 if out_dtype == torch.bfloat16 else 1e-4)


def check_bwd_input(rows, in_dim, units):
    dy, w = rnd(rows, units), rnd(in_dim, units, scale=units ** -0.5)
    dx = ops.dense_bwd_input(dy, w)
    return report(f"bwd_input rows={rows} in={in_dim} units={units}", dx, dy.float() @ w.float().t(), 1e-2)


def check_bwd_weight(rows, in_dim, units):
    x, dy = rnd(rows, in_dim), rnd(rows, units, scale=rows ** -0.5)
    dw = ops.dense_bwd_weight(x, dy)
    return report(f"bwd_weight rows={rows} in={in_dim} units={units}", dw, x.float().t() @ dy.float(), 2e-4)


def check_head(rows, in_dim, act):
    x, w = rnd(rows, in_dim), rnd(in_dim, 1, scale=in_dim ** -0.5)
    b = torch.randn(1, generator=g, device=dev)
    out = ops.dense_head_fwd(x, w.reshape(-1), b, act)
    z = (x.float() @ w.float()).reshape(-1) + b
    ref = z.sigmoid() if act == "sigmoid" else z.relu() if act == "relu" else z
    ok = report(f"head_fwd rows={rows} in={in_dim} act={act}", out[:, None], ref[:, None], 1e-5)
    dout = torch.randn(rows, generator=g, device=dev)
    dx, dw, db = ops.dense_head_bwd(dout, out, act, x, w.reshape(-1))
    dz = dout * (ref * (1 - ref) if act == "sigmoid" else (ref > 0).float() if act == "relu" else 1.0)
    ok &= report("head_bwd dx", dx, dz[:, None] * w.float().reshape(1, -1), 1e-2)
    ok &= report("head_bwd dW", dw[:, None], (x.float().t() @ dz)[:, None], 1e-4)
    ok &= report("head_bwd db", db[:, None], dz.sum().reshape(1, 1), 1e-4)
    return ok


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3     # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--only", default="fwd,bwd_input,bwd_weight,head")
    ap.add_argument("--quick", action="store_true", help="the two smallest cases of each product only")
    ap.add_argument("--stats", action="store_true", help="per-role wait cycles of the three products at 65536 x 800 x 512")
    ap.add_argument("--profile", action="store_true", help="no checks: 3 launches of each product at 65536 x 800 x 512 (for ncu)")
    a = ap.parse_args()
    only = a.only.split(",")
    ok = True
    if a.stats:
        from recommender_b200._lib import lib
        B, i, u = 65536, 800, 512
        x, w, dy = rnd(B, i), rnd(i, u, scale=i ** -0.5), rnd(B, u)
        bias = torch.zeros(u, device=dev)
        buf = torch.zeros(8 * 148, dtype=torch.int64, device=dev)
        names = ["mma_wait_operands", "mma_wait_accumulator", "mma_total", "producer_wait_stage", "producer_total", "epilogue_wait_acc", "epilogue_total"]
        for label, fn in (("fwd", lambda: ops.dense_fwd(x, w, bias, None)), ("bwd_input", lambda: ops.dense_bwd_input(dy, w)),
                          ("bwd_weight", lambda: ops.dense_bwd_weight(x, dy))):
            fn()
            torch.cuda.synchronize()
            buf.zero_()
            lib.rb_dense_debug_stats(buf.data_ptr())
            fn()
            torch.cuda.synchronize()
            lib.rb_dense_debug_stats(None)
            t = buf.reshape(148, 8).double()
            lead = t[t[:, 2] > 0]
            print(label, {n: int(lead[:, k].mean().item()) if lead.numel() else 0 for k, n in enumerate(names)},
                  "epilogue(all CTAs):", {n: int(t[:, k].mean().item()) for k, n in enumerate(names) if k >= 5}, flush=True)
        return
    if a.profile:
        B, i, u = 65536, 800, 512
        x, w, dy = rnd(B, i), rnd(i, u, scale=i ** -0.5), rnd(B, u)
        bias = torch.zeros(u, device=dev)
        for _ in range(3):
            ops.dense_fwd(x, w, bias, None)
            ops.dense_bwd_input(dy, w)
            ops.dense_bwd_weight(x, dy)
        torch.cuda.synchronize()
        print("profiled")
        return
    if "fwd" in only:
        cases = [(128, 64, 64), (256, 128, 256), (300, 800, 512), (1000, 16, 512), (128, 512, 256), (4096, 512, 256)]
        for rows, i, u in cases[:2] if a.quick else cases:
            ok &= check_fwd(rows, i, u)
        if not a.quick:
            ok &= check_fwd(77, 256, 64, "relu", torch.float32)
            ok &= check_fwd(513, 32, 32, "sigmoid", torch.float32)
            ok &= check_fwd(130, 24, 40, None, torch.bfloat16, bias=False)
    if "bwd_input" in only:
        cases = [(128, 64, 64), (256, 256, 128), (300, 800, 512), (1000, 512, 256), (4096, 256, 64), (77, 16, 512)]
        for rows, i, u in cases[:2] if a.quick else cases:
            ok &= check_bwd_input(rows, i, u)
    if "bwd_weight" in only:
        cases = [(128, 128, 64), (256, 128, 256), (300, 800, 512), (1000, 16, 512), (4096, 512, 256), (65536, 256, 64), (100, 40, 24)]
        for rows, i, u in cases[:2] if a.quick else cases:
            ok &= check_bwd_weight(rows, i, u)
    if "head" in only:
        ok &= check_head(1000, 256, "sigmoid")
        ok &= check_head(77, 32, None)
    print("ALL OK" if ok else "FAILURES", flush=True)

    if a.time:
        B = 65536
        res = {}
        for name, (i, u) in {"top1 800->512": (800, 512), "top2 512->256": (512, 256), "bot1 16->512": (16, 512), "bot3 256->64": (256, 64)}.items():
            x, w, dy = rnd(B, i), rnd(i, u, scale=i ** -0.5), rnd(B, u)
            bias = torch.zeros(u, device=dev)
            bias16 = bias.to(torch.bfloat16)
            y = torch.empty(B, u, dtype=torch.bfloat16, device=dev)
            dx = torch.empty(B, i, dtype=torch.bfloat16, device=dev)
            dw = torch.empty(i, u, dtype=torch.float32, device=dev)
            flops = 2.0 * B * i * u
            t = {
                "fwd": timeit(lambda: ops.dense_fwd(x, w, bias, None, out=y)),
                "fwd_cublas": timeit(lambda: torch.addmm(bias16, x, w, out=y)),
                "bwd_input": timeit(lambda: ops.dense_bwd_input(dy, w, out=dx)),
                "bwd_input_cublas": timeit(lambda: torch.mm(dy, w.t(), out=dx)),
                "bwd_weight": timeit(lambda: ops.dense_bwd_weight(x, dy, out=dw)),
                "bwd_weight_cublas": timeit(lambda: torch.mm(x.t(), dy, out_dtype=torch.float32)),
            }
            res[name] = {k: dict(us=round(v, 1), tflops=round(flops / v / 1e6, 1)) for k, v in t.items()}
            print(name, json.dumps(res[name]), flush=True)
        print(json.dumps(res))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
