#!/usr/bin/env python
"""Probe (2 GPUs, one process): how fast do OUR kernels read a table that lives on the other GPU?
Compares cudaMemcpyPeer, the LDG.128 gather and the cp.async interaction kernel on local vs peer tables."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommender_b200 import ops  # noqa: E402

assert torch.cuda.device_count() >= 2
d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
torch.cuda.set_device(0)
print("can_access_peer 0->1:", torch.cuda.can_device_access_peer(0, 1))
V, D, B, F = 4_000_000, 64, 65536, 26
g = torch.Generator(device=d0).manual_seed(1)
local = torch.empty(V, D, device=d0).uniform_(-0.05, 0.05, generator=g)
remote = torch.empty(V, D, device=d1).uniform_(-0.05, 0.05)
from recommender_b200._lib import check, lib  # noqa: E402
check(lib.rb_enable_peer_access(1), "rb_enable_peer_access")
idx = torch.randint(0, V, (B, F), device=d0, generator=g)
dense = torch.randn(B, D, device=d0, generator=g) * 0.1
out = torch.empty(B, 800, device=d0, dtype=torch.bfloat16)
res = {}


def timeit(name, fn, nbytes):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    res[name] = dict(ms=round(ms, 4), gbs=round(nbytes / ms / 1e6, 1))


buf = torch.empty(1 << 28, dtype=torch.uint8, device=d0)
src = torch.empty(1 << 28, dtype=torch.uint8, device=d1)
timeit("memcpy_peer_256MB", lambda: buf.copy_(src), 1 << 28)
rows_bytes = B * F * D * 4
timeit("gather_LDG_local", lambda: ops.gather_fwd(local, idx), rows_bytes)
timeit("gather_LDG_peer", lambda: ops.gather_fwd(remote, idx), rows_bytes)
timeit("interaction_cpasync_local", lambda: ops.dot_interaction_fwd(table=local, idx=idx, dense_vec=dense, tail=True, out=out, out_stride=800,
                                                                  out_dtype=torch.bfloat16), rows_bytes)
timeit("interaction_cpasync_peer", lambda: ops.dot_interaction_fwd(table=remote, idx=idx, dense_vec=dense, tail=True, out=out, out_stride=800,
                                                                 out_dtype=torch.bfloat16), rows_bytes)
print(json.dumps(res, indent=1))
