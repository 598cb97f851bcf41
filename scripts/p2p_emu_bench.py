#!/usr/bin/env python
"""Phase timings of the peer-memory sharded step with G emulated ranks on ONE GPU (p2p.LocalPeerLink) at
config-2 sizes: isolates kernel-side cost from NVLink effects.   python scripts/p2p_emu_bench.py [G]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommender_b200.p2p import LocalPeerLink, P2PShardedEmbedding  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
V, T, D, F = 1_000_000, 26, 64, 26
dev = torch.device("cuda", 0)
reg = {}
embs = [P2PShardedEmbedding(V, D, num_tables=T, link=LocalPeerLink(G, r, reg), device=dev) for r in range(G)]
g = torch.Generator(device=dev).manual_seed(3)
idx = [torch.randint(0, V, (B, F), device=dev, generator=g) for _ in range(G)]
dense = [torch.randn(B, D, device=dev, generator=g) * 0.1 for _ in range(G)]
dOut = [(torch.randn(B, 800, device=dev, generator=g) * 1e-3).to(torch.bfloat16) for _ in range(G)]
flags = (False, True, True)
times = {}


def phase(name, fn):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    times.setdefault(name, []).append(s.elapsed_time(e) * 1e3)


for it in range(4):
    for r in range(G):
        phase("route", lambda: embs[r].route(idx[r]))
    for r in range(G):
        phase("collect_and_sort", lambda: embs[r].collect_and_sort())
    for r in range(G):
        phase("fwd", lambda: embs[r]._interaction_fwd(idx[r], dense[r], flags, torch.bfloat16, 8))
    for r in range(G):
        phase("bwd", lambda: embs[r]._interaction_bwd(idx[r], dense[r], flags, dOut[r]))
    for r in range(G):
        embs[r]._routed_by_caller = True
        phase("apply", lambda: embs[r].apply_pending("adam_lazy", it + 1, 1e-3))
    for r in range(G):
        embs[r].check_overflow()
print(json.dumps({k: round(sorted(v[G:])[len(v[G:]) // 2], 1) for k, v in times.items()} | {"G": G, "B_local": B,
                 "n_valid": [int(e._n_valid.item()) for e in embs], "capacity": embs[0].capacity}))
