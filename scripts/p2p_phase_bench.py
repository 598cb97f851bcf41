#!/usr/bin/env python
"""Phase timings of the peer-memory sharded step across REAL ranks (torchrun, one process per GPU), phases
separated by barriers; plus the raw bandwidth of our gather kernel on an IPC-mapped peer shard."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _Raw:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = dict(shape=shape, typestr="<f4", data=(ptr, False), version=3, strides=None)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from recommender_b200 import ops
    from recommender_b200.p2p import DistPeerLink, P2PShardedEmbedding
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    V, T, D, F = 1_000_000, 26, 64, 26
    emb = P2PShardedEmbedding(V, D, num_tables=T, link=DistPeerLink(None, dev), device=dev)
    g = torch.Generator(device=dev).manual_seed(3 + rank)
    idx = torch.randint(0, V, (B, F), device=dev, generator=g)
    dense = torch.randn(B, D, device=dev, generator=g) * 0.1
    dOut = (torch.randn(B, 800, device=dev, generator=g) * 1e-3).to(torch.bfloat16)
    flags = (False, True, True)
    times = {}

    def phase(name, fn):
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times.setdefault(name, []).append(s.elapsed_time(e) * 1e3)

    for it in range(4):
        phase("route", lambda: emb.route(idx))
        phase("collect_and_sort", lambda: emb.collect_and_sort())
        phase("fwd", lambda: emb._interaction_fwd(idx, dense, flags, torch.bfloat16, 8))
        phase("bwd", lambda: emb._interaction_bwd(idx, dense, flags, dOut))
        emb._routed_by_caller = True
        phase("apply", lambda: emb.apply_pending("adam_lazy", it + 1, 1e-3))
        phase("barrier_allreduce", lambda: emb.link.barrier())
    emb.check_overflow()
    # raw: our LDG gather over an IPC-mapped peer shard vs the local shard
    peer = (rank + 1) % world
    ptrs = emb._resolve(emb._shard_ptrs)
    rows = emb.local_rows - 8
    if world == 1:
        os._exit(0)
    peer_t = torch.as_tensor(_Raw(ptrs[peer], (rows, D)), device=dev)
    lidx = torch.randint(0, rows, (B, F), device=dev, generator=g)
    phase("gather_peer_ipc", lambda: ops.gather_fwd(peer_t, lidx))
    phase("gather_peer_ipc", lambda: ops.gather_fwd(peer_t, lidx))
    phase("gather_local", lambda: ops.gather_fwd(emb.embeddings, lidx))
    phase("gather_local", lambda: ops.gather_fwd(emb.embeddings, lidx))
    out = {k: round(v[-1], 1) for k, v in times.items()}
    out.update(rank=rank, world=world, B_local=B, n_valid=int(emb._n_valid.item()))
    for r in range(world):
        if r == rank:
            print(json.dumps(out), flush=True)
        dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
