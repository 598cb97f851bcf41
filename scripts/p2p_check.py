#!/usr/bin/env python
"""Multi-process check of the peer-memory sharded path (run under torchrun on >= 2 GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/p2p_check.py

Every rank trains a P2PShardedDLRM for a few steps on its own batches; rank 0 also runs the same global
problem UNSHARDED (plain DLRM, gradients of the G local losses summed) and compares the reassembled table."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def init_params(seed, bottom, top, D, rows, num_int=13, num_cat=26):
    """Keras-default initial weights (table U(-0.05, 0.05), Dense Glorot-uniform, zero bias) from one seeded generator,
    identical on every rank."""
    rng = np.random.default_rng(seed)

    def mlp(in_dim, units):
        layers = []
        for u in units:
            lim = np.sqrt(6.0 / (in_dim + u))
            layers.append((rng.uniform(-lim, lim, size=(in_dim, u)).astype(np.float32), np.zeros(u, np.float32)))
            in_dim = u
        return layers

    return dict(table=rng.uniform(-0.05, 0.05, size=(rows, D)).astype(np.float32), bottom=mlp(num_int, bottom),
                top=mlp((num_cat + 1) ** 2 + D, top))


def synth_batch(B, V, seed, num_cat=26, num_int=13):
    """Zipf-like ids with 2 % forced id 0 (hot OOV row), log1p dense features, ~25 % positives."""
    rng = np.random.default_rng(seed)
    cat = (rng.pareto(1.05, size=(B, num_cat)) * 3).astype(np.int64) % V
    cat[rng.random((B, num_cat)) < 0.02] = 0
    dense = np.log1p(rng.integers(0, 1000, size=(B, num_int))).astype(np.float32)
    label = (rng.random(B) < 0.25).astype(np.int64)
    return cat, dense, label


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from recommender_b200.model import DLRM, bce_clipped
    from recommender_b200.optimizers import Adam
    from recommender_b200.p2p import P2PShardedDLRM
    V, D, B, T, steps = 5000, 32, 512, 26, 3
    params = init_params(4, [64, D], [64, 1], D, V * T)
    model = P2PShardedDLRM([64, D], [64, 1], D, V, 26, 13, num_tables=T, device=dev)
    model.embedding_layer.load_full_table(torch.tensor(params["table"]))
    model.bottom_mlp.load_arrays(params["bottom"], dev)
    model.top_mlp.load_arrays(params["top"], dev)
    opt = Adam()
    batches = [[synth_batch(B, V, seed=100 * s + r) for r in range(world)] for s in range(steps)]
    for s in range(steps):
        cat, dense_x, label = (torch.tensor(a, device=dev) for a in batches[s][rank])
        loss = bce_clipped(model({"cat_features": cat, "int_features": dense_x}), label)
        loss.backward()
        opt.apply_gradients(model)
    model.embedding_layer.check_overflow()
    torch.cuda.synchronize()
    # gather the shards on rank 0 as (unsharded row id, row) pairs
    emb = model.embedding_layer
    max_rows = (emb.total_rows + world - 1) // world
    ids = torch.full((max_rows,), -1, dtype=torch.int64, device=dev)
    rows = torch.zeros(max_rows, D, device=dev)
    ids[: emb.local_rows] = emb.full_row_ids()
    rows[: emb.local_rows] = emb.embeddings
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    all_rows = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(all_ids, ids)
    dist.all_gather(all_rows, rows)
    ok = True
    if rank == 0:
        full = torch.empty(V * T, D, device=dev)
        for k in range(world):
            okk = all_ids[k] >= 0
            full[all_ids[k][okk]] = all_rows[k][okk]
        ref = DLRM([64, D], [64, 1], D, V, 26, 13, num_tables=T, device=dev)
        ref.embedding_layer.embeddings.copy_(torch.tensor(params["table"]))
        ref.bottom_mlp.load_arrays(params["bottom"], dev)
        ref.top_mlp.load_arrays(params["top"], dev)
        ropt = Adam()
        for s in range(steps):
            total = 0
            for r in range(world):        # MirroredStrategy with Reduction.NONE: the replicas' gradients are summed
                cat, dense_x, label = (torch.tensor(a, device=dev) for a in batches[s][r])
                total = total + bce_clipped(ref({"cat_features": cat, "int_features": dense_x}), label)
            total.backward()
            ropt.apply_gradients(ref)
        torch.cuda.synchronize()
        got, want = full.cpu().numpy(), ref.embedding_layer.embeddings.cpu().numpy()
        moved = np.abs(want - params["table"]) > 0
        err = np.abs(got - want).max()
        print(f"p2p_check: world={world} rows moved={int(moved.any(1).sum())} max|sharded - unsharded|={err:.3e}")
        ok = bool(err <= 2e-6 and moved.any())
        print("p2p_check: OK" if ok else "p2p_check: MISMATCH")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
