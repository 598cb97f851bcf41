"""Real-rank self-check of the peer-memory sharded DLRM: every rank trains `p2p.P2PShardedDLRM` for a few steps on its
own batches while rank 0 also trains the same global problem UNSHARDED (`model.DLRM`, the gradients of the G local
losses summed — MirroredStrategy with Reduction.NONE), then compares the reassembled table
and the last step's probabilities.  This is the check that exercises what single-GPU emulation cannot: the device-side
symmetric-memory barriers and the kernels' reads of rows / gradient rows in PEER memory over NVLink.

Used by `bench.py` (N > 1: the `parity` object of the JSON line), `scripts/p2p_check.py` (torchrun) and the `-m gpu`
test that spawns two ranks when two GPUs are visible.  Product code checking product code: no oracle involved."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist


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


def run(dev: torch.device, *, compute_dtype: Optional[torch.dtype] = None, V: int = 5000, D: int = 32, B: int = 512, T: int = 26,
        steps: int = 3, group=None, table_rows=None, replicate_rows_upto: int = 0) -> Optional[Dict[str, float]]:
    """Collective over `group`.  Returns {'max_abs_table_diff', 'max_abs_prob_diff', 'rows_moved', 'steps', ...} on rank 0,
    None elsewhere."""
    from .model import DLRM, bce_clipped
    from .optimizers import Adam
    from .p2p import P2PShardedDLRM
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bottom, top = [64, D], [64, 1]
    rows_list = [V] * T if table_rows is None else [int(r) for r in table_rows]
    total = sum(rows_list)
    cards = np.array(rows_list, dtype=np.int64)[None]
    params = init_params(4, bottom, top, D, total)
    model = P2PShardedDLRM(bottom, top, D, V, 26, 13, num_tables=T, device=dev, compute_dtype=compute_dtype, group=group,
                           table_rows=table_rows, replicate_rows_upto=replicate_rows_upto,
                           capacity_factor=2.0 if table_rows is not None else 1.25)
    model.embedding_layer.load_full_table(torch.tensor(params["table"]))
    model.bottom_mlp.load_arrays(params["bottom"], dev)
    model.top_mlp.load_arrays(params["top"], dev)
    opt = Adam()
    batches = [[synth_batch(B, V, seed=100 * s + r) for r in range(world)] for s in range(steps)]
    if table_rows is not None:       # ids folded into every table's own cardinality
        batches = [[(c % cards, d, l) for c, d, l in per_rank] for per_rank in batches]
    prob = None
    for s in range(steps):
        cat, dense_x, label = (torch.tensor(a, device=dev) for a in batches[s][rank])
        prob = model({"cat_features": cat, "int_features": dense_x})
        loss = bce_clipped(prob, label)
        loss.backward()
        opt.apply_gradients(model)
    model.embedding_layer.check_overflow()
    torch.cuda.synchronize()
    emb = model.embedding_layer
    max_rows = (emb.total_rows + world - 1) // world
    ids = torch.full((max_rows,), -1, dtype=torch.int64, device=dev)
    rows = torch.zeros(max_rows, D, device=dev)
    ids[: emb.local_rows] = emb.full_row_ids()
    rows[: emb.local_rows] = emb.embeddings
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    all_rows = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(all_ids, ids, group=group)
    dist.all_gather(all_rows, rows, group=group)
    replica_diff = 0.0
    if emb.small_table is not None:          # every rank's copy of the replicated tables must hold the same bits
        copies = [torch.empty_like(emb.small_table) for _ in range(world)]
        dist.all_gather(copies, emb.small_table.contiguous(), group=group)
        replica_diff = max(float((c - copies[0]).abs().max().item()) for c in copies)
    out = None
    if rank == 0:
        full = torch.empty(total, D, device=dev)
        for k in range(world):
            okk = all_ids[k] >= 0
            full[all_ids[k][okk]] = all_rows[k][okk]
        emb.small_into_full(full)
        ref = DLRM(bottom, top, D, V, 26, 13, num_tables=T, device=dev, compute_dtype=compute_dtype, table_rows=table_rows)
        ref.embedding_layer.embeddings.copy_(torch.tensor(params["table"]))
        ref.bottom_mlp.load_arrays(params["bottom"], dev)
        ref.top_mlp.load_arrays(params["top"], dev)
        ropt = Adam()
        rprob = None
        for s in range(steps):
            total = 0
            for r in range(world):        # the replicas' gradients are summed
                cat, dense_x, label = (torch.tensor(a, device=dev) for a in batches[s][r])
                p = ref({"cat_features": cat, "int_features": dense_x})
                if r == 0:
                    rprob = p.detach()
                total = total + bce_clipped(p, label)
            total.backward()
            ropt.apply_gradients(ref)
        torch.cuda.synchronize()
        got, want = full.cpu().numpy(), ref.embedding_layer.embeddings.cpu().numpy()
        moved = np.abs(want - params["table"]) > 0
        out = dict(max_abs_table_diff=float(np.abs(got - want).max()), max_abs_prob_diff=float((prob.detach() - rprob).abs().max().item()),
                   rows_moved=int(moved.any(1).sum()), steps=steps, world=world, replicated_tables=len(emb._small),
                   replica_max_abs_diff=replica_diff,
                   config=f"DLRM D={D}, " + (f"{T} x {V}-row tables" if table_rows is None else f"tables of {min(rows_list)}..{max(rows_list)} rows") +
                          f", B_local={B}, Zipf ids + 2% id 0, "
                          f"{'bf16' if compute_dtype == torch.bfloat16 else 'fp32'} towers, Adam (lazy rows)")
    dist.barrier(group=group)
    return out
