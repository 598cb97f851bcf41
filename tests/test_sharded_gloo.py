"""The N > 1 path on CPU: world_size-2 `gloo`, the CUDA kernels replaced by tests/fake_ops.py.

Covers the host logic of recommender_b200/sharded.py — sharding maps, exchange plans, all-to-all
split sizes, in-place consumption through the inverse permutation, gradient routing back to the
owners, the MLP all-reduce — by comparing a sharded 2-rank DLRM training run with the
single-process oracle on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ctr_oracle as O

WORLD = 2
V, D, T, B_LOCAL = 60, 16, 26, 24
BOTTOM, TOP = [24, D], [20, 1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_run(num_tables, steps):
    """Single process, concatenated global batch, loss = SUM over replicas of the local means
    (MirroredStrategy with Reduction.NONE)."""
    params = O.init_dlrm(4, BOTTOM, TOP, D, V * num_tables)
    st = dict(m=np.zeros_like(params["table"]), v=np.zeros_like(params["table"]))
    dense_state, probs = {}, []
    off = np.arange(num_tables)[None] * V if num_tables > 1 else 0
    for step in range(1, steps + 1):
        parts = [O.synth_batch(B_LOCAL, V, seed=10 * step + r, dist="zipf") for r in range(WORLD)]
        cat = np.concatenate([p[0] for p in parts]) + off
        dense_x = np.concatenate([p[1] for p in parts])
        label = np.concatenate([p[2] for p in parts])
        prob, cache = O.dlrm_forward(params, cat, dense_x, operand_dtype="bf16")
        dprob = np.concatenate([O.bce_clipped(prob[r * B_LOCAL:(r + 1) * B_LOCAL], parts[r][2])[1] for r in range(WORLD)])
        grads = O.dlrm_backward(params, cache, dprob, operand_dtype="bf16")
        O.sparse_backward_update(params["table"], st, cat, grads["dE"], "adam_lazy", step)
        for name in ("bottom", "top"):
            for i, ((W, b), (dW, db)) in enumerate(zip(params[name], grads[name])):
                for tag, p, g in (("W", W, dW), ("b", b, db)):
                    m, v = dense_state.setdefault((name, i, tag), (np.zeros_like(p), np.zeros_like(p)))
                    O.adam_dense_param(p, m, v, g, step)
        probs.append(prob)
    return params, probs


def _worker(rank, port, sharding, num_tables, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.set_num_threads(1)
    try:
        import tests.fake_ops as fake
        import recommender_b200.layers as layers
        import recommender_b200.optimizers as optimizers
        import recommender_b200.sharded as sharded
        import recommender_b200.model as model_mod
        layers.ops = sharded.ops = optimizers.ops = model_mod.ops = fake      # inject the CPU kernels
        from recommender_b200.model import bce_clipped
        params = O.init_dlrm(4, BOTTOM, TOP, D, V * num_tables)
        model = sharded.ShardedDLRM(BOTTOM, TOP, D, V, 26, 13, num_tables=num_tables, sharding=sharding, device="cpu")
        model.embedding_layer.load_full_table(torch.tensor(params["table"]))
        model.bottom_mlp.load_arrays(params["bottom"], "cpu")
        model.top_mlp.load_arrays(params["top"], "cpu")
        opt = optimizers.Adam()
        probs = []
        for step in range(1, steps + 1):
            cat, dense_x, label = O.synth_batch(B_LOCAL, V, seed=10 * step + rank, dist="zipf")
            prob = model({"cat_features": torch.tensor(cat), "int_features": torch.tensor(dense_x)})
            loss = bce_clipped(prob, torch.tensor(label))
            loss.backward()
            opt.apply_gradients(model)
            probs.append(prob.detach().numpy().copy())
        emb = model.embedding_layer
        out[rank] = dict(probs=probs, shard=emb.shard.embeddings.numpy().copy(), rows=emb.full_row_ids().numpy().copy(),
                         W0=model.top_mlp.kernels[0].detach().numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sharding,num_tables", [("row", 1), ("row", 26), ("table", 26)])
def test_sharded_dlrm_two_ranks_match_single_process_oracle(sharding, num_tables):
    steps = 3
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(_free_port(), sharding, num_tables, steps, out), nprocs=WORLD, join=True)
    params, probs = _oracle_run(num_tables, steps)
    for r in range(WORLD):
        res = out[r]
        for step in range(steps):
            np.testing.assert_allclose(res["probs"][step], probs[step][r * B_LOCAL:(r + 1) * B_LOCAL], rtol=0, atol=3e-5 * (step + 1))
        # every shard row equals the oracle's full-table row it stands for; shards tile the table
        np.testing.assert_allclose(res["shard"][: len(res["rows"])], params["table"][res["rows"]], rtol=0, atol=2e-5)
        np.testing.assert_allclose(res["W0"], params["top"][0][0], rtol=0, atol=2e-5)
    all_rows = np.sort(np.concatenate([out[r]["rows"] for r in range(WORLD)]))
    np.testing.assert_array_equal(all_rows, np.arange(V * num_tables))
    np.testing.assert_array_equal(out[0]["W0"], out[1]["W0"])                # replicas stay in lock-step


def test_table_wise_static_plan_layout():
    """Bucket order of the table-wise plan: owner-major, then sample-major, then the owner's tables."""
    import recommender_b200.sharded as sharded
    semb = sharded.ShardedEmbedding.__new__(sharded.ShardedEmbedding)
    torch.nn.Module.__init__(semb)
    semb.world, semb.rank, semb.num_tables, semb.input_dim = 4, 1, 26, 100
    semb.table_owner = [t % 4 for t in range(26)]
    semb.owned = [t for t in range(26) if t % 4 == 1]
    semb._static = {}
    perm, inv, local_off, counts = semb._table_wise_static(3, 26, "cpu")
    assert counts == [3 * 7, 3 * 7, 3 * 6, 3 * 6]
    assert perm[:7].tolist() == [0, 4, 8, 12, 16, 20, 24]                    # sample 0, owner 0's tables
    assert perm[7:14].tolist() == [26 + c for c in (0, 4, 8, 12, 16, 20, 24)]
    assert inv[perm.long()].tolist() == list(range(78))
    assert local_off[5].item() == 100 and local_off[1].item() == 0          # table 5 is owner 1's 2nd table
