"""DeepFM / DLRM with the reference's call surface on the CUDA path, against the fixtures frozen
from the reference's own ctr/model.py (tests/golden) and against multi-step oracle training;
plus size-independent properties at BASELINE config 2's full size."""
import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O

pytestmark = pytest.mark.gpu

BF16_REL = 2.0 ** -7


def cu(a):
    return torch.as_tensor(a).cuda()


def _mlp(g, name):
    layers, i = [], 0
    while f"{name}_W{i}" in g:
        layers.append((g[f"{name}_W{i}"], g[f"{name}_b{i}"]))
        i += 1
    return layers


def _build_dlrm(g, fused=True, num_tables=1, V=None, compute_dtype=None):
    from recommender_b200.model import DLRM
    bottom, top = _mlp(g, "bottom"), _mlp(g, "top")
    D = g["table"].shape[1]
    V = V or g["table"].shape[0]
    model = DLRM([W.shape[1] for W, _ in bottom], [W.shape[1] for W, _ in top], D, V, 26, 13, fused=fused,
                 num_tables=num_tables, device="cuda", compute_dtype=compute_dtype)
    model.embedding_layer.embeddings.copy_(cu(g["table"]))
    model.bottom_mlp.load_arrays(bottom, "cuda")
    model.top_mlp.load_arrays(top, "cuda")
    return model


@pytest.mark.parametrize("name", ["dlrm_small", "dlrm_uniform"])
@pytest.mark.parametrize("fused", [True, False])
def test_dlrm_against_reference_fixture(cuda_lib, golden, name, fused):
    from recommender_b200.model import bce_clipped
    from recommender_b200.optimizers import Adam
    g = golden(name)
    model = _build_dlrm(g, fused)
    x = {"cat_features": cu(g["cat"]), "int_features": cu(g["dense"])}
    prob = model(x)
    assert prob.shape == (g["cat"].shape[0],)
    np.testing.assert_allclose(prob.detach().cpu().numpy(), g["prob"], rtol=0, atol=2e-3)     # bf16 operands in the dot
    params = dict(table=g["table"].copy(), bottom=_mlp(g, "bottom"), top=_mlp(g, "top"))
    ref_prob, cache = O.dlrm_forward(params, g["cat"], g["dense"], operand_dtype="bf16")
    np.testing.assert_allclose(prob.detach().cpu().numpy(), ref_prob, rtol=1e-4, atol=1e-6)   # same arithmetic
    loss = bce_clipped(prob, cu(g["label"]))
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=2e-3)
    loss.backward()
    for i, W in enumerate(model.top_mlp.kernels):
        ref = g[f"top_dW{i}"]
        assert np.abs(W.grad.cpu().numpy() - ref).max() <= 4 * BF16_REL * np.abs(ref).max() + 1e-7
    for i, W in enumerate(model.bottom_mlp.kernels):
        ref = g[f"bottom_dW{i}"]
        assert np.abs(W.grad.cpu().numpy() - ref).max() <= 4 * BF16_REL * np.abs(ref).max() + 1e-7
    # sparse side: one recorded lookup group; Adam step 1 against the oracle chain
    emb = model.embedding_layer
    assert len(emb.pending) == 1 and emb.pending[0].n == g["cat"].size
    _, dprob = O.bce_clipped(ref_prob, g["label"])
    grads = O.dlrm_backward(params, cache, dprob, operand_dtype="bf16")
    st = dict(m=np.zeros_like(g["table"]), v=np.zeros_like(g["table"]))
    table_ref = g["table"].copy()
    O.sparse_backward_update(table_ref, st, g["cat"], grads["dE"], "adam_lazy", 1)
    Adam().apply_gradients(model)
    assert emb.pending == []
    got = emb.embeddings.cpu().numpy()
    touched = np.unique(g["cat"])
    untouched = np.setdiff1d(np.arange(g["table"].shape[0]), touched)
    np.testing.assert_array_equal(got[untouched], g["table"][untouched])                      # lazy: untouched rows stay
    # Adam's first step is ~ lr * sign(g): compare where the gradient is well above epsilon-scale noise
    np.testing.assert_allclose(emb.opt_state["m"].cpu().numpy(), st["m"], rtol=5e-3, atol=1e-9)
    big = np.abs(st["m"]) > 1e-7
    np.testing.assert_allclose(got[big], table_ref[big], rtol=0, atol=2e-5)


def test_deepfm_against_reference_fixture(cuda_lib, golden):
    from recommender_b200.model import DeepFM, bce_logits
    from recommender_b200.optimizers import Adam
    g = golden("deepfm_small")
    mlp = _mlp(g, "mlp")
    D, V = g["table"].shape[1], g["table"].shape[0]
    for fused in (True, False):
        model = DeepFM(D, V, 13, 26, [W.shape[1] for W, _ in mlp], fused=fused, device="cuda")
        model.embedding_layer.embeddings.copy_(cu(g["table"]))
        model.mlp.load_arrays(mlp, "cuda")
        x = {"cat_features": cu(g["cat"]), "int_features": cu(g["dense"])}
        prob = model(x)
        np.testing.assert_allclose(prob.detach().cpu().numpy(), g["prob"], rtol=2e-5, atol=2e-6)   # all fp32
        loss = bce_logits(model.logits(x), cu(g["label"]))
        model.embedding_layer.pending.clear()
        np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-5)
        loss.backward()
        for i, W in enumerate(model.mlp.kernels):
            np.testing.assert_allclose(W.grad.cpu().numpy(), g[f"mlp_dW{i}"], rtol=1e-3, atol=1e-6)
        # table gradient = sum of the three consumers of cat_embedding (ctr/model.py:21,22,25):
        # apply it with SGD(lr=1) so that table_before - table_after IS the dense table gradient
        from recommender_b200.optimizers import SGD
        SGD(1.0).apply_gradients([model.embedding_layer])
        dtable = g["table"] - model.embedding_layer.embeddings.cpu().numpy()
        np.testing.assert_allclose(dtable, g["dtable"], rtol=1e-3, atol=2e-8)


def test_dlrm_multi_step_training_tracks_the_oracle(cuda_lib, golden):
    """5 Adam steps on fresh batches: probabilities, loss and the table stay within tolerance of the
    oracle's DLRM (bf16-operand dot, lazy Adam on the table, Keras Adam on the MLPs)."""
    from recommender_b200.model import bce_clipped
    from recommender_b200.optimizers import Adam
    g = golden("dlrm_uniform")
    D, V = g["table"].shape[1], g["table"].shape[0]
    model = _build_dlrm(g)
    opt = Adam()
    params = dict(table=g["table"].copy(), bottom=[(W.copy(), b.copy()) for W, b in _mlp(g, "bottom")],
                  top=[(W.copy(), b.copy()) for W, b in _mlp(g, "top")])
    st = dict(m=np.zeros_like(params["table"]), v=np.zeros_like(params["table"]))
    dense_state = {}
    for step in range(1, 6):
        cat, dense_x, label = O.synth_batch(128, V, seed=100 + step, dist="zipf")
        prob = model({"cat_features": cu(cat), "int_features": cu(dense_x)})
        loss = bce_clipped(prob, cu(label))
        loss.backward()
        opt.apply_gradients(model)
        ref_prob, cache = O.dlrm_forward(params, cat, dense_x, operand_dtype="bf16")
        ref_loss, dprob = O.bce_clipped(ref_prob, label)
        grads = O.dlrm_backward(params, cache, dprob, operand_dtype="bf16")
        O.sparse_backward_update(params["table"], st, cat, grads["dE"], "adam_lazy", step)
        for name in ("bottom", "top"):
            for i, ((W, b), (dW, db)) in enumerate(zip(params[name], grads[name])):
                for tag, p, gr in (("W", W, dW), ("b", b, db)):
                    m, v = dense_state.setdefault((name, i, tag), (np.zeros_like(p), np.zeros_like(p)))
                    O.adam_dense_param(p, m, v, gr, step)
        np.testing.assert_allclose(prob.detach().cpu().numpy(), ref_prob, rtol=0, atol=5e-4 * step)
        assert abs(loss.item() - float(ref_loss)) <= 5e-4 * step
    got = model.embedding_layer.embeddings.cpu().numpy()
    moved = np.abs(params["table"] - g["table"]).max()
    assert moved > 1e-3                                                       # training did move rows
    assert np.abs(got - params["table"]).mean() <= 0.02 * moved               # and the two tables agree


def test_esmm_style_multi_table_embedding(cuda_lib, golden):
    """config 5 through the layer classes: one Embedding per feature, bag size 1, concat on the
    last axis (esmm/esmm.py:15-19), two consumers."""
    from recommender_b200.layers import Embedding
    from recommender_b200.optimizers import SGD
    g = golden("esmm_small")
    feats = [str(f) for f in g["feats"]]
    embs = {}
    for f in feats:
        W = g[f"W_{f}"]
        embs[f] = Embedding(W.shape[0], W.shape[1], device="cuda")
        embs[f].embeddings.copy_(cu(W))
    inputs = {f: cu(g[f"idx_{f}"]) for f in feats}
    embedding = torch.cat([embs[f](inputs[f]) for f in feats], dim=-1).squeeze(1)        # esmm/esmm.py:16-18
    np.testing.assert_array_equal(embedding.detach().cpu().numpy(), g["emb"])
    rng = np.random.default_rng(3)
    w1, w2 = (cu(rng.normal(size=g["emb"].shape).astype(np.float32)) for _ in range(2))
    ((embedding * w1).sum() + (embedding * w2).sum()).backward()                          # ctr and cvr towers
    SGD(1.0).apply_gradients(list(embs.values()))
    total = (w1 + w2).cpu().numpy()
    D = g[f"W_{feats[0]}"].shape[1]
    for k, f in enumerate(feats):
        ref = g[f"W_{f}"].copy()
        O.sparse_backward_update(ref, {}, g[f"idx_{f}"], total[:, None, k * D:(k + 1) * D], "sgd", lr=1.0)
        np.testing.assert_allclose(embs[f].embeddings.cpu().numpy(), ref, rtol=0, atol=1e-5)


def test_dien_style_masked_history(cuda_lib, golden):
    """config 4 through the layer classes (dien/model.py:14-19,25-31)."""
    from recommender_b200.layers import Embedding, compute_his_average
    from recommender_b200.optimizers import SGD
    g = golden("masked_mean")
    D = g["W_item"].shape[1]
    item_emb = Embedding(g["W_item"].shape[0], D, mask_zero=True, device="cuda")
    cat_emb = Embedding(g["W_cat"].shape[0], D, mask_zero=True, device="cuda")
    item_emb.embeddings.copy_(cu(g["W_item"]))
    cat_emb.embeddings.copy_(cu(g["W_cat"]))
    item, cat = cu(g["item"]), cu(g["cat"])
    assert torch.equal(item_emb.compute_mask(item), item != 0)
    avg = torch.cat([compute_his_average(item_emb, item), compute_his_average(cat_emb, cat, mask_idx=item)], dim=-1)
    np.testing.assert_allclose(avg.detach().cpu().numpy(), g["avg"], rtol=1e-5, atol=1e-7)
    avg.backward(cu(g["davg"]))
    SGD(1.0).apply_gradients([item_emb, cat_emb])
    np.testing.assert_allclose(g["W_item"] - item_emb.embeddings.cpu().numpy(), g["dW_item"], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(g["W_cat"] - cat_emb.embeddings.cpu().numpy(), g["dW_cat"], rtol=1e-4, atol=1e-8)


# ---- BASELINE config 2 at full size: properties that need no oracle run -------------------------------------

FULL_B, FULL_D, FULL_V, FULL_T = 65536, 64, 1_000_000, 26


def test_graphed_train_step_equals_eager(cuda_lib, golden):
    """One CUDA graph per step (graph.GraphedTrainStep: forward + loss + backward + dense and sparse Adam, the
    radix sort on its side stream) must reproduce the eager step: same kernels, same order, same numbers —
    including Adam's step-dependent alpha_t, which the graph reads from device memory."""
    from recommender_b200.graph import GraphedTrainStep
    from recommender_b200.model import bce_clipped
    from recommender_b200.optimizers import Adam
    g = golden("dlrm_uniform")
    cat, dense_x, label = cu(g["cat"]), cu(g["dense"]), cu(g["label"])
    batches = [(cat, dense_x, label), (cat.flip(0).contiguous(), dense_x.flip(0).contiguous(), label.flip(0).contiguous())]
    n_steps = 5

    def run(graphed):
        torch.manual_seed(0)
        model = _build_dlrm(g)
        opt = Adam()
        losses = []
        if graphed:
            gs = GraphedTrainStep(model, opt, bce_clipped, batches[0], warmup=1)    # runs 1 eager + 1 replayed step on batch 0
            done = gs.steps_run
            for i in range(n_steps - done):
                losses.append(float(gs.step(batches[(done + i) % 2]).item()))
        else:
            for i in range(n_steps):
                b = batches[0] if i < 2 else batches[i % 2]
                prob = model({"cat_features": b[0], "int_features": b[1]})
                loss = bce_clipped(prob, b[2])
                loss.backward()
                opt.apply_gradients(model)
                losses.append(float(loss.item()))
            losses = losses[2:]
        torch.cuda.synchronize()
        return losses, model.embedding_layer.embeddings.clone(), [p.detach().clone() for p in model.parameters()], opt.iterations

    l_e, t_e, p_e, it_e = run(False)
    l_g, t_g, p_g, it_g = run(True)
    assert it_e == it_g == n_steps
    np.testing.assert_allclose(l_g, l_e, rtol=1e-6)
    assert torch.equal(t_g, t_e)
    for a, b in zip(p_g, p_e):
        assert torch.equal(a, b)


def test_fused_row_update_inside_the_backward_equals_the_two_phase_step(cuda_lib, golden):
    """optimizers.fuse_sparse_updates (opt-in): rows a step touches once are updated by the interaction backward itself, the
    rest by the sorted reduction over the compacted pairs.  Same model after 5 steps as the plain step — eagerly and as one
    CUDA graph (tables equal up to the association of duplicate sums longer than two terms)."""
    from recommender_b200.graph import GraphedTrainStep
    from recommender_b200.model import bce_clipped
    from recommender_b200.optimizers import Adam
    g = golden("dlrm_small")                 # Zipf ids: singletons, pairs and long chains in one batch
    cat, dense_x, label = cu(g["cat"]), cu(g["dense"]), cu(g["label"])
    n_steps = 5

    def run(mode):
        torch.manual_seed(0)
        model = _build_dlrm(g, compute_dtype=torch.bfloat16)
        opt = Adam()
        if mode == "graph":
            gs = GraphedTrainStep(model, opt, bce_clipped, (cat, dense_x, label), warmup=1, fuse_sparse_updates=True)
            for _ in range(n_steps - gs.steps_run):
                gs.step((cat, dense_x, label))
        else:
            if mode == "fused":
                assert opt.fuse_sparse_updates(model) == 1
            for _ in range(n_steps):
                loss = bce_clipped(model({"cat_features": cat, "int_features": dense_x}), label)
                loss.backward()
                opt.apply_gradients(model)
        torch.cuda.synchronize()
        assert opt.iterations == n_steps
        return model.embedding_layer.embeddings.clone(), [p.detach().clone() for p in model.parameters()]

    t_ref, p_ref = run("plain")
    for mode in ("fused", "graph"):
        t, ps = run(mode)
        torch.testing.assert_close(t, t_ref, rtol=1e-5, atol=1e-7)
        assert (t != t_ref).float().mean().item() < 0.05          # all but the long chains: identical bits
        for a, b in zip(ps, p_ref):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7)


def test_graphs_bound_to_input_buffers_equal_the_copying_step(cuda_lib, golden):
    """GraphedTrainStep.bind_inputs: one captured graph per registered batch, replayed on that batch in place.  Same tables, same MLP
    weights, same losses as the single graph that copies every batch into its static inputs — and an unregistered batch still
    takes the copying path."""
    from recommender_b200.graph import GraphedTrainStep
    from recommender_b200.model import bce_clipped
    from recommender_b200.optimizers import Adam
    g = golden("dlrm_uniform")
    cat, dense_x, label = cu(g["cat"]), cu(g["dense"]), cu(g["label"])
    b0 = (cat, dense_x, label)
    b1 = (cat.flip(0).contiguous(), dense_x.flip(0).contiguous(), label.flip(0).contiguous())
    b2 = (cat.roll(3, 0).contiguous(), dense_x.roll(3, 0).contiguous(), label.roll(3, 0).contiguous())
    order = [b1, b0, b1, b2, b0, b1]

    def run(bind):
        torch.manual_seed(0)
        model = _build_dlrm(g, compute_dtype=torch.bfloat16)
        opt = Adam()
        gs = GraphedTrainStep(model, opt, bce_clipped, b0, warmup=1)
        if bind:
            gs.bind_inputs([b0, b1])
            assert len(gs._bound) == 2
        it0 = opt.iterations
        losses = [float(gs.step(b).item()) for b in order]
        torch.cuda.synchronize()
        assert opt.iterations == it0 + len(order)
        return losses, model.embedding_layer.embeddings.clone(), [p.detach().clone() for p in model.parameters()]

    l_c, t_c, p_c = run(False)
    l_b, t_b, p_b = run(True)
    assert l_b == l_c
    assert torch.equal(t_b, t_c)
    for a, b in zip(p_b, p_c):
        assert torch.equal(a, b)


def test_losses_read_on_the_side_stream_are_the_steps_losses(cuda_lib, golden):
    """GraphedTrainStep.step(loss_out=pinned slot): the loss of EVERY step reaches the host without a synchronisation in the loop,
    for bound graphs (each owns its loss tensor) and for the copying graph (one loss tensor rewritten every step)."""
    from recommender_b200.graph import GraphedTrainStep
    from recommender_b200.model import bce_clipped
    from recommender_b200.optimizers import Adam
    g = golden("dlrm_uniform")
    cat, dense_x, label = cu(g["cat"]), cu(g["dense"]), cu(g["label"])
    b0 = (cat, dense_x, label)
    b1 = (cat.flip(0).contiguous(), dense_x.flip(0).contiguous(), label.flip(0).contiguous())
    b2 = (cat.roll(3, 0).contiguous(), dense_x.roll(3, 0).contiguous(), label.roll(3, 0).contiguous())
    order = [b1, b0, b1, b2, b2, b0, b1, b0, b0, b2] * 3

    def run(async_read):
        torch.manual_seed(0)
        model = _build_dlrm(g, compute_dtype=torch.bfloat16)
        gs = GraphedTrainStep(model, Adam(), bce_clipped, b0, warmup=1)
        gs.bind_inputs([b0, b1])
        host = torch.full((len(order),), float("nan")).pin_memory()
        if async_read:
            for i, b in enumerate(order):
                gs.step(b, loss_out=host[i])
            torch.cuda.synchronize()
            return host.tolist()
        return [float(gs.step(b).item()) for b in order]

    assert run(True) == run(False)


@pytest.fixture(scope="module")
def full(cuda_lib):
    torch.manual_seed(4)
    W = torch.empty(FULL_V * FULL_T, FULL_D, device="cuda").uniform_(-0.05, 0.05)
    idx = torch.randint(0, FULL_V, (FULL_B, FULL_T), device="cuda", dtype=torch.int64)
    idx[torch.rand(FULL_B, FULL_T, device="cuda") < 0.02] = 0                  # OOV -> 0 hot rows per table
    off = torch.arange(FULL_T, device="cuda", dtype=torch.int64) * FULL_V
    return W, idx, off


def test_full_size_gather_bit_exact(full):
    from recommender_b200 import ops
    W, idx, off = full
    out = ops.gather_fwd(W, idx, L=FULL_T, field_row_offset=off)
    ref = W[(idx + off[None]).reshape(-1)].reshape(FULL_B, FULL_T, FULL_D)
    assert torch.equal(out.view(torch.int32), ref.view(torch.int32))


def test_full_size_interaction_properties(full):
    from recommender_b200 import ops
    W, idx, off = full
    dv = torch.randn(FULL_B, FULL_D, device="cuda") * 0.1
    out = ops.dot_interaction_fwd(table=W, idx=idx, field_row_offset=off, dense_vec=dv, tail=True)
    E = ops.gather_fwd(W, idx, L=FULL_T, field_row_offset=off)
    out2 = ops.dot_interaction_fwd(E=E, dense_vec=dv, tail=True)
    assert torch.equal(out, out2)                                              # fused gather == materialised E
    Z = out[:, :729].reshape(FULL_B, 27, 27)
    assert (torch.tril(Z) == 0).all()                                          # strict upper triangle kept
    # power-of-two scaling is exact in bf16 and fp32: Z(2X) == 4 Z(X) bit for bit
    out4 = ops.dot_interaction_fwd(E=E * 2.0, dense_vec=dv * 2.0, tail=False)
    assert torch.equal(out4, out[:, :729] * 4.0)
    # against fp32 cuBLAS on a slice
    X = torch.cat([E[:4096], dv[:4096, None]], dim=1)
    ref = torch.triu(torch.bmm(X, X.transpose(1, 2)), diagonal=1)
    assert (Z[:4096] - ref).abs().max() <= BF16_REL * ref.abs().max()


def test_full_size_scatter_counts_and_determinism(full):
    from recommender_b200 import ops
    from recommender_b200.ops import GradSource, LookupGroup
    W, idx, off = full
    rows = (idx + off[None]).reshape(-1)
    # SGD(lr=1) with an all-ones gradient subtracts each row's multiplicity: exact integer sums
    Wc = W.clone()
    ones = torch.ones(FULL_B, FULL_T, FULL_D, device="cuda")
    grp = [LookupGroup(idx, FULL_T, GradSource.per_position(ones, FULL_T), field_row_offset=off)]
    ops.sparse_bwd_update(Wc, None, None, grp, optimizer="sgd", lr=1.0)
    counts = torch.bincount(rows, minlength=W.shape[0]).to(torch.float32)
    assert torch.equal(Wc, W - counts[:, None])
    del Wc, ones
    # Adam twice from the same state: bit-identical tables (no atomics anywhere)
    dE = torch.randn(FULL_B, FULL_T, FULL_D, device="cuda") * 1e-3
    res = []
    for _ in range(2):
        Wa, m, v = W.clone(), torch.zeros_like(W), torch.zeros_like(W)
        ops.sparse_bwd_update(Wa, m, v, [LookupGroup(idx, FULL_T, GradSource.per_position(dE, FULL_T), field_row_offset=off)],
                              optimizer="adam_lazy", step=1)
        res.append((Wa, m, v))
    for a, b in zip(*res):
        assert torch.equal(a, b)
    # gradient conservation: sum over rows of m / (1 - beta_1) == sum over lookups of dE (fp64)
    m = res[0][1]
    lhs = m.double().sum(0) / (1.0 - np.float32(0.9))
    rhs = dE.double().sum((0, 1))
    assert (lhs - rhs).abs().max() <= 1e-4 * dE.double().abs().sum((0, 1)).max()
    # untouched rows did not move
    touched = torch.zeros(W.shape[0], dtype=torch.bool, device="cuda")
    touched[rows] = True
    assert torch.equal(res[0][0][~touched], W[~touched])
    ops.check_oob("cuda")


def test_two_real_ranks_sharded_equals_unsharded(cuda_lib):
    """The peer-memory sharded step on REAL ranks (two processes, two GPUs, NVLink peer reads, device-side barriers)
    against the unsharded model on rank 0 (scripts/p2p_check.py).  Needs two visible GPUs."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(repo, "scripts", "p2p_check.py")], capture_output=True, text=True, timeout=600,
                       env=env, cwd=repo)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") >= 5         # equal tables (fp32 / bf16 towers), unequal tables, and the small ones replicated (fp32 / bf16)
    assert '"replicated_tables": 6' in r.stdout


# ---- BASELINE config 2 at full size AGAINST THE NUMPY ORACLE (not only properties) ----------------------------------------

def _full_ids(dist, tables, rng):
    """[B, 26] ids at config 2's size: uniform, or Zipf(1.05) folded mod V + 2 % forced id 0."""
    if dist == "uniform":
        ids = rng.integers(0, FULL_V, size=(FULL_B, 26), dtype=np.int64)
    else:
        ids = (rng.zipf(1.05, size=(FULL_B, 26)).astype(np.uint64) % np.uint64(FULL_V)).astype(np.int64)
        ids[rng.random((FULL_B, 26)) < 0.02] = 0
    return ids


@pytest.mark.parametrize("dist", ["uniform", "zipf"])
@pytest.mark.parametrize("tables", [26, 1])
def test_full_size_interaction_and_sparse_update_against_the_oracle(cuda_lib, dist, tables):
    """B = 65536, D = 64, 1M-row tables, T = 26 and T = 1, uniform and Zipf + 2 % id 0 ids.
    (a) the fused lookup + interaction row of ALL 65536 samples against oracle.dot_interaction in its bf16-operand mode
        (same arithmetic: 1e-5 of the largest entry; the bf16 output row: one bf16 rounding);
    (b) the deterministic backward — sort, duplicate-row sum in input order, fused lazy Adam — on >= 10 000 sampled
        touched rows (the hottest rows included) against oracle.dedup_indexed_slices + oracle.adam_lazy run on exactly
        the lookups that hit those rows: rows hit once bit-exact, summed rows to fp32 re-association."""
    from recommender_b200 import ops
    from recommender_b200.ops import GradSource, LookupGroup
    rng = np.random.default_rng(7 + tables + (dist == "zipf"))
    T, V, D, B, F = tables, FULL_V, FULL_D, FULL_B, 26
    torch.manual_seed(11)
    W = torch.empty(V * T, D, device="cuda").uniform_(-0.05, 0.05)
    ids = _full_ids(dist, T, rng)
    idx = cu(ids)
    off = torch.arange(T, device="cuda", dtype=torch.int64) * V if T > 1 else None
    rows = ids + (np.arange(T, dtype=np.int64) * V)[None] if T > 1 else ids
    dv = torch.randn(B, D, device="cuda") * 0.1
    # ---- (a) interaction, all samples
    out = ops.dot_interaction_fwd(table=W, idx=idx, field_row_offset=off, dense_vec=dv, tail=True)
    Wn, dvn = W.cpu().numpy(), dv.cpu().numpy()
    worst = 0.0
    for lo in range(0, B, 8192):                                   # the oracle in slabs: 8192 x 27 x 64 floats at a time
        hi = lo + 8192
        X = np.concatenate([Wn[rows[lo:hi]], dvn[lo:hi, None, :]], axis=1)
        ref = np.concatenate([O.dot_interaction(X, False, True, operand_dtype="bf16"), dvn[lo:hi]], axis=1)
        got = out[lo:hi].cpu().numpy()
        worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
    assert worst <= 1e-5, worst
    out16 = ops.dot_interaction_fwd(table=W, idx=idx, field_row_offset=off, dense_vec=dv, tail=True, out_dtype=torch.bfloat16, pad_to=8,
                                    ones_col=True)
    assert (out16[:, 793] == 1).all() and (out16[:, 794:] == 0).all()
    assert (out16[:, :793].float() - out).abs().max().item() <= 2.0 ** -8 * out.abs().max().item()
    # ---- (b) backward scatter + lazy Adam on sampled touched rows
    dE = torch.randn(B, F, D, device="cuda") * 1e-3
    m, v = torch.zeros_like(W), torch.zeros_like(W)
    W0 = Wn                                                         # the table before the update (host copy)
    ops.sparse_bwd_update(W, m, v, [LookupGroup(idx, F, GradSource.per_position(dE, F), field_row_offset=off)], optimizer="adam_lazy", step=3)
    flat = rows.reshape(-1)
    uniq, counts = np.unique(flat, return_counts=True)
    hot = uniq[np.argsort(-counts)[:64]]                            # the longest chains (id 0 / the Zipf head)
    sample = np.unique(np.concatenate([hot, rng.choice(uniq, size=min(10_000, uniq.size), replace=False)]))
    assert sample.size >= min(10_000, uniq.size)
    sel = np.isin(flat, sample)
    pos = np.nonzero(sel)[0]                                        # ascending = input order
    dEn = dE.reshape(-1, D).cpu().numpy()
    r_rows, r_g = O.dedup_indexed_slices(flat[pos], dEn[pos])
    ref_W, ref_m, ref_v = W0.copy(), np.zeros_like(W0), np.zeros_like(W0)
    O.adam_lazy(ref_W, ref_m, ref_v, r_rows, r_g, 3)
    got_W, got_m, got_v = W[sample].cpu().numpy(), m[sample].cpu().numpy(), v[sample].cpu().numpy()
    cnt = counts[np.searchsorted(uniq, sample)]
    once = cnt == 1
    if once.any():                                                  # no summation involved: bit-exact
        np.testing.assert_array_equal(got_m[once], ref_m[sample][once])
        np.testing.assert_array_equal(got_v[once], ref_v[sample][once])
        np.testing.assert_array_equal(got_W[once], ref_W[sample][once])
    # fp32 re-association of the duplicate sums: a hot row adds ~2000 gradients of magnitude 1e-3 whose sum nearly cancels,
    # so the error scales with sum |g| (~2) x 2^-24 x (1 - beta_1), not with the result
    # (one shared table under Zipf ids: the hottest row sums ~170 000 gradients, relative error of the sum ~1e-4)
    np.testing.assert_allclose(got_m, ref_m[sample], rtol=5e-4, atol=5e-8)
    np.testing.assert_allclose(got_v, ref_v[sample], rtol=1e-3, atol=1e-12)
    # Adam's update is ~ alpha * m / sqrt(v): compare where the row moved at all, at a tolerance far below one lr step
    np.testing.assert_allclose(got_W, ref_W[sample], rtol=0, atol=2e-5)
    # rows outside the batch did not move
    probe = rng.integers(0, V * T, size=4096)
    probe = probe[~np.isin(probe, uniq)]
    assert torch.equal(W[probe].cpu(), torch.tensor(W0[probe]))
    ops.check_oob("cuda")
