"""DIN attention pooling (csrc/din.cu + the tcgen05 Dense kernels) against the golden fixture frozen from
the reference's own LocalActivationUnit (dien/layers.py:34-59) and against the oracle's bf16-operand restatement.

Bars: valid-position numbering and the bf16 feature rows bit-exact; attention output / gradients within the bf16-operand bound
of the fp32 reference (stated per assertion) and tight against the oracle's bf16 mode, which rounds where the kernels round."""
import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.as_tensor(a).cuda()


@pytest.fixture(scope="module")
def ops(cuda_lib):
    from recommender_b200 import ops as _ops
    return _ops


def _case(g):
    his = O.compute_flat_embedding(g["W_item"], g["W_cat"], g["item"], g["cat"])
    tgt = O.compute_flat_embedding(g["W_item"], g["W_cat"], g["t_item"], g["t_cat"])[:, 0, :]
    layers = [(g[f"att_W{i}"], g[f"att_b{i}"]) for i in range(3)]
    return tgt, his, g["item"] != 0, layers


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("itype", [np.int32, np.int64])
def test_valid_positions_and_feature_rows_bit_exact(ops, golden, itype):
    g = golden("din_attention")
    tgt, his, mask, _ = _case(g)
    B, L, E = his.shape
    h = ops.DinHistory(cu(g["W_item"]), cu(g["item"].astype(itype)), cu(g["W_cat"]), cu(g["cat"].astype(itype)))
    off = ops.din_offsets(h).cpu().numpy()
    np.testing.assert_array_equal(off, np.concatenate([[0], np.cumsum(mask.sum(1))]))
    P = int(off[-1])
    X = ops.din_build_features(h, cu(tgt), cu(off.astype(np.int32)), P, 4 * E).float().cpu().numpy()
    t = np.broadcast_to(tgt[:, None, :], his.shape)
    ref = O.round_bf16(np.concatenate([t, his, t - his, t * his], axis=-1).astype(np.float32))[mask]      # rows in (b, l) order
    np.testing.assert_array_equal(X[:P], ref)
    # the weighted sum and its weight gradient on fp32 rows
    w = np.random.default_rng(0).normal(size=P).astype(np.float32)
    rep = ops.din_pool_fwd(h, cu(off.astype(np.int32)), cu(w)).cpu().numpy()
    wfull = np.zeros((B, L), dtype=np.float32)
    wfull[mask] = w
    np.testing.assert_allclose(rep, np.einsum("bl,ble->be", wfull, his), rtol=1e-5, atol=1e-7)
    d_rep = g["d_rep"]
    dw = ops.din_pool_bwd_weights(h, cu(off.astype(np.int32)), cu(d_rep), P).cpu().numpy()
    np.testing.assert_allclose(dw[:P], np.einsum("be,ble->bl", d_rep, his)[mask], rtol=1e-4, atol=1e-8)


def _build_din(g):
    from recommender_b200.din import DIN
    D = g["W_item"].shape[1]
    model = DIN(g["W_item"].shape[0], D, g["W_cat"].shape[0], D, device="cuda")
    model.item_embedding.embeddings.copy_(cu(g["W_item"]))
    model.cat_embedding.embeddings.copy_(cu(g["W_cat"]))
    model.local_activation_unit.load_arrays([(g[f"att_W{i}"], g[f"att_b{i}"]) for i in range(3)], "cuda")
    return model


def test_din_attention_forward_backward_golden(cuda_lib, golden):
    """dien/model.py:42-51 through the fused path: output, attention-MLP gradients and BOTH tables' gradients (history rows +
    the target row, one optimizer call per table over the concatenated lookup groups) against torch autograd of the reference."""
    from recommender_b200.optimizers import SGD
    g = golden("din_attention")
    tgt, his, mask, layers = _case(g)
    model = _build_din(g)
    inputs = {k: cu(g[n]) for k, n in (("target_item", "t_item"), ("target_cat", "t_cat"), ("pos_his_item", "item"), ("pos_his_cat", "cat"))}
    out = model(inputs)                                             # [target | history_representation]
    E = tgt.shape[1]
    np.testing.assert_array_equal(out[:, :E].detach().cpu().numpy(), tgt)               # the lookup itself is bit-exact
    rep = out[:, E:].detach().cpu().numpy()
    rep16, cache = O.local_activation_unit(tgt, his, mask, layers, operand_dtype="bf16")
    assert _rel(rep, g["rep"]) <= 3e-2                              # bf16 operands vs the fp32 reference
    assert _rel(rep, rep16) <= 2e-3                                 # same rounding points: accumulation order and exp() only
    W_item0, W_cat0 = model.item_embedding.embeddings.clone(), model.cat_embedding.embeddings.clone()
    out.backward(torch.cat([torch.zeros_like(out[:, :E]), cu(g["d_rep"])], dim=1))
    dt16, dh16, grads16 = O.local_activation_unit_backward(cache, g["d_rep"])
    unit = model.local_activation_unit
    for i in range(3):
        dW, db = unit.kernels[i].grad.cpu().numpy(), unit.biases[i].grad.cpu().numpy()
        assert _rel(dW, g[f"att_dW{i}"]) <= 6e-2 and _rel(db, g[f"att_db{i}"]) <= 6e-2, i
        assert _rel(dW, grads16[i][0]) <= 1e-2 and _rel(db, grads16[i][1]) <= 1e-2, i
    SGD(learning_rate=1.0).apply_gradients(model)                   # W -= dW: reads the tables' gradients off the update
    dWi = (W_item0 - model.item_embedding.embeddings).cpu().numpy()
    dWc = (W_cat0 - model.cat_embedding.embeddings).cpu().numpy()
    assert _rel(dWi, g["dW_item"]) <= 6e-2 and _rel(dWc, g["dW_cat"]) <= 6e-2
    D = g["W_item"].shape[1]
    for got, h_idx, t_idx, c0 in ((dWi, g["item"], g["t_item"], 0), (dWc, g["cat"], g["t_cat"], D)):
        ref = np.zeros_like(got)
        np.add.at(ref, h_idx.reshape(-1), dh16[:, :, c0:c0 + D].reshape(-1, D))
        np.add.at(ref, t_idx.reshape(-1), dt16[:, c0:c0 + D])
        assert _rel(got, ref) <= 1e-2
        untouched = np.setdiff1d(np.arange(got.shape[0]), np.concatenate([h_idx.reshape(-1), t_idx.reshape(-1)]))
        assert (got[untouched] == 0).all()


def test_materialised_history_form_matches_the_fused_one(cuda_lib, golden):
    """The reference's call surface `unit((target, history), mask=mask)` (dien/layers.py:42) on tensors: same numbers as the
    fused form, and a dense d_history with exact zeros at the masked positions."""
    from recommender_b200.din import LocalActivationUnit
    g = golden("din_attention")
    tgt, his, mask, layers = _case(g)
    unit = LocalActivationUnit()
    unit.load_arrays(layers, "cuda")
    t = cu(tgt[:, None, :]).requires_grad_()
    h = cu(his).requires_grad_()
    rep = unit((t, h), mask=cu(mask))
    rep.backward(cu(g["d_rep"]))
    rep16, cache = O.local_activation_unit(tgt, his, mask, layers, operand_dtype="bf16")
    dt16, dh16, _ = O.local_activation_unit_backward(cache, g["d_rep"])
    assert _rel(rep.detach().cpu().numpy(), rep16) <= 2e-3 and _rel(rep.detach().cpu().numpy(), g["rep"]) <= 3e-2
    dh = h.grad.cpu().numpy()
    assert (dh[~mask] == 0).all()
    assert _rel(dh, dh16) <= 1e-2 and _rel(dh, g["d_his"]) <= 6e-2
    assert _rel(t.grad[:, 0, :].cpu().numpy(), dt16) <= 1e-2


def test_all_pad_histories(cuda_lib):
    """A sample whose history is all padding contributes a zero representation (weights *= mask, dien/layers.py:54 — unlike
    compute_his_average there is no division); a batch without any valid position runs no GEMM at all."""
    from recommender_b200.din import DIN
    torch.manual_seed(0)
    model = DIN(50, 8, 10, 8, device="cuda")
    B, L = 4, 6
    item = torch.randint(1, 50, (B, L), device="cuda")
    cat = torch.randint(1, 10, (B, L), device="cuda")
    item[1] = 0
    inputs = dict(target_item=torch.randint(1, 50, (B, 1), device="cuda"), target_cat=torch.randint(1, 10, (B, 1), device="cuda"),
                  pos_his_item=item, pos_his_cat=cat)
    out = model(inputs)
    assert (out[1, 16:] == 0).all() and (out[0, 16:] != 0).any()
    out.sum().backward()
    inputs["pos_his_item"] = torch.zeros_like(item)
    out0 = model(inputs)
    assert (out0[:, 16:] == 0).all()
    out0.sum().backward()


def test_config4_size_runs_and_is_deterministic(cuda_lib):
    """BASELINE config 4 shape (history 100, D = 32 + 32) at a batch the test budget allows: two runs give identical bits."""
    from recommender_b200.din import DIN
    B, L, V = 8192, 100, 100_000
    gen = torch.Generator(device="cuda").manual_seed(3)
    lens = torch.randint(1, L + 1, (B, 1), device="cuda", generator=gen)
    item = torch.randint(1, V, (B, L), device="cuda", generator=gen)
    item = torch.where(torch.arange(L, device="cuda")[None] < lens, item, torch.zeros_like(item))
    cat = item % 1000
    inputs = dict(target_item=torch.randint(1, V, (B, 1), device="cuda", generator=gen), pos_his_item=item, pos_his_cat=cat)
    inputs["target_cat"] = inputs["target_item"] % 1000
    outs = []
    for _ in range(2):
        model = DIN(V, 32, 1000, 32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
        out = model(inputs)
        out.square().sum().backward()
        outs.append((out.detach().clone(), model.local_activation_unit.kernels[0].grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.isfinite(outs[0][0]).all() and torch.isfinite(outs[0][1]).all()


def test_three_uses_of_the_tables_in_one_step_with_the_negative_history(cuda_lib, golden):
    """DIEN looks the same two tables up a third time for the NEGATIVE history (dien/model.py:69-71: compute_flat_embedding on
    neg_his_item / neg_his_cat, un-pooled, consumed by the auxiliary loss).  Target lookup + attention-pooled positive history +
    plain gather of the negative history in ONE step: each table's update is one call over three concatenated lookup groups
    (SURVEY A.1), and equals the sum of the three gradients."""
    from recommender_b200.optimizers import SGD
    g = golden("din_attention")
    tgt, his, mask, layers = _case(g)
    model = _build_din(g)
    rng = np.random.default_rng(23)
    B, L = g["item"].shape
    D = g["W_item"].shape[1]
    neg_item = rng.integers(1, g["W_item"].shape[0], size=(B, L)).astype(np.int32)
    neg_cat = rng.integers(1, g["W_cat"].shape[0], size=(B, L)).astype(np.int32)
    d_neg = rng.normal(0, 1e-2, size=(B, L, 2 * D)).astype(np.float32)
    inputs = {k: cu(g[n]) for k, n in (("target_item", "t_item"), ("target_cat", "t_cat"), ("pos_his_item", "item"), ("pos_his_cat", "cat"))}
    out = model(inputs)
    neg = model.compute_flat_embedding((cu(neg_item), cu(neg_cat)))                    # dien/model.py:69-71
    np.testing.assert_array_equal(neg.detach().cpu().numpy(), O.compute_flat_embedding(g["W_item"], g["W_cat"], neg_item, neg_cat))
    E = 2 * D
    W_item0, W_cat0 = model.item_embedding.embeddings.clone(), model.cat_embedding.embeddings.clone()
    torch.autograd.backward([out, neg], [torch.cat([torch.zeros_like(out[:, :E]), cu(g["d_rep"])], dim=1), cu(d_neg)])
    assert len(model.item_embedding.pending) == 3 and len(model.cat_embedding.pending) == 3
    SGD(learning_rate=1.0).apply_gradients(model)
    rep16, cache = O.local_activation_unit(tgt, his, mask, layers, operand_dtype="bf16")
    dt16, dh16, _ = O.local_activation_unit_backward(cache, g["d_rep"])
    for W0, emb, h_idx, t_idx, n_idx, c0 in ((W_item0, model.item_embedding, g["item"], g["t_item"], neg_item, 0),
                                             (W_cat0, model.cat_embedding, g["cat"], g["t_cat"], neg_cat, D)):
        got = (W0 - emb.embeddings).cpu().numpy()
        ref = np.zeros_like(got)
        np.add.at(ref, h_idx.reshape(-1), dh16[:, :, c0:c0 + D].reshape(-1, D))
        np.add.at(ref, t_idx.reshape(-1), dt16[:, c0:c0 + D])
        np.add.at(ref, n_idx.reshape(-1), d_neg[:, :, c0:c0 + D].reshape(-1, D))
        assert _rel(got, ref) <= 1e-2
        only_neg = np.setdiff1d(n_idx.reshape(-1), np.concatenate([h_idx.reshape(-1), t_idx.reshape(-1)]))
        exact = np.zeros_like(got)
        np.add.at(exact, n_idx.reshape(-1), d_neg[:, :, c0:c0 + D].reshape(-1, D))
        np.testing.assert_allclose(got[only_neg], exact[only_neg], rtol=2e-6, atol=8e-9)     # fp32 rows, same summation order; W0 - W rounds at 2^-24 |W|


@pytest.mark.parametrize("D0,D1", [(17, 16), (64, 64), (20, 0), (4, 4)])
def test_attention_pooling_other_row_widths(cuda_lib, D0, D1):
    """Row widths that take the other kernel instantiations (odd widths: one column per lane; E = 128: two column groups per lane;
    a single table) against the oracle's bf16-operand mode — forward, d_target and the history gradient rows."""
    from recommender_b200.din import LocalActivationUnit
    from recommender_b200.layers import Embedding
    rng = np.random.default_rng(D0 * 100 + D1)
    B, L, Vi, Vc = 37, 45, 200, 30
    E = D0 + D1
    Wi = O.init_table(rng, Vi, D0)
    Wc = O.init_table(rng, Vc, D1) if D1 else None
    lens = rng.integers(0, L + 1, size=B)
    item = np.zeros((B, L), dtype=np.int64)
    cat = np.zeros((B, L), dtype=np.int64)
    for b, n in enumerate(lens):
        item[b, :n] = rng.integers(1, Vi, size=n)
        cat[b, :n] = rng.integers(1, Vc, size=n)
    tgt = rng.normal(0, 0.05, size=(B, E)).astype(np.float32)
    layers = O.init_mlp(rng, 4 * E, [80, 40, 1])
    layers = [(W, rng.normal(0, 0.05, size=b.shape).astype(np.float32)) for W, b in layers]
    his = O.embedding_lookup(Wi, item) if not D1 else O.compute_flat_embedding(Wi, Wc, item, cat)
    mask = item != 0
    rep16, cache = O.local_activation_unit(tgt, his, mask, layers, operand_dtype="bf16")
    d_rep = rng.normal(0, 0.1, size=(B, E)).astype(np.float32)
    dt16, dh16, _ = O.local_activation_unit_backward(cache, d_rep)
    ei = Embedding(Vi, D0, mask_zero=True, device="cuda")
    ei.embeddings.copy_(cu(Wi))
    ei.presort = False
    ec = None
    if D1:
        ec = Embedding(Vc, D1, mask_zero=True, device="cuda")
        ec.embeddings.copy_(cu(Wc))
        ec.presort = False
    unit = LocalActivationUnit()
    unit.load_arrays(layers, "cuda")
    t = cu(tgt).requires_grad_()
    rep = unit.attend(t, ei, cu(item), ec, cu(cat) if D1 else None)
    rep.backward(cu(d_rep))
    assert _rel(rep.detach().cpu().numpy(), rep16) <= 3e-3
    assert _rel(t.grad.cpu().numpy(), dt16) <= 2e-2
    src = ei.pending[0].grad.srcs[0]                       # dh[B, L, E]: rows of the valid positions
    got = src.reshape(B, L, E).cpu().numpy()
    assert _rel(got[mask], dh16[mask]) <= 2e-2
    assert ei.pending[0].grad.scale == "masked" and (ec is None or len(ec.pending) == 1)
