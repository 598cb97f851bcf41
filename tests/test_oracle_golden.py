"""The numpy oracle against the fixtures frozen from the reference's own Python files
(tests/golden/make_golden.py: ctr/model.py, ctr/layers.py, dien/layers.py, dien/model.py and
esmm/esmm.py imported byte-for-byte under the tensorflow shim).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import ctr_oracle as O

RTOL, ATOL = 2e-5, 2e-6   # fp32 oracle vs fp32 torch-CPU replay of the reference graph


def _mlp(g, name):
    layers, i = [], 0
    while f"{name}_W{i}" in g:
        layers.append((g[f"{name}_W{i}"], g[f"{name}_b{i}"]))
        i += 1
    return layers


@pytest.mark.parametrize("name", ["dlrm_small", "dlrm_uniform"])
def test_dlrm_forward_backward(golden, name):
    g = golden(name)
    params = dict(table=g["table"], bottom=_mlp(g, "bottom"), top=_mlp(g, "top"))
    prob, cache = O.dlrm_forward(params, g["cat"], g["dense"])
    np.testing.assert_allclose(prob, g["prob"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(cache["X"], g["X"], rtol=0, atol=0)           # gather + concat: bit-exact
    np.testing.assert_allclose(cache["inter"], g["inter"], rtol=RTOL, atol=ATOL)
    loss, dprob = O.bce_clipped(prob, g["label"])
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-5)
    grads = O.dlrm_backward(params, cache, dprob)
    np.testing.assert_allclose(grads["dE"], g["dX"][:, :26], rtol=1e-4, atol=1e-7)
    for i, (dW, db) in enumerate(grads["top"]):
        np.testing.assert_allclose(dW, g[f"top_dW{i}"], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(db, g[f"top_db{i}"], rtol=1e-3, atol=1e-6)
    for i, (dW, db) in enumerate(grads["bottom"]):
        np.testing.assert_allclose(dW, g[f"bottom_dW{i}"], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(db, g[f"bottom_db{i}"], rtol=1e-3, atol=1e-6)
    # dense table gradient == scatter-add of the IndexedSlices (A.1/A.2)
    rows, summed = O.dedup_indexed_slices(*O.gather_backward(g["cat"], grads["dE"]))
    dtable = np.zeros_like(g["table"])
    dtable[rows] = summed
    np.testing.assert_allclose(dtable, g["dtable"], rtol=1e-4, atol=1e-7)


def test_dot_interaction_operand_rounding_is_small(golden):
    g = golden("dlrm_small")
    a = O.dot_interaction(g["X"])
    b = O.dot_interaction(g["X"], operand_dtype="bf16")
    np.testing.assert_allclose(a, g["inter"], rtol=RTOL, atol=ATOL)
    scale = np.abs(a).max()
    assert np.abs(a - b).max() <= 2.0 ** -7 * scale      # two bf16-rounded operands, fp32 accumulate


def test_deepfm_forward_backward(golden):
    g = golden("deepfm_small")
    params = dict(table=g["table"], mlp=_mlp(g, "mlp"))
    prob, cache = O.deepfm_forward(params, g["cat"], g["dense"])
    np.testing.assert_allclose(prob, g["prob"], rtol=RTOL, atol=ATOL)
    np.testing.assert_array_equal(cache["E"], g["E"])
    np.testing.assert_allclose(cache["fm"], g["fm"], rtol=1e-4, atol=1e-6)
    loss, dlogit = O.bce_logits(cache["logit"], g["label"])
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-5)
    grads = O.deepfm_backward(params, cache, dlogit)
    np.testing.assert_allclose(grads["dE"], g["dE"], rtol=1e-4, atol=1e-7)
    for i, (dW, db) in enumerate(grads["mlp"]):
        np.testing.assert_allclose(dW, g[f"mlp_dW{i}"], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(db, g[f"mlp_db{i}"], rtol=1e-3, atol=1e-6)
    rows, summed = O.dedup_indexed_slices(*O.gather_backward(g["cat"], grads["dE"]))
    dtable = np.zeros_like(g["table"])
    dtable[rows] = summed
    np.testing.assert_allclose(dtable, g["dtable"], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("si", [False, True])
@pytest.mark.parametrize("sg", [False, True])
def test_dot_interaction_all_modes(golden, si, sg):
    g = golden("dot_interaction")
    tag = f"si{int(si)}_sg{int(sg)}"
    out = O.dot_interaction(g["X"], si, sg)
    assert out.shape == g[f"out_{tag}"].shape
    np.testing.assert_allclose(out, g[f"out_{tag}"], rtol=RTOL, atol=ATOL)
    dX = O.dot_interaction_backward(g["X"], g[f"dout_{tag}"], si, sg)
    np.testing.assert_allclose(dX, g[f"dX_{tag}"], rtol=1e-4, atol=1e-5)


def test_masked_mean(golden):
    g = golden("masked_mean")
    avg = O.masked_mean_lookup(g["W_item"], g["W_cat"], g["item"], g["cat"])
    np.testing.assert_allclose(avg, g["avg"], rtol=RTOL, atol=ATOL)
    mask = g["item"] != 0
    dhis = O.masked_mean_backward(g["davg"], mask)
    np.testing.assert_allclose(dhis, g["dhis"], rtol=1e-5, atol=1e-9)
    D = g["W_item"].shape[1]
    for W, idx, d, key in ((g["W_item"], g["item"], dhis[..., :D], "dW_item"), (g["W_cat"], g["cat"], dhis[..., D:], "dW_cat")):
        rows, summed = O.dedup_indexed_slices(*O.gather_backward(idx, np.ascontiguousarray(d)))
        dW = np.zeros_like(W)
        dW[rows] = summed
        np.testing.assert_allclose(dW, g[key], rtol=1e-4, atol=1e-8)


def test_esmm_multi_table(golden):
    g = golden("esmm_small")
    feats = [str(f) for f in g["feats"]]
    tables = {f: g[f"W_{f}"] for f in feats}
    inputs = {f: g[f"idx_{f}"] for f in feats}
    emb = O.compute_embedding_multi(tables, inputs)
    np.testing.assert_array_equal(emb, g["emb"])


# ---- self-checks of the restated TF-internal semantics --------------------

def test_fm_identity_and_finite_difference():
    rng = np.random.default_rng(0)
    E = rng.normal(0, 0.3, size=(5, 7, 4))
    fm = O.fm_second_order_f64(E)
    pair = sum((E[:, i] * E[:, j]).sum(1) for i in range(7) for j in range(i + 1, 7))
    np.testing.assert_allclose(fm, pair, rtol=1e-12)
    g = rng.normal(size=5)
    dE = O.fm_backward(E.astype(np.float32), g.astype(np.float32))
    eps = 1e-6
    for (b, f, d) in [(0, 0, 0), (2, 3, 1), (4, 6, 3)]:
        Ep, Em = E.copy(), E.copy()
        Ep[b, f, d] += eps
        Em[b, f, d] -= eps
        num = ((O.fm_second_order_f64(Ep) - O.fm_second_order_f64(Em)) * g).sum() / (2 * eps)
        assert abs(num - dE[b, f, d]) < 1e-4


def test_dot_interaction_structure():
    rng = np.random.default_rng(1)
    X = rng.normal(size=(3, 27, 8)).astype(np.float32)
    out = O.dot_interaction(X, False, True)
    assert out.shape == (3, 729)
    nz = out.reshape(3, 27, 27) != 0
    assert nz.sum() == 3 * 351 and not np.tril(nz).any()
    assert O.dot_interaction(X, False, False).shape == (3, 351)
    assert O.dot_interaction(X, True, False).shape == (3, 378)


def test_dedup_first_occurrence_order_and_input_order_sum():
    idx = np.array([5, 2, 5, 9, 2, 5], dtype=np.int64)
    vals = np.array([[1e8], [1.0], [-1e8], [3.0], [2.0], [1.0]], dtype=np.float32)
    rows, summed = O.dedup_indexed_slices(idx, vals)
    assert rows.tolist() == [5, 2, 9]
    # input-order fp32 sum: (1e8 + -1e8) + 1 == 1 ; a different association would give 0
    assert summed[:, 0].tolist() == [1.0, 3.0, 3.0]


def test_adam_lazy_equals_tf_dense_on_step_one_and_differs_later():
    rng = np.random.default_rng(2)
    W0 = rng.uniform(-0.05, 0.05, size=(50, 8)).astype(np.float32)
    idx1 = rng.integers(0, 50, size=(16, 3))
    idx2 = rng.integers(0, 50, size=(16, 3))
    d1 = rng.normal(0, 1e-3, size=(16, 3, 8)).astype(np.float32)
    d2 = rng.normal(0, 1e-3, size=(16, 3, 8)).astype(np.float32)
    A = dict(W=W0.copy(), m=np.zeros_like(W0), v=np.zeros_like(W0))
    B = dict(W=W0.copy(), m=np.zeros_like(W0), v=np.zeros_like(W0))
    O.sparse_backward_update(A["W"], A, idx1, d1, "adam_lazy", 1)
    O.sparse_backward_update(B["W"], B, idx1, d1, "adam_tf_dense", 1)
    np.testing.assert_array_equal(A["W"], B["W"])
    O.sparse_backward_update(A["W"], A, idx2, d2, "adam_lazy", 2)
    O.sparse_backward_update(B["W"], B, idx2, d2, "adam_tf_dense", 2)
    untouched = np.setdiff1d(np.unique(idx1), np.unique(idx2))
    assert untouched.size and not np.array_equal(A["W"][untouched], B["W"][untouched])   # momentum keeps moving them


def test_adam_squares_the_summed_gradient():
    W = np.zeros((4, 2), np.float32)
    st = dict(m=np.zeros_like(W), v=np.zeros_like(W))
    idx = np.array([[1, 1]])
    dE = np.array([[[1.0, 2.0], [3.0, 4.0]]], np.float32)
    O.sparse_backward_update(W, st, idx, dE, "adam_lazy", 1)
    np.testing.assert_allclose(st["v"][1], 0.001 * np.array([16.0, 36.0]), rtol=1e-4)   # (1+3)^2, (2+4)^2


def test_id_to_row_and_sharding_are_consistent():
    ids = np.array([-1, 0, 5, 2 ** 40 + 3, 999_999], dtype=np.int64)
    rows = O.id_to_row(ids, 1000)
    assert rows.tolist() == [int((2 ** 64 - 1) % 1000), 0, 5, (2 ** 40 + 3) % 1000, 999]
    owner, local = O.shard_of_row(rows, 8)
    np.testing.assert_array_equal(local * 8 + owner, rows)


def test_bf16_rounding_matches_torch():
    import torch
    x = np.random.default_rng(3).normal(size=4096).astype(np.float32)
    ref = torch.tensor(x).to(torch.bfloat16).float().numpy()
    np.testing.assert_array_equal(O.round_bf16(x), ref)


# ---- independent cross-checks of the TF-internal restatements --------------------------------------
# TensorFlow is absent, so these pieces cannot be pinned on the reference's own run.  Two implementations that ARE in
# this image follow the same published definitions and serve as known-answer checks of the oracle: scikit-learn's
# AdamOptimizer (Kingma & Ba's "epsilon-hat" form, the one Keras OptimizerV2.Adam uses: alpha_t = lr*sqrt(1-b2^t)/(1-b1^t),
# var -= alpha_t * m / (sqrt(v) + eps)) and pandas.factorize (uniques in first-occurrence order, as tf.unique).

def test_keras_adam_formula_against_sklearn_adam():
    from sklearn.neural_network._stochastic_optimizers import AdamOptimizer
    rng = np.random.default_rng(3)
    p0 = rng.normal(0, 0.05, size=(37, 16))
    sk_params = [p0.astype(np.float64).copy()]
    sk = AdamOptimizer(sk_params, learning_rate_init=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7)
    p = p0.astype(np.float32).copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    for step in range(1, 8):
        g = rng.normal(0, 1e-2, size=p.shape)
        sk.update_params(sk_params, [g.astype(np.float64)])
        O.adam_dense_param(p, m, v, g.astype(np.float32), step)
        np.testing.assert_allclose(p, sk_params[0], rtol=2e-6, atol=2e-8, err_msg=f"step {step}")   # fp32 vs fp64: a few ULP of 0.05


def test_keras_sparse_adam_is_dense_adam_on_the_scattered_gradient():
    """A.3: Keras Adam._resource_apply_sparse decays m and v and moves var on EVERY row; rows without a gradient see
    g = 0.  That is dense Adam on the scattered gradient — checked against scikit-learn's Adam over several steps, with
    duplicate ids (summed first, A.2)."""
    from sklearn.neural_network._stochastic_optimizers import AdamOptimizer
    rng = np.random.default_rng(5)
    V, D = 50, 8
    W0 = O.init_table(rng, V, D)
    sk_params = [W0.astype(np.float64).copy()]
    sk = AdamOptimizer(sk_params, learning_rate_init=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7)
    W = W0.copy()
    st = dict(m=np.zeros_like(W), v=np.zeros_like(W))
    for step in range(1, 6):
        idx = rng.integers(0, V, size=(12, 3))
        dE = rng.normal(0, 1e-2, size=(12, 3, D)).astype(np.float32)
        dense_g = np.zeros((V, D))
        np.add.at(dense_g, idx.reshape(-1), dE.reshape(-1, D).astype(np.float64))
        sk.update_params(sk_params, [dense_g])
        O.sparse_backward_update(W, st, idx, dE, "adam_tf_dense", step=step)
        np.testing.assert_allclose(W, sk_params[0], rtol=5e-6, atol=2e-8, err_msg=f"step {step}")
    untouched = np.setdiff1d(np.arange(V), idx.reshape(-1))
    assert untouched.size and not np.array_equal(W[untouched], W0[untouched])      # rows without a gradient moved too


def test_lazy_adam_against_torch_sparse_adam():
    """The row-sparse rule the benchmarked step uses (oracle.adam_lazy = the fused CUDA row update): moments and rows move on
    touched rows only, the bias correction counts GLOBAL steps.  torch.optim.SparseAdam implements the same published rule
    (coalesced gradient, m / v updated at its indices only, step_size = lr * sqrt(1 - b2^t) / (1 - b1^t), eps added to sqrt(v))."""
    import torch
    rng = np.random.default_rng(21)
    V, D = 60, 8
    W0 = rng.normal(0, 0.05, size=(V, D)).astype(np.float32)
    W = W0.copy()
    st = dict(m=np.zeros_like(W), v=np.zeros_like(W))
    p = torch.nn.Parameter(torch.tensor(W0))
    opt = torch.optim.SparseAdam([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-7)
    never = np.ones(V, bool)
    for step in range(1, 7):
        idx = rng.integers(0, 40, size=(16, 3))                       # duplicates inside a step; rows 40.. never touched
        dE = rng.normal(0, 1e-2, size=(16, 3, D)).astype(np.float32)
        never[idx.reshape(-1)] = False
        O.sparse_backward_update(W, st, idx, dE, "adam_lazy", step=step)
        p.grad = torch.sparse_coo_tensor(torch.tensor(idx.reshape(1, -1)), torch.tensor(dE.reshape(-1, D)), (V, D))
        opt.step()
        # a step moves a weight by <= lr = 1e-3; the 1.3e-5 in v (below) is 6.5e-6 of that through the square root
        np.testing.assert_allclose(W, p.detach().numpy(), rtol=2e-6, atol=1e-8 * step, err_msg=f"step {step}")
        sd = opt.state[p]
        np.testing.assert_allclose(st["m"], sd["exp_avg"].numpy(), rtol=2e-6, atol=1e-9)     # m + (g - m)(1 - b1) there: cancellation
        # v: the oracle forms 1 - beta_2 in fp32 as TF does (0.00099998713), torch in double (0.001): 1.3e-5 apart
        np.testing.assert_allclose(st["v"], sd["exp_avg_sq"].numpy(), rtol=3e-5, atol=1e-14)
    assert never.any() and np.array_equal(W[never], W0[never])        # the lazy rule: rows without a gradient stay bit for bit


def test_sparse_adagrad_and_sgd_against_torch():
    """A.4: acc += g^2 on the touched rows, var -= lr * g / (sqrt(acc) + eps) — torch.optim.Adagrad's sparse branch with the
    same initial accumulator and epsilon; and plain SGD."""
    import torch
    rng = np.random.default_rng(22)
    V, D = 50, 4
    W0 = rng.normal(0, 0.05, size=(V, D)).astype(np.float32)
    Wa, Ws = W0.copy(), W0.copy()
    st = dict(acc=np.full_like(W0, 0.1))
    pa, ps = torch.nn.Parameter(torch.tensor(W0)), torch.nn.Parameter(torch.tensor(W0))
    oa = torch.optim.Adagrad([pa], lr=1e-3, initial_accumulator_value=0.1, eps=1e-7)
    os_ = torch.optim.SGD([ps], lr=1e-2)
    for step in range(1, 5):
        idx = rng.integers(0, 30, size=(12, 2))
        dE = rng.normal(0, 1e-1, size=(12, 2, D)).astype(np.float32)
        O.sparse_backward_update(Wa, st, idx, dE, "adagrad", lr=1e-3, epsilon=1e-7)
        O.sparse_backward_update(Ws, {}, idx, dE, "sgd", lr=1e-2)
        g = torch.sparse_coo_tensor(torch.tensor(idx.reshape(1, -1)), torch.tensor(dE.reshape(-1, D)), (V, D))
        pa.grad, ps.grad = g, g.to_dense()
        oa.step()
        os_.step()
        np.testing.assert_allclose(Wa, pa.detach().numpy(), rtol=2e-6, atol=2e-9, err_msg=f"adagrad step {step}")
        np.testing.assert_allclose(st["acc"], oa.state[pa]["sum"].numpy(), rtol=2e-6)
        np.testing.assert_allclose(Ws, ps.detach().numpy(), rtol=2e-6, atol=2e-9, err_msg=f"sgd step {step}")


def test_losses_and_initialisers_against_torch():
    """Keras binary_crossentropy on probabilities (clip to [1e-7, 1 - 1e-7], log(p + 1e-7)) and on logits, against
    torch.nn.functional on float64 with autograd gradients; Glorot-uniform limit against torch.nn.init's."""
    import torch
    import torch.nn.functional as Fn
    rng = np.random.default_rng(23)
    n = 257
    prob = rng.uniform(0.01, 0.99, size=n).astype(np.float32)
    label = rng.integers(0, 2, size=n).astype(np.float32)
    edge = prob.copy()
    edge[:3] = (0.0, 1.0, 5e-8)                                      # outside the clip: zero gradient
    _, dp_edge = O.bce_clipped(edge, label)
    assert np.all(dp_edge[:3] == 0) and np.all(dp_edge[3:] != 0)     # Keras' clip_by_value passes no gradient outside the range
    loss, dp = O.bce_clipped(prob, label)
    pt = torch.tensor(prob, dtype=torch.float64, requires_grad=True)
    pc = pt.clamp(1e-7, 1 - 1e-7)
    y = torch.tensor(label, dtype=torch.float64)
    ref = -(y * torch.log(pc + 1e-7) + (1 - y) * torch.log(1 - pc + 1e-7)).mean()
    ref.backward()
    np.testing.assert_allclose(loss, ref.item(), rtol=2e-6)
    np.testing.assert_allclose(dp, pt.grad.numpy(), rtol=2e-5, atol=1e-9)
    # inside the clip it is torch's own BCE up to the +eps in the log
    np.testing.assert_allclose(loss, Fn.binary_cross_entropy(pc.detach(), y).item(), rtol=1e-5)
    logit = rng.normal(0, 3, size=n).astype(np.float32)
    l2, dl = O.bce_logits(logit, label)
    lt = torch.tensor(logit, dtype=torch.float64, requires_grad=True)
    r2 = Fn.binary_cross_entropy_with_logits(lt, y)
    r2.backward()
    np.testing.assert_allclose(l2, r2.item(), rtol=2e-6)
    np.testing.assert_allclose(dl, lt.grad.numpy(), rtol=2e-5, atol=1e-9)
    W = O.glorot_uniform(np.random.default_rng(0), 800, 512)
    t = torch.empty(512, 800)
    torch.manual_seed(0)
    torch.nn.init.xavier_uniform_(t)
    lim = np.sqrt(6.0 / (800 + 512))
    assert np.abs(W).max() <= lim and np.abs(W).max() > 0.99 * lim
    assert t.abs().max().item() <= lim * (1 + 1e-6) and t.abs().max().item() > 0.99 * lim


def test_dedup_order_against_pandas_factorize():
    """A.2: tf.unique returns the unique ids in FIRST-OCCURRENCE order; pandas.factorize has the same contract."""
    import pandas as pd
    rng = np.random.default_rng(9)
    ids = rng.integers(0, 40, size=500).astype(np.int64)
    vals = rng.normal(0, 1, size=(500, 4)).astype(np.float32)
    rows, summed = O.dedup_indexed_slices(ids, vals)
    codes, uniques = pd.factorize(ids)
    np.testing.assert_array_equal(rows, uniques)
    ref = np.zeros((uniques.size, 4), np.float32)
    for k in range(500):                                                          # unsorted_segment_sum, input order
        ref[codes[k]] += vals[k]
    np.testing.assert_array_equal(summed, ref)


def test_bf16_operand_mlp_mode_restates_the_cuda_rounding_points():
    """oracle.mlp_forward / mlp_backward(operand_dtype='bf16') (the restatement of csrc/mlp.cu's arithmetic used by the GPU
    parity tests): every GEMM operand is a bf16 value, hidden activations and travelling gradients are bf16 values, the last
    activation sees the fp32 accumulator, a Dense(1) head keeps its fp32 dz for dW / db; and the whole thing stays within
    bf16 rounding of the fp32 layers."""
    rng = np.random.default_rng(11)
    layers = O.init_mlp(rng, 24, [64, 32, 1])
    x = rng.normal(size=(50, 24)).astype(np.float32)
    y32, a32 = O.mlp_forward(x, layers, "sigmoid")
    y16, a16 = O.mlp_forward(x, layers, "sigmoid", "bf16")
    for a in a16[:-1]:
        np.testing.assert_array_equal(a, O.round_bf16(a))            # what the GEMMs read is representable in bf16
    assert not np.array_equal(a16[-1], O.round_bf16(a16[-1]))         # the output is the fp32 activation of the fp32 accumulator
    assert np.abs(y16 - y32).max() <= 2.0 ** -6
    dy = rng.normal(size=y32.shape).astype(np.float32)
    dx32, g32 = O.mlp_backward(dy, a32, layers, "sigmoid")
    dx16, g16 = O.mlp_backward(dy, a16, layers, "sigmoid", "bf16")
    np.testing.assert_array_equal(dx16, O.round_bf16(dx16))
    assert np.abs(dx16 - dx32).max() <= 2.0 ** -5 * np.abs(dx32).max()
    for (W16, b16), (W32, b32) in zip(g16, g32):
        assert np.abs(W16 - W32).max() <= 2.0 ** -5 * np.abs(W32).max()
        assert np.abs(b16 - b32).max() <= 2.0 ** -5 * np.abs(b32).max()
    # the head's bias gradient is the fp32 sum of dz, not of a rounded dz
    dz = dy * y16 * (1 - y16)
    np.testing.assert_allclose(g16[-1][1], dz.sum(0), rtol=1e-6)
    # DLRM / DeepFM accept the mode end to end
    p = O.init_dlrm(4, [32, 16], [32, 1], 16, 500)
    cat, dx, lab = O.synth_batch(32, 500, seed=4)
    prob, c = O.dlrm_forward(p, cat, dx, operand_dtype="bf16", mlp_dtype="bf16")
    ref, _ = O.dlrm_forward(p, cat, dx, operand_dtype="bf16")
    assert np.abs(prob - ref).max() <= 2.0 ** -6
    g = O.dlrm_backward(p, c, O.bce_clipped(prob, lab)[1], operand_dtype="bf16", mlp_dtype="bf16")
    assert g["dE"].shape == (32, 26, 16) and np.isfinite(g["dE"]).all()


def _din_case(g):
    his = O.compute_flat_embedding(g["W_item"], g["W_cat"], g["item"], g["cat"])
    tgt = O.compute_flat_embedding(g["W_item"], g["W_cat"], g["t_item"], g["t_cat"])[:, 0, :]
    layers = [(g[f"att_W{i}"], g[f"att_b{i}"]) for i in range(3)]
    return tgt, his, g["item"] != 0, layers


def test_din_local_activation_unit(golden):
    """dien/layers.py:34-59 as DIN.call wires it (dien/model.py:42-50): the oracle's forward and hand-derived backward against
    torch autograd of the reference's own class (tests/golden/make_golden.py::golden_din_attention)."""
    g = golden("din_attention")
    tgt, his, mask, layers = _din_case(g)
    rep, cache = O.local_activation_unit(tgt, his, mask, layers)
    np.testing.assert_allclose(rep, g["rep"], rtol=1e-5, atol=1e-7)
    dt, dh, grads = O.local_activation_unit_backward(cache, g["d_rep"])
    np.testing.assert_allclose(dt, g["d_target"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(dh, g["d_his"], rtol=1e-4, atol=1e-7)
    assert (dh[~mask] == 0).all()
    for i, (dW, db) in enumerate(grads):
        np.testing.assert_allclose(dW, g[f"att_dW{i}"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(db, g[f"att_db{i}"], rtol=1e-4, atol=1e-6)
    # table gradients: history rows + the target row, concatenated IndexedSlices summed per row
    D = g["W_item"].shape[1]
    for name, W, h_idx, t_idx, c0 in (("dW_item", g["W_item"], g["item"], g["t_item"], 0), ("dW_cat", g["W_cat"], g["cat"], g["t_cat"], D)):
        dense = np.zeros_like(W)
        np.add.at(dense, h_idx.reshape(-1), dh[:, :, c0:c0 + D].reshape(-1, D))
        np.add.at(dense, t_idx.reshape(-1), dt[:, c0:c0 + D])
        np.testing.assert_allclose(dense, g[name], rtol=1e-4, atol=1e-7)


def test_din_bf16_operand_mode_stays_near_fp32(golden):
    """The rounding points the CUDA path uses (bf16 feature rows, kernels and hidden activations; fp32 accumulation, logit and
    weighted sum) move the result by bf16-level amounts only."""
    g = golden("din_attention")
    tgt, his, mask, layers = _din_case(g)
    rep16, cache = O.local_activation_unit(tgt, his, mask, layers, operand_dtype="bf16")
    assert np.abs(rep16 - g["rep"]).max() <= 3e-2 * np.abs(g["rep"]).max()
    dt, dh, grads = O.local_activation_unit_backward(cache, g["d_rep"])
    assert np.abs(dh - g["d_his"]).max() <= 6e-2 * np.abs(g["d_his"]).max()
    assert np.abs(dt - g["d_target"]).max() <= 5e-2 * np.abs(g["d_target"]).max()
