"""Host-side Python logic that needs no GPU: the dense (MLP) mirror, the Keras-form dense
optimizers, loss forms and shape helpers — each against the oracle."""
import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O
from recommender_b200 import ops
from recommender_b200.layers import MLP
from recommender_b200.model import bce_clipped, bce_logits
from recommender_b200.optimizers import Adagrad, Adam, SGD


def test_interaction_ncols():
    assert ops.interaction_ncols(27, False, True) == 729
    assert ops.interaction_ncols(27, True, True) == 729
    assert ops.interaction_ncols(27, False, False) == 351
    assert ops.interaction_ncols(27, True, False) == 378


@pytest.mark.parametrize("act", [None, "relu", "sigmoid"])
def test_mlp_matches_oracle_hidden_layers_are_linear(act):
    rng = np.random.default_rng(0)
    layers = O.init_mlp(rng, 13, [32, 16, 4])
    x = rng.normal(size=(9, 13)).astype(np.float32)
    mlp = MLP([32, 16, 4], act)
    mlp.load_arrays(layers, "cpu")
    y = mlp(torch.tensor(x)).detach().numpy()
    ref, _ = O.mlp_forward(x, layers, act)
    np.testing.assert_allclose(y, ref, rtol=1e-5, atol=1e-6)


def test_mlp_lazy_build_uses_glorot_uniform():
    g = torch.Generator().manual_seed(0)
    mlp = MLP([64, 8], "relu", generator=g)
    mlp(torch.zeros(2, 20))
    lim = np.sqrt(6.0 / (20 + 64))
    W = mlp.kernels[0].detach()
    assert W.shape == (20, 64) and float(W.abs().max()) <= lim and float(W.abs().max()) > 0.8 * lim
    assert float(mlp.biases[0].abs().max()) == 0.0


def test_dense_adam_is_the_keras_formula(cuda_lib, fake_kernels):
    rng = np.random.default_rng(1)
    p0 = rng.normal(size=(6, 5)).astype(np.float32)
    p = torch.nn.Parameter(torch.tensor(p0.copy()))
    ref, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    opt = Adam()
    for step in range(1, 4):
        g = rng.normal(size=p0.shape).astype(np.float32)
        p.grad = torch.tensor(g)
        opt.apply_gradients([p])
        O.adam_dense_param(ref, m, v, g, step)
        np.testing.assert_allclose(p.detach().numpy(), ref, rtol=2e-6, atol=1e-7)
    assert opt.iterations == 3 and p.grad is None


def test_dense_adagrad_and_sgd(cuda_lib, fake_kernels):
    p0 = np.ones((3, 2), np.float32)
    g = np.full((3, 2), 0.5, np.float32)
    p = torch.nn.Parameter(torch.tensor(p0.copy()))
    p.grad = torch.tensor(g)
    Adagrad().apply_gradients([p])
    acc = 0.1 + g * g
    np.testing.assert_allclose(p.detach().numpy(), p0 - 1e-3 * g / (np.sqrt(acc) + 1e-7), rtol=1e-6)
    q = torch.nn.Parameter(torch.tensor(p0.copy()))
    q.grad = torch.tensor(g)
    SGD(0.1).apply_gradients([q])
    np.testing.assert_allclose(q.detach().numpy(), p0 - 0.1 * g, rtol=1e-6)


def test_loss_forms_match_oracle(fake_kernels):
    rng = np.random.default_rng(2)
    prob = rng.uniform(0, 1, size=64).astype(np.float32)
    prob[:2] = [0.0, 1.0]
    y = (rng.random(64) < 0.3).astype(np.int64)
    ref, dref = O.bce_clipped(prob, y)
    pt = torch.tensor(prob, requires_grad=True)
    loss = bce_clipped(pt, torch.tensor(y))
    loss.backward()
    np.testing.assert_allclose(loss.item(), ref, rtol=1e-5)
    np.testing.assert_allclose(pt.grad.numpy(), dref, rtol=1e-4, atol=1e-7)
    logit = rng.normal(0, 3, size=64).astype(np.float32)
    ref, dref = O.bce_logits(logit, y)
    lt = torch.tensor(logit, requires_grad=True)
    loss = bce_logits(lt, torch.tensor(y))
    loss.backward()
    np.testing.assert_allclose(loss.item(), ref, rtol=1e-5)
    np.testing.assert_allclose(lt.grad.numpy(), dref, rtol=1e-4, atol=1e-7)


def test_embedding_needs_cuda():
    from recommender_b200.layers import Embedding
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        Embedding(10, 4)


def test_mlp_bf16_path_forward_backward_close_to_fp32(fake_kernels):
    rng = np.random.default_rng(5)
    layers = O.init_mlp(rng, 13, [64, 32, 8])
    x = rng.normal(size=(128, 13)).astype(np.float32)
    ref, acts = O.mlp_forward(x, layers, "relu")
    dy = rng.normal(size=ref.shape).astype(np.float32)
    _, ref_grads = O.mlp_backward(dy, acts, layers, "relu")
    mlp = MLP([64, 32, 8], "relu", compute_dtype=torch.bfloat16)
    mlp.load_arrays(layers, "cpu")
    assert mlp.padded_in_dim() == 16
    y = mlp(torch.tensor(x))
    assert y.dtype == torch.float32
    np.testing.assert_allclose(y.detach().numpy(), ref, rtol=0, atol=0.03 * np.abs(ref).max())
    y.backward(torch.tensor(dy))
    for W, b, (dW, db) in zip(mlp.kernels, mlp.biases, ref_grads):
        assert W.grad.shape == W.shape and W.grad.dtype == torch.float32
        assert np.abs(W.grad.numpy() - dW).max() <= 0.05 * np.abs(dW).max()
        assert np.abs(b.grad.numpy() - db).max() <= 0.05 * np.abs(db).max()
    # pre-padded bf16 input (what the fused interaction kernel emits) gives the same result
    xp = torch.zeros(128, 16, dtype=torch.bfloat16)
    xp[:, :13] = torch.tensor(x)
    y2 = mlp(xp)
    np.testing.assert_array_equal(y2.detach().numpy(), y.detach().numpy())


def test_prepare_step_advances_the_counter_exactly_once(fake_kernels):
    """graph.GraphedTrainStep calls Adam.prepare_step() before every replay and apply_gradients() inside the captured
    body: together they must advance `iterations` by one, and the dense update must use that step's alpha_t."""
    import torch
    from recommender_b200.optimizers import Adam
    rng = np.random.default_rng(0)
    w0 = rng.normal(size=(4, 3)).astype(np.float32)
    g = rng.normal(size=(4, 3)).astype(np.float32)
    a, b = torch.nn.Parameter(torch.tensor(w0)), torch.nn.Parameter(torch.tensor(w0))
    opt_a, opt_b = Adam(), Adam()
    for _ in range(3):
        a.grad = torch.tensor(g)
        opt_a.apply_gradients([a])                     # plain: the call advances the counter itself
        b.grad = torch.tensor(g)
        opt_b.prepare_step()                           # graph style: host half first ...
        opt_b.apply_gradients([b])                     # ... then the (captured) device half
    assert opt_a.iterations == opt_b.iterations == 3
    assert torch.equal(a.detach(), b.detach())
    ref, m, v = w0.copy(), np.zeros_like(w0), np.zeros_like(w0)
    for t in (1, 2, 3):
        O.adam_dense_param(ref, m, v, g, t)
    np.testing.assert_allclose(a.detach().numpy(), ref, rtol=1e-6, atol=1e-7)


def test_keras_style_auc_and_accuracy():
    """recommender_b200.train.AUC / BinaryAccuracy restate tf.keras.metrics.AUC() / BinaryAccuracy() with their defaults
    (ctr/train.py:86): 200 thresholds, confusion matrices accumulated over batches, trapezoids over the ROC points."""
    from recommender_b200.train import AUC, BinaryAccuracy
    rng = np.random.default_rng(0)
    y = (rng.random(5000) < 0.3).astype(np.int64)
    p = np.clip(0.3 * y + rng.random(5000) * 0.7, 0, 1).astype(np.float32)
    p[:5] = [0.0, 1.0, 0.5, 1 / 199, 198 / 199]                                   # values that sit ON thresholds
    auc, acc = AUC(), BinaryAccuracy()
    for s in range(0, 5000, 777):
        auc.update_state(torch.from_numpy(y[s:s + 777]), torch.from_numpy(p[s:s + 777]))
        acc.update_state(torch.from_numpy(y[s:s + 777]), torch.from_numpy(p[s:s + 777]))
    th = np.array([-1e-7] + [(i + 1) / 199 for i in range(198)] + [1 + 1e-7], dtype=np.float32)
    tpr = np.array([((p > t) & (y == 1)).sum() for t in th], float) / (y == 1).sum()
    fpr = np.array([((p > t) & (y == 0)).sum() for t in th], float) / (y == 0).sum()
    ref = ((fpr[:-1] - fpr[1:]) * (tpr[:-1] + tpr[1:]) / 2).sum()
    assert abs(auc.result() - ref) < 1e-12
    order = np.argsort(p, kind="stable")
    ranks = np.empty(5000)
    ranks[order] = np.arange(1, 5001)
    n1 = y.sum()
    exact = (ranks[y == 1].sum() - n1 * (n1 + 1) / 2) / (n1 * (5000 - n1))
    assert abs(auc.result() - exact) < 5e-3                                       # the 200-threshold curve is close to the exact AUC
    assert acc.result() == ((p > 0.5) == (y > 0)).mean()
    assert AUC().result() == 0.0                                                  # no samples: div_no_nan


def test_adam_device_scalar_ring_never_reuses_a_slot_in_flight(fake_kernels):
    """optimizers.Adam._refresh_device_scalars: alpha_t of step k is staged in pinned slot k mod 4096 before its
    asynchronous copy to the device, so a replay loop running many steps ahead of the GPU cannot overwrite the value a
    still-queued copy is going to read (emulated here with plain host tensors)."""
    opt = Adam()
    opt._alpha_dev = torch.zeros(1)
    opt._alpha_host = torch.zeros(opt._ALPHA_SLOTS)
    seen = {}
    for k in range(1, opt._ALPHA_SLOTS + 50):
        opt.iterations = k
        opt._refresh_device_scalars()
        slot = k % opt._ALPHA_SLOTS
        assert float(opt._alpha_dev) == float(opt._alpha_host[slot]) == ops.adam_alpha_t(1e-3, 0.9, 0.999, k)
        seen[slot] = k
    assert len(seen) == opt._ALPHA_SLOTS                                          # every slot used before any is reused
    assert float(opt._alpha_host[60]) == ops.adam_alpha_t(1e-3, 0.9, 0.999, 60)    # untouched for 4096 steps
    assert float(opt._alpha_host[5]) == ops.adam_alpha_t(1e-3, 0.9, 0.999, opt._ALPHA_SLOTS + 5)   # reused one lap later


@pytest.mark.parametrize("units,act,in_dim", [([512, 256, 64], "relu", 13), ([512, 256, 1], "sigmoid", 793), ([32, 1], None, 429),
                                              ([24, 16, 8, 4], "relu", 10)])
def test_collapsed_affine_mlp_matches_the_layerwise_oracle(units, act, in_dim, fake_kernels):
    """MLP(collapse_linear=True): the linear hidden layers (ctr/layers.py:8) make each tower ONE affine map followed by
    the last activation; output, input gradient and every layer's own dW / db must equal the oracle's layer-by-layer
    forward / backward to fp32 re-association accuracy."""
    rng = np.random.default_rng(in_dim)
    layers = O.init_mlp(rng, in_dim, units)
    layers = [(W, rng.normal(0, 0.1, size=b.shape).astype(np.float32)) for W, b in layers]       # non-zero biases
    x = rng.normal(0, 0.5, size=(129, in_dim)).astype(np.float32)
    dy = rng.normal(size=(129, units[-1])).astype(np.float32)
    y_ref, acts = O.mlp_forward(x, layers, act)
    dx_ref, grads_ref = O.mlp_backward(dy, acts, layers, act)
    mlp = MLP(units, act, collapse_linear=True)
    mlp.load_arrays(layers, "cpu")
    xt = torch.tensor(x, requires_grad=True)
    y = mlp(xt)
    (y * torch.tensor(dy)).sum().backward()

    def close(a, b):
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-5 * max(np.abs(b).max(), 1e-6))
    close(y.detach().numpy(), y_ref)
    close(xt.grad.numpy(), dx_ref)
    for W, b, (dW, db) in zip(mlp.kernels, mlp.biases, grads_ref):
        close(W.grad.numpy(), dW)
        close(b.grad.numpy(), db)


def test_collapsed_affine_mlp_bf16_path_with_padding_and_ones_column(fake_kernels):
    """The bf16 form the fused interaction kernel feeds: input padded to a multiple of 8 columns, pad column in_dim = 1.
    The padded rows of the collapsed kernel are zero, and the bias-gradient vector s is read off row in_dim of x^T dz.
    No worse than the layer-by-layer bf16 path against the fp32 oracle."""
    rng = np.random.default_rng(7)
    in_dim, units = 13, [64, 32, 8]
    layers = O.init_mlp(rng, in_dim, units)
    layers = [(W, rng.normal(0, 0.1, size=b.shape).astype(np.float32)) for W, b in layers]
    x = rng.normal(0, 0.5, size=(200, in_dim)).astype(np.float32)
    dy = rng.normal(size=(200, units[-1])).astype(np.float32)
    y_ref, acts = O.mlp_forward(x, layers, "relu")
    _, grads_ref = O.mlp_backward(dy, acts, layers, "relu")
    errs = {}
    for collapse in (False, True):
        mlp = MLP(units, "relu", compute_dtype=torch.bfloat16, collapse_linear=collapse)
        mlp.load_arrays(layers, "cpu")
        y = mlp(torch.tensor(x))
        (y * torch.tensor(dy)).sum().backward()
        e = [np.abs(y.detach().numpy() - y_ref).max() / np.abs(y_ref).max()]
        for W, b, (dW, db) in zip(mlp.kernels, mlp.biases, grads_ref):
            e.append(np.abs(W.grad.numpy() - dW).max() / np.abs(dW).max())
            e.append(np.abs(b.grad.numpy() - db).max() / np.abs(db).max())
        errs[collapse] = max(e)
    assert errs[True] < 0.05 and errs[True] <= 1.5 * errs[False]
