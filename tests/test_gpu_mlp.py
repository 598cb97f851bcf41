"""The Dense layers of the towers (ctr/layers.py:5-14) on the tcgen05 kernels (csrc/mlp.cu), through the C ABI:
every product against fp32 matmuls of the same bf16 operands, the MLP module against the oracle's bf16-operand MLP
(oracle.ctr_oracle.mlp_forward / mlp_backward, which restate the kernels' rounding points) and against the fp32
reference layers, and the exact configuration bench.py times (fused DLRM, bf16 towers) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O

pytestmark = pytest.mark.gpu

BF16_EPS = 2.0 ** -8          # half an ulp of bf16, relative: the rounding of a bf16 output


def cu(a):
    return torch.as_tensor(a).cuda()


def _rnd(g, *shape, scale=1.0):
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def _close(got, ref, tol):
    got, ref = got.float(), ref.float()
    assert not torch.isnan(got).any()
    assert (got - ref).abs().max().item() <= tol * ref.abs().max().item() + 1e-30


@pytest.mark.parametrize("rows,in_dim,units", [(128, 64, 64), (256, 128, 256), (300, 800, 512), (1000, 16, 512), (4096, 512, 256),
                                               (130, 24, 40), (65, 832, 264)])
def test_dense_forward_matches_fp32_matmul_of_the_bf16_operands(cuda_lib, rows, in_dim, units):
    from recommender_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + in_dim)
    x, w = _rnd(g, rows, in_dim), _rnd(g, in_dim, units, scale=in_dim ** -0.5)
    b = torch.randn(units, generator=g, device="cuda")
    ref = x.float() @ w.float() + b
    _close(ops.dense_fwd(x, w, b, None, torch.bfloat16), ref, 2 * BF16_EPS)            # bf16 output rounding
    _close(ops.dense_fwd(x, w, b, None, torch.float32), ref, 1e-5)                     # fp32 accumulation order only
    _close(ops.dense_fwd(x, w, b, "relu", torch.float32), ref.relu(), 1e-5)
    _close(ops.dense_fwd(x, w, None, "sigmoid", torch.float32), (x.float() @ w.float()).sigmoid(), 1e-5)


@pytest.mark.parametrize("rows,in_dim,units", [(128, 64, 64), (256, 256, 128), (300, 800, 512), (1000, 512, 256), (77, 16, 512),
                                               (4096, 256, 64)])
def test_dense_input_gradient(cuda_lib, rows, in_dim, units):
    from recommender_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + units)
    dy, w = _rnd(g, rows, units), _rnd(g, in_dim, units, scale=units ** -0.5)
    _close(ops.dense_bwd_input(dy, w), dy.float() @ w.float().t(), 2 * BF16_EPS)


@pytest.mark.parametrize("rows,in_dim,units", [(128, 128, 64), (256, 128, 256), (300, 800, 512), (1000, 16, 512), (100, 40, 24),
                                               (65536, 256, 64)])
def test_dense_weight_gradient_is_exact_to_fp32_and_deterministic(cuda_lib, rows, in_dim, units):
    from recommender_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + in_dim + units)
    x, dy = _rnd(g, rows, in_dim), _rnd(g, rows, units, scale=rows ** -0.5)
    dw = ops.dense_bwd_weight(x, dy)
    _close(dw, (x.double().t() @ dy.double()).float(), 1e-5)
    assert torch.equal(dw, ops.dense_bwd_weight(x, dy))        # split over the batch, summed in split order: no atomics


@pytest.mark.parametrize("act", [None, "relu", "sigmoid"])
def test_dense_head_forward_and_backward(cuda_lib, act):
    from recommender_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    rows, in_dim = 1000, 256
    x, w = _rnd(g, rows, in_dim), _rnd(g, in_dim, scale=in_dim ** -0.5)
    b = torch.randn(1, generator=g, device="cuda")
    out = ops.dense_head_fwd(x, w, b, act)
    z = x.float() @ w.float() + b
    ref = z.sigmoid() if act == "sigmoid" else z.relu() if act == "relu" else z
    _close(out, ref, 1e-5)
    dout = torch.randn(rows, generator=g, device="cuda")
    dx, dw, db = ops.dense_head_bwd(dout, out, act, x, w)
    dz = dout * (ref * (1 - ref) if act == "sigmoid" else (ref > 0).float() if act == "relu" else 1.0)
    _close(dx, dz[:, None] * w.float()[None], 2 * BF16_EPS)
    _close(dw, x.float().t() @ dz, 1e-5)
    _close(db, dz.sum().reshape(1), 1e-5)


def test_pack_input_and_activation_backward(cuda_lib):
    from recommender_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(300, 13, generator=g, device="cuda")
    p = ops.dense_pack_input(x, 16, True)
    assert torch.equal(p[:, :13], x.to(torch.bfloat16)) and (p[:, 13] == 1).all() and (p[:, 14:] == 0).all()
    dy, y = torch.randn(300, 64, generator=g, device="cuda"), torch.randn(300, 64, generator=g, device="cuda").relu()
    assert torch.equal(ops.dense_act_bwd(dy, y, "relu"), (dy * (y > 0)).to(torch.bfloat16))
    s = torch.rand(300, 64, generator=g, device="cuda")
    assert torch.equal(ops.dense_act_bwd(dy, s, "sigmoid"), (dy * s * (1 - s)).to(torch.bfloat16))


def _oracle_layers(rng, in_dim, units):
    return O.init_mlp(rng, in_dim, units)


@pytest.mark.parametrize("units,act,in_dim", [([512, 256, 1], "sigmoid", 793), ([512, 256, 64], "relu", 13), ([64, 32, 1], None, 429),
                                              ([32], "relu", 40)])
def test_mlp_module_matches_the_bf16_oracle_and_the_cublas_path(cuda_lib, units, act, in_dim):
    """MLP(compute_dtype=bf16) on the tcgen05 kernels vs (a) the oracle restating its rounding points: outputs to fp32
    accumulation order, gradients to one bf16 ulp of the travelling activation gradients; (b) the same module on
    torch / cuBLASLt (backend='cublas'); (c) the fp32 layers, at the bf16 tolerance."""
    from recommender_b200.layers import MLP
    rng = np.random.default_rng(5)
    layers = _oracle_layers(rng, in_dim, units)
    layers = [(W, rng.normal(0, 0.1, b.shape).astype(np.float32)) for W, b in layers]       # non-zero biases
    B = 300
    x = rng.normal(0, 1, (B, in_dim)).astype(np.float32)
    dy = rng.normal(0, 1e-2, (B, units[-1])).astype(np.float32)

    def run(backend, dtype):
        m = MLP(units, act, compute_dtype=dtype)
        m.backend = backend
        m.load_arrays(layers, "cuda")
        xt = cu(x).requires_grad_(True)
        y = m(xt)
        y.backward(cu(dy))
        return (y.detach().cpu().numpy(), xt.grad.cpu().numpy(), [k.grad.cpu().numpy() for k in m.kernels],
                [b.grad.cpu().numpy() for b in m.biases])

    y, dx, dWs, dbs = run("tcgen05", torch.bfloat16)
    ry, acts = O.mlp_forward(x, layers, act, "bf16")
    rdx, rg = O.mlp_backward(dy.copy(), acts, layers, act, "bf16")
    # a hidden activation that sits on a bf16 rounding boundary may round the other way when the fp32 sums are associated
    # differently: one such flip moves an output by ~ |w| |h| 2^-8
    np.testing.assert_allclose(y, ry, rtol=0, atol=2 * BF16_EPS * max(1.0, np.abs(ry).max()))
    tol = 4 * BF16_EPS
    assert np.abs(dx - rdx).max() <= tol * np.abs(rdx).max()
    for (rW, rb), dW, db in zip(rg, dWs, dbs):
        assert np.abs(dW - rW).max() <= tol * np.abs(rW).max() + 1e-9
        assert np.abs(db - rb).max() <= tol * np.abs(rb).max() + 1e-9
    y2, dx2, dWs2, _ = run("cublas", torch.bfloat16)
    assert np.abs(y - y2).max() <= 4 * BF16_EPS * max(1.0, np.abs(y2).max())
    # (a relu unit whose pre-activation is a rounding error away from 0 may switch between the two paths: mean, not max)
    assert np.abs(dx - dx2).mean() <= 2 * BF16_EPS * np.abs(dx2).max()
    for a, b in zip(dWs, dWs2):
        assert np.abs(a - b).mean() <= 2 * BF16_EPS * np.abs(b).max() + 1e-9
    y3, dx3, dWs3, _ = run("tcgen05", None)                    # fp32 layers (torch fp32)
    assert np.abs(y - y3).max() <= 8 * BF16_EPS * max(1.0, np.abs(y3).max())
    for a, b in zip(dWs, dWs3):
        assert np.abs(a - b).mean() <= 4 * BF16_EPS * np.abs(b).max() + 1e-9


def _mlp(g, name):
    layers, i = [], 0
    while f"{name}_W{i}" in g:
        layers.append((g[f"{name}_W{i}"], g[f"{name}_b{i}"]))
        i += 1
    return layers


def _build_bench_dlrm(g):
    """The model bench.py times: fused lookup + interaction emitting the bf16 row with its ones column, bf16 towers."""
    from recommender_b200.model import DLRM
    bottom, top = _mlp(g, "bottom"), _mlp(g, "top")
    D, V = g["table"].shape[1], g["table"].shape[0]
    model = DLRM([W.shape[1] for W, _ in bottom], [W.shape[1] for W, _ in top], D, V, 26, 13, fused=True, device="cuda",
                 compute_dtype=torch.bfloat16)
    model.embedding_layer.embeddings.copy_(cu(g["table"]))
    model.bottom_mlp.load_arrays(bottom, "cuda")
    model.top_mlp.load_arrays(top, "cuda")
    return model


def _oracle_train(g, batches, steps):
    params = dict(table=g["table"].copy(), bottom=[(W.copy(), b.copy()) for W, b in _mlp(g, "bottom")],
                  top=[(W.copy(), b.copy()) for W, b in _mlp(g, "top")])
    st = dict(m=np.zeros_like(params["table"]), v=np.zeros_like(params["table"]))
    dense_state, out = {}, []
    for step in range(1, steps + 1):
        cat, dense_x, label = batches[step - 1]
        prob, cache = O.dlrm_forward(params, cat, dense_x, operand_dtype="bf16", mlp_dtype="bf16")
        loss, dprob = O.bce_clipped(prob, label)
        grads = O.dlrm_backward(params, cache, dprob, operand_dtype="bf16", mlp_dtype="bf16")
        out.append((prob, float(loss), grads))
        O.sparse_backward_update(params["table"], st, cat, grads["dE"], "adam_lazy", step)
        for name in ("bottom", "top"):
            for i, ((W, b), (dW, db)) in enumerate(zip(params[name], grads[name])):
                for tag, p, gr in (("W", W, dW), ("b", b, db)):
                    m, v = dense_state.setdefault((name, i, tag), (np.zeros_like(p), np.zeros_like(p)))
                    O.adam_dense_param(p, m, v, gr, step)
    return params, st, out


@pytest.mark.parametrize("graphed", [False, True])
def test_the_benchmarked_dlrm_tracks_the_oracle(cuda_lib, golden, graphed):
    """bench.py's model (fused interaction -> bf16 row + ones column -> tcgen05 towers, lazy Adam) against the oracle with
    the same rounding points, step 1 in detail (prob, loss, every dense gradient, the updated table) and 5 steps of training,
    launched eagerly and as one CUDA graph.  Stated tolerance: 4 bf16 ulps (2^-6) of each tensor's largest entry for
    gradients that travelled through bf16 GEMMs; 5e-4 per step for probabilities and loss."""
    from recommender_b200.graph import GraphedTrainStep
    from recommender_b200.model import bce_clipped
    from recommender_b200.optimizers import Adam
    g = golden("dlrm_uniform")
    V = g["table"].shape[0]
    steps = 5
    batches = [(g["cat"], g["dense"], g["label"])] + [O.synth_batch(g["cat"].shape[0], V, seed=100 + s, dist="zipf") for s in range(1, steps)]
    params, st, ref = _oracle_train(g, batches, steps)

    model = _build_bench_dlrm(g)
    opt = Adam()
    dev = [tuple(cu(t) for t in b) for b in batches]
    if not graphed:
        # step 1 in detail
        prob = model({"cat_features": dev[0][0], "int_features": dev[0][1]})
        loss = bce_clipped(prob, dev[0][2])
        loss.backward()
        rprob, rloss, rg = ref[0]
        np.testing.assert_allclose(prob.detach().cpu().numpy(), rprob, rtol=0, atol=5e-4)
        assert abs(loss.item() - rloss) <= 5e-4
        tol = 2.0 ** -6
        for name, mlp in (("bottom", model.bottom_mlp), ("top", model.top_mlp)):
            for i, (W, b) in enumerate(zip(mlp.kernels, mlp.biases)):
                rW, rb = rg[name][i]
                assert np.abs(W.grad.cpu().numpy() - rW).max() <= tol * np.abs(rW).max() + 1e-9, (name, i)
                assert np.abs(b.grad.cpu().numpy() - rb).max() <= tol * np.abs(rb).max() + 1e-9, (name, i)
        opt.apply_gradients(model)
        for s in range(1, steps):
            prob = model({"cat_features": dev[s][0], "int_features": dev[s][1]})
            loss = bce_clipped(prob, dev[s][2])
            loss.backward()
            opt.apply_gradients(model)
            np.testing.assert_allclose(prob.detach().cpu().numpy(), ref[s][0], rtol=0, atol=5e-4 * (s + 1))
            assert abs(loss.item() - ref[s][1]) <= 5e-4 * (s + 1)
    else:
        # GraphedTrainStep runs its warm-up steps on the first batch: give it batch 0 `done` times in the oracle's order instead
        gs = GraphedTrainStep(model, opt, bce_clipped, dev[0], warmup=1)
        done = gs.steps_run
        batches_g = [batches[0]] * done + batches[1:steps - done + 1]
        params, st, ref = _oracle_train(g, batches_g, len(batches_g))
        for s in range(done, len(batches_g)):
            loss = gs.step(dev[s - done + 1])
            assert abs(float(loss.item()) - ref[s][1]) <= 5e-4 * (s + 1)
    torch.cuda.synchronize()
    got = model.embedding_layer.embeddings.cpu().numpy()
    moved = np.abs(params["table"] - g["table"]).max()
    assert moved > 1e-3
    assert np.abs(got - params["table"]).mean() <= 0.02 * moved
    for name, mlp in (("bottom", model.bottom_mlp), ("top", model.top_mlp)):
        for i, W in enumerate(mlp.kernels):
            rW = params[name][i][0]
            assert np.abs(W.detach().cpu().numpy() - rW).mean() <= 0.02 * np.abs(rW - _mlp(g, name)[i][0]).max() + 1e-7


def test_deepfm_with_the_fused_deep_input_row_tracks_the_oracle(cuda_lib, golden):
    """BASELINE config 1's model in the bf16 configuration: rb_gather_fm_deep_fwd writes [flatten(E) | int | 1 | 0...] as the
    MLP's bf16 K operand (ctr/model.py:19-27 without E, reshape, concat or pad copy); prob, loss, the MLP gradients and the
    table gradient (three consumers of E: MLP, sum-square, square-sum) against the oracle with the same rounding points."""
    from recommender_b200.model import DeepFM, bce_logits
    from recommender_b200.optimizers import SGD
    g = golden("deepfm_small")
    mlp = _mlp(g, "mlp")
    D, V = g["table"].shape[1], g["table"].shape[0]
    params = dict(table=g["table"].copy(), mlp=mlp)
    rprob, cache = O.deepfm_forward(params, g["cat"], g["dense"], mlp_dtype="bf16")
    rloss, dlogit = O.bce_logits(cache["logit"], g["label"])
    rg = O.deepfm_backward(params, cache, dlogit, mlp_dtype="bf16")
    model = DeepFM(D, V, 13, 26, [W.shape[1] for W, _ in mlp], fused=True, device="cuda", compute_dtype=torch.bfloat16)
    model.embedding_layer.embeddings.copy_(cu(g["table"]))
    model.mlp.load_arrays(mlp, "cuda")
    x = {"cat_features": cu(g["cat"]), "int_features": cu(g["dense"])}
    logit = model.logits(x)
    np.testing.assert_allclose(torch.sigmoid(logit).detach().cpu().numpy(), rprob, rtol=0, atol=2e-3)
    np.testing.assert_allclose(torch.sigmoid(logit).detach().cpu().numpy(), g["prob"], rtol=0, atol=2e-2)      # the fp32 reference run
    loss = bce_logits(logit, cu(g["label"]))
    assert abs(loss.item() - float(rloss)) <= 2e-3
    loss.backward()
    tol = 2.0 ** -6
    for i, (W, b) in enumerate(zip(model.mlp.kernels, model.mlp.biases)):
        rW, rb = rg["mlp"][i]
        assert np.abs(W.grad.cpu().numpy() - rW).max() <= tol * np.abs(rW).max() + 1e-9, i
        assert np.abs(b.grad.cpu().numpy() - rb).max() <= tol * np.abs(rb).max() + 1e-9, i
    SGD(1.0).apply_gradients([model.embedding_layer])
    dtable = g["table"] - model.embedding_layer.embeddings.cpu().numpy()
    ref = np.zeros_like(g["table"])
    np.add.at(ref, g["cat"].reshape(-1), rg["dE"].reshape(-1, D))
    assert np.abs(dtable - ref).max() <= tol * np.abs(ref).max() + 1e-9


@pytest.mark.parametrize("in_dim,rows,ltype", [(256, 65536, torch.float32), (256, 1000, torch.int64), (64, 777, torch.float32), (8, 33, torch.int64)])
def test_head_loss_and_head_backward_in_one_kernel(cuda_lib, in_dim, rows, ltype):
    """rb_dense_head_bce against the three calls it replaces (rb_dense_head_fwd -> rb_bce_clipped -> rb_dense_head_bwd): the same
    bits for prob, dx, dW, db and the dx column sums; the loss to the order its terms are added."""
    from recommender_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(in_dim + rows)
    x = (torch.randn(rows, in_dim, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(in_dim, generator=g, device="cuda") * in_dim ** -0.5).to(torch.bfloat16)
    b = torch.randn(1, generator=g, device="cuda")
    label = (torch.rand(rows, generator=g, device="cuda") < 0.3).to(ltype)
    prob = ops.dense_head_fwd(x, w, b, "sigmoid")
    loss, dprob = ops.bce_clipped(prob, label, want_grad=True)
    dx, dw, db, cs = ops.dense_head_bwd(dprob, prob, "sigmoid", x, w, want_dx=True, want_dx_colsum=True)
    p2, l2, dx2, dw2, db2, cs2 = ops.dense_head_bce(x, w, b, label, want_dx=True, want_dx_colsum=True)
    if in_dim > 64:      # narrower heads: rb_dense_head_fwd adds the row's products in another order (one thread per row)
        assert torch.equal(p2, prob)
        assert torch.equal(dx2, dx) and torch.equal(dw2, dw) and torch.equal(db2, db) and torch.equal(cs2, cs)
    else:
        torch.testing.assert_close(p2, prob, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(dx2.float(), dx.float(), rtol=2e-2, atol=1e-9)
        torch.testing.assert_close(dw2, dw, rtol=1e-4, atol=1e-8)
        torch.testing.assert_close(db2, db, rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(l2.reshape(()), loss.reshape(()), rtol=2e-6, atol=0)


def test_dlrm_forward_bce_equals_model_then_loss(cuda_lib):
    """DLRM.forward_bce (what GraphedTrainStep runs for bce_clipped) against bce_clipped(model(inputs), label): identical parameter
    gradients, identical embedding gradient rows, the same loss."""
    from recommender_b200.model import DLRM, bce_clipped
    B, D, V = 512, 64, 3000

    def run(fused):
        gen = torch.Generator(device="cuda").manual_seed(7)
        model = DLRM([64, 32, D], [128, 256, 1], D, V, 26, 13, num_tables=26, device="cuda", compute_dtype=torch.bfloat16, generator=gen)
        g2 = torch.Generator(device="cuda").manual_seed(8)
        inputs = {"cat_features": torch.randint(0, V, (B, 26), device="cuda", generator=g2),
                  "int_features": torch.rand(B, 13, device="cuda", generator=g2)}
        label = (torch.rand(B, device="cuda", generator=g2) < 0.25).to(torch.int64)
        loss = model.forward_bce(inputs, label) if fused else bce_clipped(model(inputs), label)
        loss.backward()
        torch.cuda.synchronize()
        grads = [p.grad.clone() for p in model.parameters()]
        dE = model.embedding_layer.pending[0].grad.srcs[0].clone()
        return loss.detach().clone(), grads, dE, model

    l0, g0, e0, _ = run(False)
    l1, g1, e1, m1 = run(True)
    torch.testing.assert_close(l1, l0, rtol=2e-6, atol=0)
    assert torch.equal(e0, e1)
    for a, b in zip(g0, g1):
        assert torch.equal(a, b)
    assert m1.top_mlp.last_prob is not None and m1.top_mlp.last_prob.shape == (B,)
