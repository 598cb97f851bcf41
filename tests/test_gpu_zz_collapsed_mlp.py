"""MLP(collapse_linear=True) on the GPU (opt-in, layers._CollapsedAffineFn): the towers' linear hidden layers
(ctr/layers.py:8) evaluated as one affine map, against the layer-by-layer path and the reference fixture.
Runs last (file name): the default path does not depend on it."""
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


@pytest.mark.parametrize("units,act,in_dim", [([512, 256, 64], "relu", 13), ([512, 256, 1], "sigmoid", 793)])
def test_collapsed_mlp_fp32_equals_layerwise_oracle(cuda_lib, units, act, in_dim):
    from recommender_b200.layers import MLP
    rng = np.random.default_rng(in_dim)
    layers = O.init_mlp(rng, in_dim, units)
    layers = [(W, rng.normal(0, 0.1, size=b.shape).astype(np.float32)) for W, b in layers]
    x = rng.normal(0, 0.5, size=(1000, in_dim)).astype(np.float32)
    dy = rng.normal(size=(1000, units[-1])).astype(np.float32)
    y_ref, acts = O.mlp_forward(x, layers, act)
    dx_ref, grads_ref = O.mlp_backward(dy, acts, layers, act)
    mlp = MLP(units, act, collapse_linear=True)
    mlp.load_arrays(layers, "cuda")
    xt = cu(x).requires_grad_(True)
    y = mlp(xt)
    (y * cu(dy)).sum().backward()

    def close(a, b):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b, rtol=0, atol=1e-4 * max(np.abs(b).max(), 1e-6))   # TF32-free fp32 GEMMs
    close(y, y_ref)
    close(xt.grad, dx_ref)
    for W, b, (dW, db) in zip(mlp.kernels, mlp.biases, grads_ref):
        close(W.grad, dW)
        close(b.grad, db)


@pytest.mark.parametrize("name", ["dlrm_small", "dlrm_uniform"])
def test_collapsed_dlrm_against_reference_fixture_and_layerwise_model(cuda_lib, golden, name):
    """Same bars as tests/test_gpu_models.py::test_dlrm_against_reference_fixture for the MLP gradients, and the table
    after one Adam step equals the layer-by-layer model's within the bf16 bound."""
    from recommender_b200.model import DLRM, bce_clipped
    from recommender_b200.optimizers import Adam
    g = golden(name)
    bottom, top = _mlp(g, "bottom"), _mlp(g, "top")
    D, V = g["table"].shape[1], g["table"].shape[0]
    tables = {}
    for collapse in (False, True):
        model = DLRM([W.shape[1] for W, _ in bottom], [W.shape[1] for W, _ in top], D, V, 26, 13, device="cuda",
                     collapse_linear=collapse)
        model.embedding_layer.embeddings.copy_(cu(g["table"]))
        model.bottom_mlp.load_arrays(bottom, "cuda")
        model.top_mlp.load_arrays(top, "cuda")
        prob = model({"cat_features": cu(g["cat"]), "int_features": cu(g["dense"])})
        np.testing.assert_allclose(prob.detach().cpu().numpy(), g["prob"], rtol=0, atol=2e-3)
        loss = bce_clipped(prob, cu(g["label"]))
        np.testing.assert_allclose(loss.item(), g["loss"], rtol=2e-3)
        loss.backward()
        for tower, mlp in (("top", model.top_mlp), ("bottom", model.bottom_mlp)):
            for i, (W, b) in enumerate(zip(mlp.kernels, mlp.biases)):
                ref = g[f"{tower}_dW{i}"]
                assert np.abs(W.grad.cpu().numpy() - ref).max() <= 4 * BF16_REL * np.abs(ref).max() + 1e-7
                assert torch.isfinite(b.grad).all()
        Adam().apply_gradients(model)
        tables[collapse] = model.embedding_layer.embeddings.cpu().numpy()
    moved = np.abs(tables[False] - g["table"]).max()
    assert moved > 0 and np.abs(tables[True] - tables[False]).max() <= 0.05 * moved + 1e-7


def test_collapsed_dlrm_trains_as_a_cuda_graph(cuda_lib):
    from recommender_b200.graph import GraphedTrainStep
    from recommender_b200.model import DLRM, bce_clipped
    from recommender_b200.optimizers import Adam
    gen = torch.Generator(device="cuda").manual_seed(4)
    model = DLRM([64, 32, 16], [64, 32, 1], 16, 5000, 26, 13, device="cuda", compute_dtype=torch.bfloat16, generator=gen,
                 collapse_linear=True)
    B = 512
    cat = torch.randint(0, 5000, (B, 26), device="cuda", generator=gen)
    dense = torch.rand(B, 13, device="cuda", generator=gen)
    label = (torch.rand(B, device="cuda", generator=gen) < 0.25).float()
    step = GraphedTrainStep(model, Adam(), bce_clipped, (cat, dense, label))
    first = float(step.loss)
    for _ in range(20):
        loss = step.step((cat, dense, label))
    assert np.isfinite(first) and float(loss) < first                           # it fits the batch it keeps seeing
