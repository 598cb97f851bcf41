"""Freeze golden vectors from the reference's OWN Python files.

Run in the build container only (needs /root/reference, which is absent on the GPU box):

    python tests/golden/make_golden.py

It imports /root/reference/ctr/{model,layers}.py, dien/{layers,model}.py and esmm/esmm.py
byte-for-byte under the `tensorflow` shim in oracle/tf_shim (torch-CPU tensors), feeds them
seeded synthetic Criteo-shaped inputs with weights owned by the oracle's initialisers, and
stores inputs, weights, forward outputs and torch-autograd gradients of that exact graph as
small .npz fixtures next to this script.  The committed fixtures pin (a) the numpy oracle and
(b) the CUDA path on the GPU box.  TF-internal semantics (dedup order, Adam) are NOT pinned by
this — they are restated in oracle/ctr_oracle.py.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle", "tf_shim"))

from oracle import ctr_oracle as O  # noqa: E402


def import_ref(package, *modules):
    """Import reference modules of one flat package (they use sibling imports)."""
    for name in ("layers", "model", "esmm", "mmoe", "base"):
        sys.modules.pop(name, None)
    sys.path.insert(0, os.path.join(REF, package))
    try:
        return [importlib.import_module(m) for m in modules]
    finally:
        sys.path.pop(0)


def set_dense(mlp_layers, arrays):
    for layer, (W, b) in zip(mlp_layers, arrays):
        layer.kernel = torch.tensor(W, requires_grad=True)
        layer.bias = torch.tensor(b, requires_grad=True)


def np_(t):
    return t.detach().numpy().copy()


def bce_clipped_t(p, y):           # SURVEY A.5, probability form (DLRM)
    eps = 1e-7
    p = p.clamp(eps, 1 - eps)
    return (-(y * torch.log(p + eps) + (1 - y) * torch.log(1 - p + eps))).mean()


def bce_logits_t(x, y):            # SURVEY A.5, logits form (DeepFM)
    return (x.clamp(min=0) - x * y + torch.log1p(torch.exp(-x.abs()))).mean()


def golden_dlrm(B=64, D=16, V=1000, bottom=(32, 16), top=(32, 1), seed=4, dist="zipf"):
    (model_mod,) = import_ref("ctr", "model")
    params = O.init_dlrm(seed, list(bottom), list(top), D, V)
    cat, dense_x, label = O.synth_batch(B, V, seed=seed, dist=dist)
    model = model_mod.DLRM(list(bottom), list(top), D, V, 26, 13)
    model.embedding_layer.embeddings = torch.tensor(params["table"], requires_grad=True)
    x = {"cat_features": torch.tensor(cat), "int_features": torch.tensor(dense_x)}
    # Dense layers build lazily: run once, then overwrite with the oracle-owned weights
    model(x)
    set_dense(model.bottom_mlp.mlp, params["bottom"])
    set_dense(model.top_mlp.mlp, params["top"])

    # intermediates, by replaying ctr/model.py:49-55 with hooks on the reference's own layers
    captured = {}
    orig_inter = model.interaction.call

    def spy(inputs):
        inputs.retain_grad()
        captured["X"] = inputs
        out = orig_inter(inputs)
        out.retain_grad()
        captured["inter"] = out
        return out

    model.interaction.call = spy
    prob = model(x)
    loss = bce_clipped_t(prob, torch.tensor(label, dtype=torch.float32))
    loss.backward()
    X = captured["X"]
    out = dict(cat=cat, dense=dense_x, label=label, table=params["table"], prob=np_(prob), loss=np_(loss),
               X=np_(X), dX=np_(X.grad), inter=np_(captured["inter"]), dinter=np_(captured["inter"].grad),
               dtable=np_(model.embedding_layer.embeddings.grad))
    for name, layers, arrs in (("bottom", model.bottom_mlp.mlp, params["bottom"]), ("top", model.top_mlp.mlp, params["top"])):
        for i, (layer, (W, b)) in enumerate(zip(layers, arrs)):
            out[f"{name}_W{i}"], out[f"{name}_b{i}"] = W, b
            out[f"{name}_dW{i}"], out[f"{name}_db{i}"] = np_(layer.kernel.grad), np_(layer.bias.grad)
    return out


def golden_deepfm(B=64, D=16, V=1000, units=(32, 16, 1), seed=4, dist="zipf"):
    (model_mod,) = import_ref("ctr", "model")
    params = O.init_deepfm(seed, D, V, 13, 26, list(units))
    cat, dense_x, label = O.synth_batch(B, V, seed=seed + 1, dist=dist)
    model = model_mod.DeepFM(D, V, 13, 26, list(units))
    table = torch.tensor(params["table"], requires_grad=True)
    model.embedding_layer.embeddings = table
    x = {"cat_features": torch.tensor(cat), "int_features": torch.tensor(dense_x)}
    model(x)
    set_dense(model.mlp.mlp, params["mlp"])
    # capture E and the pre-sigmoid logit by spying on the reference's own layer objects
    captured = {}
    orig_emb = model.embedding_layer.call

    def spy_emb(idx):
        e = orig_emb(idx)
        e.retain_grad()
        captured["E"] = e
        return e

    model.embedding_layer.call = spy_emb
    prob = model(x)
    logit = torch.log(prob / (1 - prob))        # only used for the logits-form loss value
    y = torch.tensor(label, dtype=torch.float32)
    loss = bce_logits_t(logit, y)
    # gradient: d/dlogit of the logits-form loss is (sigmoid(logit) - y)/B = (prob - y)/B
    prob.backward(gradient=((prob.detach() - y) / B) / (prob.detach() * (1 - prob.detach())))
    E = captured["E"]
    fm = 0.5 * ((E.sum(1)) ** 2 - (E ** 2).sum(1)).sum(1)
    out = dict(cat=cat, dense=dense_x, label=label, table=params["table"], prob=np_(prob), loss=np_(loss),
               E=np_(E), dE=np_(E.grad), fm=np_(fm), dtable=np_(table.grad))
    for i, (layer, (W, b)) in enumerate(zip(model.mlp.mlp, params["mlp"])):
        out[f"mlp_W{i}"], out[f"mlp_b{i}"] = W, b
        out[f"mlp_dW{i}"], out[f"mlp_db{i}"] = np_(layer.kernel.grad), np_(layer.bias.grad)
    return out


def golden_dot_interaction(B=8, Fp=27, D=16, seed=7):
    (layers_mod,) = import_ref("ctr", "layers")
    rng = np.random.default_rng(seed)
    X = rng.normal(0, 0.5, size=(B, Fp, D)).astype(np.float32)
    out = dict(X=X)
    for si in (False, True):
        for sg in (False, True):
            layer = layers_mod.DotInteraction(si, sg)
            xt = torch.tensor(X, requires_grad=True)
            y = layer(xt)
            dy = rng.normal(0, 1.0, size=tuple(y.shape)).astype(np.float32)
            y.backward(torch.tensor(dy))
            tag = f"si{int(si)}_sg{int(sg)}"
            out[f"out_{tag}"], out[f"dout_{tag}"], out[f"dX_{tag}"] = np_(y), dy, np_(xt.grad)
    return out


def golden_masked_mean(B=16, L=100, V_item=500, V_cat=40, D=18, seed=11):
    """dien BaseModel fragments: compute_flat_embedding + mask + compute_his_average
    (dien/model.py:14-19,25-31; dien/layers.py:5-17).  Reference D = 18 + 18 (dien/train.py:91-92)."""
    layers_mod, model_mod = import_ref("dien", "layers", "model")
    rng = np.random.default_rng(seed)
    W_item, W_cat = O.init_table(rng, V_item, D), O.init_table(rng, V_cat, D)
    lengths = rng.integers(1, L + 1, size=B)
    item = np.zeros((B, L), dtype=np.int32)
    cat = np.zeros((B, L), dtype=np.int32)
    for b, n in enumerate(lengths):     # post-padding with zeros (dien/data_loader.py:44,48)
        item[b, :n] = rng.integers(1, V_item, size=n)
        cat[b, :n] = rng.integers(1, V_cat, size=n)
    model = model_mod.BaseModel(V_item, D, V_cat, D, [8, 1])
    model.item_embedding.embeddings = torch.tensor(W_item, requires_grad=True)
    model.cat_embedding.embeddings = torch.tensor(W_cat, requires_grad=True)
    mask = model.item_embedding.compute_mask(torch.tensor(item))                       # dien/model.py:25
    his = model.compute_flat_embedding((torch.tensor(item), torch.tensor(cat)))        # dien/model.py:29-30
    his.retain_grad()
    # `his_embedding *= mask` (dien/layers.py:13) rebinds a name in TF but mutates a torch tensor in
    # place: hand the reference a clone so his.grad is the gradient w.r.t. the un-masked embedding.
    avg = layers_mod.compute_his_average(his.clone(), mask)                            # dien/model.py:31
    davg = rng.normal(0, 1e-2, size=tuple(avg.shape)).astype(np.float32)
    avg.backward(torch.tensor(davg))
    return dict(item=item, cat=cat, W_item=W_item, W_cat=W_cat, avg=np_(avg), davg=davg, dhis=np_(his.grad),
                dW_item=np_(model.item_embedding.embeddings.grad), dW_cat=np_(model.cat_embedding.embeddings.grad))


def golden_din_attention(B=16, L=20, V_item=500, V_cat=40, D=18, seed=17):
    """DIN's attention pooling from the reference's own classes: compute_flat_embedding (dien/model.py:14-19), the
    mask_zero mask (dien/model.py:43) and LocalActivationUnit (dien/layers.py:34-59) as DIN.call wires them
    (dien/model.py:42-50).  Reference D = 18 + 18 (dien/train.py:91-92), so E = 36 and the attention MLP reads 144 columns."""
    layers_mod, model_mod = import_ref("dien", "layers", "model")
    rng = np.random.default_rng(seed)
    W_item, W_cat = O.init_table(rng, V_item, D), O.init_table(rng, V_cat, D)
    lengths = rng.integers(1, L + 1, size=B)
    lengths[0], lengths[1] = L, 1
    item = np.zeros((B, L), dtype=np.int32)
    cat = np.zeros((B, L), dtype=np.int32)
    for b, n in enumerate(lengths):     # post-padding with zeros (dien/data_loader.py:44,48)
        item[b, :n] = rng.integers(1, V_item, size=n)
        cat[b, :n] = rng.integers(1, V_cat, size=n)
    item[2, 0] = 0                      # a hole inside a history: the mask is item != 0 wherever it occurs
    t_item = rng.integers(1, V_item, size=(B, 1)).astype(np.int32)
    t_cat = rng.integers(1, V_cat, size=(B, 1)).astype(np.int32)
    t_item[3, 0] = item[3, 0]           # a target that also occurs in its own history: two uses of one table row
    model = model_mod.DIN(item_vocab_size=V_item, item_embedding_size=D, cat_vocab_size=V_cat, cat_embedding_size=D, mlp_units=[8, 1])
    model.item_embedding.embeddings = torch.tensor(W_item, requires_grad=True)
    model.cat_embedding.embeddings = torch.tensor(W_cat, requires_grad=True)
    unit = model.local_activation_unit
    mask = model.item_embedding.compute_mask(torch.tensor(item))                                    # dien/model.py:43
    target = model.compute_flat_embedding((torch.tensor(t_item), torch.tensor(t_cat)))              # :44-45  [B, 1, E]
    his = model.compute_flat_embedding((torch.tensor(item), torch.tensor(cat)))                     # :46-47  [B, L, E]
    unit((target, his), mask=mask)                                                                  # builds the Dense layers
    att = O.init_mlp(rng, 4 * 2 * D, [80, 40, 1])
    att = [(W, rng.normal(0, 0.05, size=b.shape).astype(np.float32)) for W, b in att]               # non-zero biases
    set_dense([unit.layer_1, unit.layer_2, unit.layer_3], att)
    target.retain_grad()
    his.retain_grad()
    rep = unit((target, his), mask=mask)                                                            # :48
    d_rep = rng.normal(0, 1e-1, size=tuple(rep.shape)).astype(np.float32)
    rep.backward(torch.tensor(d_rep))
    out = dict(item=item, cat=cat, t_item=t_item, t_cat=t_cat, W_item=W_item, W_cat=W_cat, rep=np_(rep), d_rep=d_rep,
               d_target=np_(target.grad)[:, 0, :], d_his=np_(his.grad), dW_item=np_(model.item_embedding.embeddings.grad),
               dW_cat=np_(model.cat_embedding.embeddings.grad))
    for i, (layer, (W, b)) in enumerate(zip([unit.layer_1, unit.layer_2, unit.layer_3], att)):
        out[f"att_W{i}"], out[f"att_b{i}"] = W, b
        out[f"att_dW{i}"], out[f"att_db{i}"] = np_(layer.kernel.grad), np_(layer.bias.grad)
    return out


def golden_esmm(B=32, D=18, seed=13):
    """esmm/esmm.py:15-27 with two consumers (ctr and cvr towers) of one concat embedding."""
    (esmm_mod,) = import_ref("esmm", "esmm")
    rng = np.random.default_rng(seed)
    feat_vocab = {"101": 300, "121": 98, "122": 14, "124": 3, "125": 8, "126": 4}     # subset of esmm/train.py:197-215
    tables = {f: O.init_table(rng, v, D) for f, v in feat_vocab.items()}
    inputs = {f: rng.integers(0, v, size=(B, 1)).astype(np.int32) for f, v in feat_vocab.items()}
    model = esmm_mod.ESMM([16, 1], feat_vocab, D)
    for f in feat_vocab:
        model.embedding_layer[f].embeddings = torch.tensor(tables[f], requires_grad=True)
    tin = {f: torch.tensor(a) for f, a in inputs.items()}
    emb = model.compute_embedding(tin)
    out = model(tin)                                   # builds + runs both towers on the shared embedding
    w = torch.tensor(rng.normal(0, 1, size=tuple(out.shape)).astype(np.float32))
    (out * w).sum().backward()
    res = dict(emb=np_(emb), feats=np.array(list(feat_vocab)), vocab=np.array(list(feat_vocab.values())))
    for f in feat_vocab:
        res[f"W_{f}"], res[f"idx_{f}"] = tables[f], inputs[f]
        res[f"dW_{f}"] = np_(model.embedding_layer[f].embeddings.grad)
    return res


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    jobs = dict(dlrm_small=golden_dlrm, deepfm_small=golden_deepfm, dot_interaction=golden_dot_interaction,
                masked_mean=golden_masked_mean, esmm_small=golden_esmm, din_attention=golden_din_attention,
                dlrm_uniform=lambda: golden_dlrm(B=32, D=64, V=4096, bottom=(32, 64), top=(16, 1), seed=5, dist="uniform"))
    only = sys.argv[1:]          # python make_golden.py [name ...]: regenerate just these fixtures
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        arrays = fn()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, keys={len(arrays)}")


if __name__ == "__main__":
    main()
