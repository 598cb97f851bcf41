"""numpy restatement of the CTR hot path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Every function cites the reference line it follows (paths relative to
/root/reference) or, for TensorFlow-internal behaviour, the SURVEY.md appendix
item that restates the published Keras semantics.  All arithmetic is fp32 unless
a name ends in ``_f64``.  No FMA contraction happens in numpy, so the CUDA row
update (which uses explicit round-to-nearest mul/add) can match bit for bit.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------
# ids -> rows (the reference has no hashing; SURVEY §8c "index hashing": the build defines it)
# --------------------------------------------------------------------------------------


def id_to_row(ids: np.ndarray, vocab_size: int) -> np.ndarray:
    """row = uint64(id) mod V.  Ids already in [0, V) (ctr/train.py:64) map to themselves."""
    return (ids.astype(np.int64).view(np.uint64) % np.uint64(vocab_size)).astype(np.int64)


def shard_of_row(row: np.ndarray, world: int):
    """Row-wise sharding: owner = row mod G, local = row div G."""
    row = np.asarray(row, dtype=np.int64)
    return row % world, row // world


def table_owner(num_tables: int, world: int) -> np.ndarray:
    """Table-wise sharding: table t lives on rank t mod G (equal-size tables)."""
    return np.arange(num_tables, dtype=np.int64) % world


# --------------------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------------------


def embedding_lookup(W: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """E[b,f,:] = W[idx[b,f],:]  (ctr/model.py:19, :49; keras Embedding -> ResourceGather, A.6).

    Out-of-range ids raise, like TF's CPU kernel (InvalidArgument)."""
    idx = np.asarray(idx)
    if idx.size and (idx.min() < 0 or idx.max() >= W.shape[0]):
        raise IndexError("embedding index out of range [0, %d)" % W.shape[0])
    return W[idx]


def fm_second_order(E: np.ndarray) -> np.ndarray:
    """ctr/model.py:21-23: 0.5 * sum_d[(sum_f e)^2 - sum_f e^2] -> f32[B]."""
    sum_square = np.square(E.sum(axis=1, dtype=F32))
    square_sum = np.square(E).sum(axis=1, dtype=F32)
    return (F32(0.5) * (sum_square - square_sum).sum(axis=1, dtype=F32)).astype(F32)


def fm_second_order_f64(E: np.ndarray) -> np.ndarray:
    E = E.astype(np.float64)
    return 0.5 * (np.square(E.sum(1)) - np.square(E).sum(1)).sum(1)


def keep_mask(num_feat: int, self_interaction: bool) -> np.ndarray:
    """The kept (i,j) set of DotInteraction (ctr/layers.py:27-33).

    self_interaction=False: ones - band_part(ones,-1,0) = strict upper triangle (j > i).
    self_interaction=True : the variable named upper_matrix is band_part(ones,-1,0)
    = lower triangle including the diagonal (j <= i)  (SURVEY Appendix B)."""
    lower_incl = np.tril(np.ones((num_feat, num_feat), dtype=bool))
    return lower_incl if self_interaction else ~lower_incl


def dot_interaction(X: np.ndarray, self_interaction: bool = False, skip_gather: bool = True,
                    operand_dtype=None) -> np.ndarray:
    """ctr/layers.py:23-43.  X f32[B,F',D] -> f32[B,F'^2] (skip_gather) or compact [B,count].

    operand_dtype='bf16' rounds the operands to bfloat16 first (fp32 accumulate), the
    arithmetic the sm_100a tensor-core kernel performs (north star)."""
    B, Fp, _ = X.shape
    Xo = round_bf16(X) if operand_dtype == "bf16" else X
    Z = np.matmul(Xo, Xo.transpose(0, 2, 1)).astype(F32)         # :25
    keep = keep_mask(Fp, self_interaction)                        # :27-33
    if skip_gather:                                               # :35-38
        return np.where(keep[None], Z, F32(0)).reshape(B, Fp * Fp).astype(F32)
    return Z[:, keep].reshape(B, int(keep.sum())).astype(F32)     # :39-42 (row-major kept order)


def dot_interaction_backward(X, dOut, self_interaction=False, skip_gather=True, operand_dtype=None):
    """TF autodiff of ctr/layers.py:25-42: G = mask (.) dOut, dX = (G + G^T) X  (SURVEY a11)."""
    B, Fp, _ = X.shape
    keep = keep_mask(Fp, self_interaction)
    if skip_gather:
        G = np.where(keep[None], dOut.reshape(B, Fp, Fp), F32(0)).astype(F32)
    else:
        G = np.zeros((B, Fp, Fp), dtype=F32)
        G[:, keep] = dOut
    S = G + G.transpose(0, 2, 1)
    Xo = X
    if operand_dtype == "bf16":
        S, Xo = round_bf16(S), round_bf16(X)
    return np.matmul(S, Xo).astype(F32)


def round_bf16(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32, bit-exact with cvt.rn.bf16.f32."""
    u = np.ascontiguousarray(x, dtype=F32).view(np.uint32)
    rounded = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))
    out = (rounded & np.uint32(0xFFFF0000)).view(F32)
    nan = np.isnan(x)
    if nan.any():
        out = np.where(nan, x, out)
    return out.reshape(x.shape)


def fm_backward(E: np.ndarray, g: np.ndarray) -> np.ndarray:
    """TF autodiff of ctr/model.py:21-23: dE[b,f,:] = g[b] * (s[b,:] - E[b,f,:])."""
    s = E.sum(axis=1, dtype=F32)
    return (g[:, None, None] * (s[:, None, :] - E)).astype(F32)


def dense(x, W, b, activation=None):
    """keras.layers.Dense: x.W + b, optional activation (ctr/layers.py:8-9)."""
    y = (x @ W + b).astype(F32)
    if activation == "relu":
        y = np.maximum(y, F32(0))
    elif activation == "sigmoid":
        y = sigmoid(y)
    elif activation is not None:
        raise ValueError(activation)
    return y


def sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x.astype(F32)))).astype(F32)


def mlp_forward(x, layers, final_activation, operand_dtype=None):
    """ctr/layers.py:5-14: hidden Dense layers are linear, only the last has an activation.

    operand_dtype='bf16' restates the rounding points of the CUDA towers (csrc/mlp.cu): every GEMM reads its input
    activations and its kernel rounded to bfloat16, accumulates in fp32, adds the fp32 bias; hidden activations are
    stored in bf16; the last layer's activation acts on the fp32 accumulator.  acts[i] is what layer i's GEMMs read."""
    if operand_dtype != "bf16":
        acts = [x]
        for i, (W, b) in enumerate(layers):
            x = dense(x, W, b, final_activation if i == len(layers) - 1 else None)
            acts.append(x)
        return x, acts
    x = round_bf16(x)
    acts = [x]
    for i, (W, b) in enumerate(layers):
        last = i == len(layers) - 1
        x = dense(x, round_bf16(W), b, final_activation if last else None)
        if not last:
            x = round_bf16(x)
        acts.append(x)
    return x, acts


def mlp_backward(dy, acts, layers, final_activation, operand_dtype=None):
    """Returns dx and [(dW, db)] for mlp_forward.  operand_dtype='bf16': activation gradients travel between layers in
    bf16 (a Dense(1) last layer keeps its fp32 dz for dW / db, as rb_dense_head_bwd does)."""
    y = acts[-1]
    if final_activation == "relu":
        dy = dy * (y > 0)
    elif final_activation == "sigmoid":
        dy = dy * y * (F32(1) - y)
    bf16 = operand_dtype == "bf16"
    grads = []
    for i in range(len(layers) - 1, -1, -1):
        W, _ = layers[i]
        x = acts[i]
        if bf16:
            W = round_bf16(W)
            if not (i == len(layers) - 1 and W.shape[1] == 1):
                dy = round_bf16(dy)
        grads.append(((x.T @ dy).astype(F32), dy.sum(0, dtype=F32)))
        dy = (dy @ W.T).astype(F32)
    if bf16:
        dy = round_bf16(dy)
    return dy, grads[::-1]


# --------------------------------------------------------------------------------------
# models (ctr/model.py)
# --------------------------------------------------------------------------------------


def deepfm_forward(params, cat_features, int_features, num_int_fea=13, num_cat_fea=26, mlp_dtype=None):
    """ctr/model.py:15-31.  params: {'table', 'mlp': [(W,b)...]}.  Returns (prob[B], cache).  mlp_dtype='bf16': the MLP's
    GEMMs read bf16 operands (mlp_forward)."""
    int_features = np.reshape(int_features, (-1, num_int_fea)).astype(F32)           # :17
    cat_features = np.reshape(cat_features, (-1, num_cat_fea))                      # :18
    E = embedding_lookup(params["table"], cat_features)                             # :19
    interaction = fm_second_order(E)                                                # :21-23
    D = E.shape[2]
    deep_cat_input = E.reshape(-1, num_cat_fea * D)                                 # :25
    deep_input = np.concatenate([deep_cat_input, int_features], axis=1)             # :26
    dense_output, acts = mlp_forward(deep_input, params["mlp"], None, mlp_dtype)    # :27
    logit = interaction + dense_output[:, 0]                                        # :28-29
    prob = sigmoid(logit)                                                           # :30
    return prob, dict(E=E, acts=acts, logit=logit, idx=cat_features, fm=interaction)


def deepfm_backward(params, cache, dlogit, num_cat_fea=26, mlp_dtype=None):
    """Gradients for deepfm_forward given dL/dlogit.  Table grad = sum of the three consumers
    of cat_embedding (ctr/model.py:21, :22, :25)."""
    E = cache["E"]
    B, Fc, D = E.shape
    dx, mlp_grads = mlp_backward(dlogit[:, None].astype(F32), cache["acts"], params["mlp"], None, mlp_dtype)
    dE = dx[:, : Fc * D].reshape(B, Fc, D) + fm_backward(E, dlogit.astype(F32))
    return dict(dE=dE.astype(F32), mlp=mlp_grads)


def dlrm_forward(params, cat_features, int_features, num_int_fea=13, num_cat_fea=26,
                 operand_dtype=None, mlp_dtype=None):
    """ctr/model.py:45-58.  params: {'table', 'bottom': [...], 'top': [...]}.  mlp_dtype='bf16': the towers' GEMMs read
    bf16 operands (mlp_forward), the configuration bench.py times."""
    int_features = np.reshape(int_features, (-1, num_int_fea)).astype(F32)           # :47
    cat_features = np.reshape(cat_features, (-1, num_cat_fea))                      # :48
    E = embedding_lookup(params["table"], cat_features)                             # :49
    bmlp, bacts = mlp_forward(int_features, params["bottom"], "relu", mlp_dtype)    # :50
    X = np.concatenate([E, bmlp[:, None, :]], axis=1)                               # :51-52
    inter = dot_interaction(X, False, True, operand_dtype)                          # :53
    tmlp_input = np.concatenate([inter, bmlp], axis=1)                              # :54
    D = E.shape[2]
    tmlp_input = tmlp_input.reshape(-1, (num_cat_fea + 1) ** 2 + D)                 # :55
    out, tacts = mlp_forward(tmlp_input, params["top"], "sigmoid", mlp_dtype)       # :56
    prob = out[:, 0]                                                                # :57
    return prob, dict(E=E, X=X, bacts=bacts, tacts=tacts, idx=cat_features, inter=inter)


def dlrm_backward(params, cache, dprob, num_cat_fea=26, operand_dtype=None, mlp_dtype=None):
    X = cache["X"]
    Fp = num_cat_fea + 1
    dtin, top_grads = mlp_backward(dprob[:, None].astype(F32), cache["tacts"], params["top"], "sigmoid", mlp_dtype)
    dinter, dbmlp_direct = dtin[:, : Fp * Fp], dtin[:, Fp * Fp:]
    dX = dot_interaction_backward(X, dinter, False, True, operand_dtype)
    dE = dX[:, :num_cat_fea]
    dbmlp = dX[:, num_cat_fea] + dbmlp_direct
    _, bottom_grads = mlp_backward(dbmlp.astype(F32), cache["bacts"], params["bottom"], "relu", mlp_dtype)
    return dict(dE=np.ascontiguousarray(dE, dtype=F32), top=top_grads, bottom=bottom_grads)


# --------------------------------------------------------------------------------------
# loss (ctr/train.py:85-87)
# --------------------------------------------------------------------------------------


def bce_clipped(prob, label):
    """Keras binary_crossentropy on probabilities (DLRM: last op is Squeeze).  Returns
    (mean loss, dL/dprob).  The clip has zero gradient outside [eps, 1-eps]."""
    eps = F32(1e-7)
    y = label.astype(F32)
    p = np.clip(prob.astype(F32), eps, F32(1) - eps)
    loss = -(y * np.log(p + eps) + (F32(1) - y) * np.log(F32(1) - p + eps))
    inside = (prob >= eps) & (prob <= F32(1) - eps)
    dp = (-(y / (p + eps)) + (F32(1) - y) / (F32(1) - p + eps)) * inside
    n = F32(prob.shape[0])
    return F32(loss.mean(dtype=F32)), (dp / n).astype(F32)


def bce_logits(logit, label):
    """sigmoid_cross_entropy_with_logits, the form Keras recovers for DeepFM in graph mode
    (its last op is Sigmoid, ctr/model.py:30).  Returns (mean loss, dL/dlogit)."""
    y = label.astype(F32)
    x = logit.astype(F32)
    loss = np.maximum(x, F32(0)) - x * y + np.log1p(np.exp(-np.abs(x)))
    n = F32(x.shape[0])
    return F32(loss.mean(dtype=F32)), ((sigmoid(x) - y) / n).astype(F32)


# --------------------------------------------------------------------------------------
# backward of the gather + dedup + sparse optimizers
# --------------------------------------------------------------------------------------


def gather_backward(idx: np.ndarray, dE: np.ndarray):
    """A.1: IndexedSlices(values = dE.reshape(N,D), indices = idx.reshape(N)); nothing is summed."""
    D = dE.shape[-1]
    return idx.reshape(-1).astype(np.int64), np.ascontiguousarray(dE, dtype=F32).reshape(-1, D)


def concat_indexed_slices(slices):
    """A.1: several uses of one table concatenate their slices in use order."""
    return (np.concatenate([s[0] for s in slices]), np.concatenate([s[1] for s in slices]))


def dedup_indexed_slices(indices: np.ndarray, values: np.ndarray):
    """A.2 (_deduplicate_indexed_slices): unique in FIRST-OCCURRENCE order, then
    unsorted_segment_sum adding rows in INPUT order (the CPU kernel's order), fp32."""
    uniq_sorted, first_pos, inverse = np.unique(indices, return_index=True, return_inverse=True)
    order = np.argsort(first_pos, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    unique_indices = uniq_sorted[order]
    positions = rank[inverse.reshape(-1)]
    summed = np.zeros((unique_indices.size, values.shape[1]), dtype=F32)
    np.add.at(summed, positions, values.astype(F32))   # unbuffered, in input order
    return unique_indices.astype(np.int64), summed


def dedup_indexed_slices_f64(indices, values):
    uniq, inverse = np.unique(indices, return_inverse=True)
    summed = np.zeros((uniq.size, values.shape[1]), dtype=np.float64)
    np.add.at(summed, inverse.reshape(-1), values.astype(np.float64))
    return uniq.astype(np.int64), summed


ADAM_DEFAULTS = dict(lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7)       # ctr/train.py:80,84
ADAGRAD_DEFAULTS = dict(lr=1e-3, initial_accumulator_value=0.1, epsilon=1e-7)


def _libm_powf():
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.powf.restype = ctypes.c_float
    libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]
    return libm.powf


_powf = _libm_powf()


def adam_alpha_t(step: int, lr=1e-3, beta_1=0.9, beta_2=0.999) -> np.float32:
    """A.3: alpha_t = lr * sqrt(1 - b2^t) / (1 - b1^t), evaluated in fp32 like Keras does
    (math_ops.pow on float32 scalars = std::pow<float> = libm powf on TF's CPU device; numpy's
    SIMD float32 power differs from powf in the last bit for some t, so libm is called directly)."""
    t = float(step)
    b1p = F32(_powf(float(F32(beta_1)), t))
    b2p = F32(_powf(float(F32(beta_2)), t))
    return F32(F32(lr) * np.sqrt(F32(1) - b2p, dtype=F32) / (F32(1) - b1p))


def adam_tf_dense(var, m, v, rows, g, step, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
    """A.3 Keras Adam._resource_apply_sparse: m, v decay and var moves on EVERY row each step;
    touched rows additionally receive the (deduplicated) gradient.  In place."""
    b1, b2, eps = F32(beta_1), F32(beta_2), F32(epsilon)
    omb1, omb2 = F32(1) - b1, F32(1) - b2
    alpha = adam_alpha_t(step, lr, beta_1, beta_2)
    m *= b1
    m[rows] += g * omb1
    v *= b2
    v[rows] += (g * g) * omb2
    var -= (alpha * m) / (np.sqrt(v) + eps)
    return var, m, v


def adam_lazy(var, m, v, rows, g, step, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
    """Row-sparse Adam (LazyAdam): the same formula on touched rows only.  Equal to
    adam_tf_dense on step 1 from zero state (SURVEY §7 hard parts).  In place."""
    b1, b2, eps = F32(beta_1), F32(beta_2), F32(epsilon)
    omb1, omb2 = F32(1) - b1, F32(1) - b2
    alpha = adam_alpha_t(step, lr, beta_1, beta_2)
    g = g.astype(F32)
    m_new = m[rows] * b1 + g * omb1
    v_new = v[rows] * b2 + (g * g) * omb2
    m[rows] = m_new
    v[rows] = v_new
    var[rows] = var[rows] - (alpha * m_new) / (np.sqrt(v_new) + eps)
    return var, m, v


def adagrad(var, acc, rows, g, lr=1e-3, epsilon=1e-7):
    """A.4 Keras Adagrad sparse apply: acc[r] += g^2; var[r] -= lr*g/(sqrt(acc[r]) + eps)."""
    g = g.astype(F32)
    acc_new = acc[rows] + g * g
    acc[rows] = acc_new
    var[rows] = var[rows] - (F32(lr) * g) / (np.sqrt(acc_new) + F32(epsilon))
    return var, acc


def sgd(var, rows, g, lr=1e-2):
    """keras SGD sparse apply: var[r] -= lr * g (the commented-out option at ctr/train.py:79)."""
    var[rows] = var[rows] - F32(lr) * g.astype(F32)
    return var


def adam_dense_param(p, m, v, g, step, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
    """A.3 _resource_apply_dense (ResourceApplyAdam): same formula on a dense variable."""
    b1, b2 = F32(beta_1), F32(beta_2)
    alpha = adam_alpha_t(step, lr, beta_1, beta_2)
    m[...] = m * b1 + g * (F32(1) - b1)
    v[...] = v * b2 + (g * g) * (F32(1) - b2)
    p -= (alpha * m) / (np.sqrt(v) + F32(epsilon))


def sparse_backward_update(table, state, idx, dE, optimizer="adam_lazy", step=1, **hp):
    """Whole K7-K9 chain: IndexedSlices -> dedup -> row update.  Mutates table/state.
    Returns (unique rows in first-occurrence order, summed gradients)."""
    indices, values = gather_backward(idx, dE)
    rows, g = dedup_indexed_slices(indices, values)
    if optimizer == "adam_lazy":
        adam_lazy(table, state["m"], state["v"], rows, g, step, **hp)
    elif optimizer == "adam_tf_dense":
        adam_tf_dense(table, state["m"], state["v"], rows, g, step, **hp)
    elif optimizer == "adagrad":
        adagrad(table, state["acc"], rows, g, **hp)
    elif optimizer == "sgd":
        sgd(table, rows, g, **hp)
    else:
        raise ValueError(optimizer)
    return rows, g


# --------------------------------------------------------------------------------------
# config 4: masked mean over a behaviour history (dien/)
# --------------------------------------------------------------------------------------


def compute_flat_embedding(W_item, W_cat, item_idx, cat_idx):
    """dien/model.py:14-19: concat(item_emb, cat_emb) on the last axis."""
    return np.concatenate([embedding_lookup(W_item, item_idx), embedding_lookup(W_cat, cat_idx)], axis=-1)


def compute_his_average(his_embedding, mask):
    """dien/layers.py:5-17: sum_l(m*e) / sum_l(m); no guard for an all-pad history (NaN/Inf)."""
    m = mask[..., None].astype(his_embedding.dtype)                 # :11-12
    his = his_embedding * m                                         # :13
    mask_sum = m.sum(axis=1, dtype=F32)                             # :14
    embedding_sum = his.sum(axis=1, dtype=F32)                      # :15
    with np.errstate(divide="ignore", invalid="ignore"):
        return (embedding_sum / mask_sum).astype(F32)               # :16


def masked_mean_lookup(W_item, W_cat, item_idx, cat_idx):
    """dien/model.py:25-31: mask = item idx != 0 (compute_mask of mask_zero=True), applied to both tables."""
    mask = item_idx != 0                                            # dien/model.py:25
    return compute_his_average(compute_flat_embedding(W_item, W_cat, item_idx, cat_idx), mask)


def masked_mean_backward(dout, mask):
    """autodiff of compute_his_average: every position gets mask * dout / count (masked
    positions receive exact zeros but still appear in the IndexedSlices)."""
    m = mask.astype(F32)
    cnt = m.sum(axis=1, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        per_bag = dout / cnt[:, None]
    return (m[:, :, None] * per_bag[:, None, :]).astype(F32)


# --------------------------------------------------------------------------------------
# DIN attention pooling (dien/layers.py:34-59 at dien/model.py:42-53)
# --------------------------------------------------------------------------------------

def local_activation_unit(target, history, mask, layers, operand_dtype=None):
    """dien/layers.py:42-59.  target f32[B, E] (the reference's [B, 1, E] squeezed), history f32[B, L, E], mask bool[B, L],
    layers = [(W1[4E,80], b1), (W2[80,40], b2), (W3[40,1], b3)] (:37-39: sigmoid, sigmoid, none).
    Returns (history_representation f32[B, E], cache).

    operand_dtype='bf16' restates the rounding points of the CUDA path (csrc/din.cu + csrc/mlp.cu): the feature row
    [t, h, t - h, t * h] is formed in fp32 and stored in bf16, every GEMM reads bf16 operands and accumulates in fp32, hidden
    activations are stored in bf16; the Dense(1) logit, the mask and the weighted sum over fp32 history rows stay fp32."""
    bf16 = operand_dtype == "bf16"
    rnd = round_bf16 if bf16 else (lambda a: a)
    B, L, E = history.shape
    t = np.broadcast_to(target[:, None, :], (B, L, E))                                  # :47 tf.repeat
    x = rnd(np.concatenate([t, history, (t - history).astype(F32), (t * history).astype(F32)], axis=-1).astype(F32))   # :48
    (W1, b1), (W2, b2), (W3, b3) = layers
    a1 = rnd(dense(x, rnd(W1), b1, "sigmoid"))                                           # :49
    a2 = rnd(dense(a1, rnd(W2), b2, "sigmoid"))                                          # :50
    w = dense(a2, rnd(W3), b3, None)                                                     # :51  [B, L, 1]
    m = mask[..., None].astype(F32)                                                      # :52-53
    w = (w * m).astype(F32)                                                              # :54
    rep = np.zeros((B, E), dtype=F32)
    for l in range(L):                                                                   # :55 weights^T . history, summed in position order
        rep = (rep + (w[:, l, :] * history[:, l, :]).astype(F32)).astype(F32)
    return rep, dict(target=target, history=history, mask=m, x=x, a1=a1, a2=a2, w=w, layers=layers, bf16=bf16)


def local_activation_unit_backward(cache, d_rep):
    """Hand-derived backward of local_activation_unit (pinned against torch autograd of the reference's own class in
    tests/golden/din_attention.npz).  Returns d_target[B, E], d_history[B, L, E] (zero rows at masked positions — they still
    appear in the IndexedSlices, like compute_his_average's) and [(dW, db)] x 3."""
    rnd = round_bf16 if cache["bf16"] else (lambda a: a)
    t, h, m, x, a1, a2, w = (cache[k] for k in ("target", "history", "mask", "x", "a1", "a2", "w"))
    (W1, b1), (W2, b2), (W3, b3) = cache["layers"]
    B, L, E = h.shape
    dw = ((d_rep[:, None, :] * h).sum(axis=-1, keepdims=True, dtype=F32) * m).astype(F32)     # d/dw of w^T h, through `weights *= mask`
    dW3 = np.einsum("blk,blo->ko", a2, dw).astype(F32)
    db3 = dw.sum(axis=(0, 1), dtype=F32)
    da2 = rnd((dw * rnd(W3)[None, None, :, 0]).astype(F32))
    dz2 = rnd((da2 * a2 * (F32(1) - a2)).astype(F32))
    dW2 = np.einsum("blk,blo->ko", a1, dz2).astype(F32)
    db2 = dz2.sum(axis=(0, 1), dtype=F32)
    da1 = rnd((dz2 @ rnd(W2).T).astype(F32))
    dz1 = rnd((da1 * a1 * (F32(1) - a1)).astype(F32))
    dW1 = np.einsum("blk,blo->ko", x, dz1).astype(F32)
    db1 = dz1.sum(axis=(0, 1), dtype=F32)
    dx = rnd((dz1 @ rnd(W1).T).astype(F32))
    d0, d1, d2, d3 = dx[..., :E], dx[..., E:2 * E], dx[..., 2 * E:3 * E], dx[..., 3 * E:]
    tb = t[:, None, :]
    dh = ((((d1 - d2).astype(F32) + (d3 * tb).astype(F32)).astype(F32) + (w * d_rep[:, None, :]).astype(F32)).astype(F32) * m).astype(F32)
    dt_pos = (((d0 + d2).astype(F32) + (d3 * h).astype(F32)).astype(F32) * m).astype(F32)
    dt = np.zeros((B, E), dtype=F32)
    for l in range(L):
        dt = (dt + dt_pos[:, l, :]).astype(F32)
    return dt, dh, [(dW1, db1), (dW2, db2), (dW3, db3)]


def bag_pool(W, idx, mode="sum", mask=None):
    """Generic bag pooling used by the gather kernel's modes: sum / mean over L, or masked mean."""
    E = embedding_lookup(W, idx)
    if mode == "sum":
        return E.sum(axis=1, dtype=F32)
    if mode == "mean":
        return (E.sum(axis=1, dtype=F32) / F32(idx.shape[1])).astype(F32)
    if mode == "masked_mean":
        return compute_his_average(E, mask if mask is not None else idx != 0)
    raise ValueError(mode)


# --------------------------------------------------------------------------------------
# config 5: one table per feature, bag size 1, several consumers (esmm/)
# --------------------------------------------------------------------------------------


def compute_embedding_multi(tables, inputs):
    """esmm/esmm.py:15-19: [W_f[idx_f[b,0]] for f in inputs] concatenated on the last axis,
    then squeeze(axis=1) -> f32[B, sum_f D_f].  `inputs` is an ordered dict feat -> int[B,1]."""
    embs = [embedding_lookup(tables[feat], inputs[feat]) for feat in inputs]   # :16
    return np.concatenate(embs, axis=-1)[:, 0, :]                                # :17-18


def multi_consumer_grad(consumer_grads):
    """autodiff adds the consumers' gradients at the concat output (esmm/esmm.py:23-24,
    esmm/mmoe.py:28-29,38), left to right in fp32."""
    total = consumer_grads[0].astype(F32).copy()
    for g in consumer_grads[1:]:
        total += g
    return total


def split_multi_table_grad(dconcat, dims):
    """Split d(concat) back per table -> list of f32[B,1,D_f] (gradient of esmm/esmm.py:17-18)."""
    out, o = [], 0
    for d in dims:
        out.append(np.ascontiguousarray(dconcat[:, None, o:o + d]))
        o += d
    return out


# --------------------------------------------------------------------------------------
# synthetic Criteo-shaped inputs and initialisers (Keras defaults)
# --------------------------------------------------------------------------------------


def synth_batch(batch, vocab_size, num_cat=26, num_int=13, seed=4, dist="uniform"):
    """seed 4 is the reference's default --seed (ctr/train.py:18).  dist: 'uniform' on [0,V) or
    'zipf' (alpha 1.05 folded mod V, 2% forced id 0 = OOV, ctr/tfrecord_io.py:61-64)."""
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        cat = rng.integers(0, vocab_size, size=(batch, num_cat), dtype=np.int64)
    elif dist == "zipf":
        cat = (rng.zipf(1.05, size=(batch, num_cat)).astype(np.uint64) % np.uint64(vocab_size)).astype(np.int64)
        cat[rng.random((batch, num_cat)) < 0.02] = 0
    else:
        raise ValueError(dist)
    dense_x = np.log1p(rng.integers(0, 1000, size=(batch, num_int)).astype(F32)).astype(F32)  # tfrecord_io.py:48-53
    label = (rng.random(batch) < 0.25).astype(np.int64)
    return cat, dense_x, label


def glorot_uniform(rng, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(F32)


def init_mlp(rng, in_dim, units):
    layers = []
    for u in units:
        layers.append((glorot_uniform(rng, in_dim, u), np.zeros(u, dtype=F32)))
        in_dim = u
    return layers


def init_table(rng, vocab_size, dim):
    """keras Embedding default initialiser: U(-0.05, 0.05)."""
    return rng.uniform(-0.05, 0.05, size=(vocab_size, dim)).astype(F32)


def init_dlrm(seed, bottom_mlp_units, top_mlp_units, embedding_size, vocab_size, num_cat_fea=26, num_int_fea=13):
    rng = np.random.default_rng(seed)
    assert bottom_mlp_units[-1] == embedding_size       # ctr/model.py:52,55 hard requirement
    return dict(table=init_table(rng, vocab_size, embedding_size),
                bottom=init_mlp(rng, num_int_fea, bottom_mlp_units),
                top=init_mlp(rng, (num_cat_fea + 1) ** 2 + embedding_size, top_mlp_units))


def init_deepfm(seed, embedding_size, vocab_size, num_int_fea, num_cat_fea, mlp_units):
    rng = np.random.default_rng(seed)
    return dict(table=init_table(rng, vocab_size, embedding_size),
                mlp=init_mlp(rng, num_cat_fea * embedding_size + num_int_fea, mlp_units))
