"""CPU stand-in for recommender_b200.ops, backed by the numpy oracle (TEST INFRASTRUCTURE).

Lets the host-side logic of the multi-GPU path (sharding maps, exchange plans, all-to-all split
sizes, permutations, optimizer plumbing) run under `gloo` with world_size 2 in the GPU-less
container.  Same function names and argument meaning as ops.py; tensors are torch-CPU.  The
product never imports this module."""
import numpy as np
import torch

from oracle import ctr_oracle as O
from recommender_b200.ops import GradSource, LookupGroup, interaction_ncols  # noqa: F401  (plain descriptors)

_launches = [0]


def kernel_launches():
    return _launches[0]


def check_oob(device):
    return None


def adam_alpha_t(lr, beta_1, beta_2, step):
    return float(O.adam_alpha_t(step, lr, beta_1, beta_2))


def _rows(idx, L, field_row_offset, hash_mod):
    ids = idx.numpy().astype(np.int64)
    if hash_mod:
        ids = O.id_to_row(ids, hash_mod)
    if field_row_offset is not None:
        ids = (ids.reshape(-1, L) + field_row_offset.numpy()[None]).reshape(ids.shape)
    return ids


def hash_ids(ids, vocab, world=1):
    rows = O.id_to_row(ids.numpy(), vocab)
    owner, local = O.shard_of_row(rows, max(world, 1))
    return torch.tensor(rows), torch.tensor(owner.astype(np.int32)), torch.tensor(local)


def gather_fwd(table, idx, *, L=1, field_row_offset=None, hash_mod=0, out=None, out_stride=None):
    _launches[0] += 1
    rows = _rows(idx, L, field_row_offset, hash_mod)
    return torch.tensor(O.embedding_lookup(table.detach().numpy(), rows))


def _X(E, table, idx, field_row_offset, dense_vec):
    if E is None:
        F = idx.shape[1]
        E = gather_fwd(table, idx, L=F, field_row_offset=field_row_offset)
    X = E.numpy()
    if dense_vec is not None:
        X = np.concatenate([X, dense_vec.detach().numpy()[:, None, :]], axis=1)
    return X


def dot_interaction_fwd(*, E=None, table=None, idx=None, field_row_offset=None, dense_vec=None, self_interaction=False,
                        skip_gather=True, tail=False, out=None, out_stride=None, out_dtype=torch.float32, pad_to=1, ones_col=False,
                        row_cache=False, row_cache_hint=None):
    _launches[0] += 1
    X = _X(E, table, idx, field_row_offset, dense_vec)
    res = O.dot_interaction(X, self_interaction, skip_gather, operand_dtype="bf16")
    if tail:
        res = np.concatenate([res, dense_vec.detach().numpy()], axis=1)
    res = torch.tensor(res)
    if out_dtype == torch.bfloat16:
        width = res.shape[1]
        padded = torch.zeros(res.shape[0], (width + pad_to - 1) // pad_to * pad_to, dtype=torch.bfloat16)
        padded[:, :width] = res.to(torch.bfloat16)
        if ones_col and padded.shape[1] > width:
            padded[:, width] = 1.0
        return padded
    return res


def dot_interaction_bwd(dOut, *, E=None, table=None, idx=None, field_row_offset=None, dense_vec=None, self_interaction=False,
                        skip_gather=True, tail=False, want_dE=True, row_cache=False, row_cache_hint=None):
    _launches[0] += 1
    X = _X(E, table, idx, field_row_offset, dense_vec)
    Fp = X.shape[1]
    ncols = interaction_ncols(Fp, self_interaction, skip_gather)
    d = dOut.float().numpy()
    dX = O.dot_interaction_backward(X, np.ascontiguousarray(d[:, :ncols]), self_interaction, skip_gather, operand_dtype="bf16")
    if dense_vec is None:
        return torch.tensor(dX), None
    dd = dX[:, -1] + (d[:, ncols:] if tail else 0)
    return torch.tensor(np.ascontiguousarray(dX[:, :-1])), torch.tensor(np.ascontiguousarray(dd, dtype=np.float32))


def bucket_by_owner(idx, world, *, L=1, field_row_offset=None, hash_mod=0):
    _launches[0] += 1
    rows = _rows(idx, L, field_row_offset, hash_mod).reshape(-1)
    owner, local = O.shard_of_row(rows, world)
    perm = np.argsort(owner, kind="stable")
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    return (torch.tensor(local[perm]), torch.tensor(perm.astype(np.int32)), torch.tensor(inv.astype(np.int32)),
            torch.tensor(np.bincount(owner, minlength=world).astype(np.int64)))


def _group_slices(g: LookupGroup, D):
    """(indices[n], values[n,D]) of one lookup group, interpreting the rb_grad_source addressing rule."""
    n, L = g.n, g.L
    bags = n // L
    gs = g.grad
    total = None
    for src, bs, ps in zip(gs.srcs, gs.bag_strides, gs.pos_strides):
        v = torch.as_strided(src, (bags, L, D), (bs, ps, 1)).numpy().astype(np.float32)
        total = v.copy() if total is None else total + v
    if gs.scale == "mean":
        total = total / np.float32(L)
    elif gs.scale == "masked_mean":
        mask = (gs.mask_idx if gs.mask_idx is not None else g.idx).numpy().reshape(bags, L) != 0
        total = np.where(mask[..., None], total / gs.count.numpy()[:, None, None], np.float32(0))
    assert gs.fm_g is None, "the fake does not model the FM term"
    rows = _rows(g.idx, L, g.field_row_offset, g.hash_mod).reshape(-1)
    return rows, total.reshape(n, D)


def sparse_bwd_update(table, state0, state1, groups, *, optimizer="adam_lazy", step=1, lr=1e-3, beta_1=0.9, beta_2=0.999,
                      epsilon=1e-7, alpha_dev=None):
    _launches[0] += 1
    D = table.shape[1]
    ind, val = O.concat_indexed_slices([_group_slices(g, D) for g in groups])
    rows, gsum = O.dedup_indexed_slices(ind, val)
    W = table.numpy()
    if optimizer == "adam_lazy":
        O.adam_lazy(W, state0.numpy(), state1.numpy(), rows, gsum, step, lr, beta_1, beta_2, epsilon)
    elif optimizer == "adam_tf_dense":
        O.adam_tf_dense(W, state0.numpy(), state1.numpy(), rows, gsum, step, lr, beta_1, beta_2, epsilon)
    elif optimizer == "adagrad":
        O.adagrad(W, state0.numpy(), rows, gsum, lr, epsilon)
    else:
        O.sgd(W, rows, gsum, lr)


# ---- dense side of the step (host-logic tests run the towers, the loss and the dense optimizers on torch-CPU) ------------

def mm_f32_out(a, b):
    return a.float() @ b.float()


def colsum(x, out=None, ws=None):
    r = x.float().sum(0)
    if out is not None:
        out.copy_(r)
        return out
    return r


def bce_clipped(prob, label, want_grad=True):
    loss, dprob = O.bce_clipped(prob.detach().numpy().astype(np.float32), label.numpy())
    return torch.tensor(np.float32(loss)), (torch.tensor(dprob) if want_grad else None)


def dense_opt_step(params, grads, state0, state1, *, optimizer="adam_lazy", step=1, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
                   alpha_dev=None):
    """Keras `_resource_apply_dense` with torch foreach ops."""
    _launches[0] += 1
    if optimizer in ("adam_lazy", "adam_tf_dense"):
        alpha = adam_alpha_t(lr, beta_1, beta_2, step)
        torch._foreach_mul_(state0, beta_1)
        torch._foreach_add_(state0, grads, alpha=1.0 - beta_1)
        torch._foreach_mul_(state1, beta_2)
        torch._foreach_addcmul_(state1, grads, grads, value=1.0 - beta_2)
        denom = torch._foreach_sqrt(state1)
        torch._foreach_add_(denom, epsilon)
        torch._foreach_addcdiv_(params, state0, denom, value=-alpha)
    elif optimizer == "adagrad":
        torch._foreach_addcmul_(state0, grads, grads, value=1.0)
        denom = torch._foreach_sqrt(state0)
        torch._foreach_add_(denom, epsilon)
        torch._foreach_addcdiv_(params, grads, denom, value=-lr)
    else:
        torch._foreach_add_(params, grads, alpha=-lr)
