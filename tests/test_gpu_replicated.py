"""The entry points behind "small tables are replicated on every rank, large ones row-sharded" (DESIGN §5), each on ONE GPU against
numpy / the unsharded kernels: routing that leaves replicated rows out of every owner bucket, the sharded forward that reads them
from the local copy, the backward that sends their gradient rows to a compact tensor, and the update from all-gathered
(row, summed gradient) lists.  The collective composition of the four runs on real ranks (scripts/p2p_check.py, the 2-GPU test)."""
import ctypes as C

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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_routing_leaves_replicated_rows_out(ops, world):
    rng = np.random.default_rng(world)
    B, F = 777, 6
    rows = [5000, 7, 3000, 40, 9000, 3]                     # fields 1, 3, 5: replicated
    big = [0, 2, 4]
    starts, end = [0] * F, 0
    for t in big:
        starts[t] = end
        end += rows[t]
    small_base = (end + world - 1) // world * world
    off = small_base
    for t in (1, 3, 5):
        starts[t] = off
        off += rows[t]
    ids = np.stack([rng.integers(0, r, size=B) for r in rows], axis=1).astype(np.int64)
    local, perm, inv, counts = ops.bucket_by_owner(cu(ids), world, L=F, field_row_offset=cu(np.array(starts, dtype=np.int64)),
                                                   skip_from_row=small_base)
    local, perm, inv, counts = (t.cpu().numpy() for t in (local, perm, inv, counts))
    g = (ids + np.array(starts)[None]).reshape(-1)
    keep = np.nonzero(g < small_base)[0]                    # positions that have an owner, in position order
    owner = g[keep] % world
    order = keep[np.argsort(owner, kind="stable")]
    n_kept = keep.size
    assert counts.sum() == n_kept
    np.testing.assert_array_equal(counts, np.bincount(owner, minlength=world))
    np.testing.assert_array_equal(perm[:n_kept], order)
    np.testing.assert_array_equal(local[:n_kept], g[order] // world)
    np.testing.assert_array_equal(inv[order], np.arange(n_kept))
    assert (inv[g >= small_base] == -1).all()


@pytest.mark.parametrize("kind", ["adam_lazy", "adagrad", "sgd"])
def test_replicated_rows_update_from_gathered_lists(ops, kind):
    """G sorted, padded (row, summed gradient) lists -> one optimizer update per touched row, gradients added in list order;
    untouched rows (and the optimizer state of untouched rows) do not move."""
    from recommender_b200._lib import check, lib
    rng = np.random.default_rng(5)
    rows, D, G, cap = 300, 32, 4, 120
    W = O.init_table(rng, rows, D)
    lists = np.zeros((G, cap, D + 1), dtype=np.float32)
    total = np.zeros((rows, D), dtype=np.float32)
    touched = np.zeros(rows, dtype=bool)
    for k in range(G):
        n_k = int(rng.integers(1, cap))
        r_k = np.sort(rng.choice(rows, size=n_k, replace=False))
        g_k = rng.normal(0, 1e-2, size=(n_k, D)).astype(np.float32)
        lists[k, :n_k, :D] = g_k
        ids = np.full(cap, rows, dtype=np.int32)            # pads: the scratch id, behind the valid records
        ids[:n_k] = r_k
        lists[k, :, D] = ids.view(np.float32)
        total[r_k] = (total[r_k] + g_k).astype(np.float32)  # rank order
        touched[r_k] = True
    Wd = cu(W.copy())
    if kind == "adam_lazy":
        s0, s1 = torch.zeros_like(Wd), torch.zeros_like(Wd)
    elif kind == "adagrad":
        s0, s1 = torch.full_like(Wd, 0.1), None
    else:
        s0 = s1 = None
    shadow = torch.zeros(rows, D, dtype=torch.bfloat16, device="cuda")
    opt = ops._opt_params(kind, 3, 1e-2)
    check(lib.rb_replicated_rows_update(Wd.data_ptr(), ops._ptr(s0), ops._ptr(s1), rows, D, cu(lists).data_ptr(), G, cap, C.byref(opt),
                                        shadow.data_ptr(), ops._stream()), "rb_replicated_rows_update")
    ref = W.copy()
    r_idx = np.nonzero(touched)[0]
    if kind == "adam_lazy":
        m, v = np.zeros_like(W), np.zeros_like(W)
        O.adam_lazy(ref, m, v, r_idx, total[r_idx], 3, lr=1e-2)
        np.testing.assert_array_equal(s0.cpu().numpy(), m)
        np.testing.assert_array_equal(s1.cpu().numpy(), v)
    elif kind == "adagrad":
        acc = np.full_like(W, 0.1)
        O.adagrad(ref, acc, r_idx, total[r_idx], lr=1e-2)
        np.testing.assert_array_equal(s0.cpu().numpy(), acc)
    else:
        O.sgd(ref, r_idx, total[r_idx], lr=1e-2)
    got = Wd.cpu().numpy()
    np.testing.assert_array_equal(got, ref)                 # explicitly rounded ops: the oracle's bits
    np.testing.assert_array_equal(got[~touched], W[~touched])
    np.testing.assert_array_equal(shadow.float().cpu().numpy()[touched], O.round_bf16(ref)[touched])


@pytest.mark.parametrize("use_shadow", [False, True])
def test_sharded_forward_and_split_backward_with_a_local_replica(ops, use_shadow):
    """world = 1: one 'shard' holds the large tables, a replica the small ones behind `small_base`.  The sharded forward must equal
    the unsharded fused forward over the same rows laid out as ONE table; the split backward must put the small fields' gradient
    rows into the compact tensor and everything else where the plain backward puts it."""
    from recommender_b200 import _lib
    from recommender_b200._lib import check, lib
    rng = np.random.default_rng(9)
    B, F, D = 300, 26, 64
    rows = [4000 if f % 3 else 50 for f in range(F)]        # every third field: a small (replicated) table
    small = [f for f in range(F) if rows[f] == 50]
    big = [f for f in range(F) if rows[f] != 50]
    starts, end = [0] * F, 0
    for f in big:
        starts[f] = end
        end += rows[f]
    small_base = end + 5                                     # a gap: rows between the shard's end and small_base are never looked up
    off = small_base
    for f in small:
        starts[f] = off
        off += rows[f]
    total = off
    full = O.init_table(rng, total, D)
    shard, replica = cu(full[:end].copy()), cu(full[small_base:].copy())
    ids = cu(np.stack([rng.integers(0, r, size=B) for r in rows], axis=1).astype(np.int64))
    offs = cu(np.array(starts, dtype=np.int64))
    dense = cu(rng.normal(0, 0.1, size=(B, D)).astype(np.float32))
    width = 27 * 27 + D
    ref = ops.dot_interaction_fwd(table=cu(full), idx=ids, field_row_offset=offs, dense_vec=dense, tail=True)
    out = torch.empty(B, width, dtype=torch.float32, device="cuda")
    x_saved = torch.empty(B, F + 1, D, dtype=torch.bfloat16, device="cuda")
    shard_ptrs = torch.tensor([shard.data_ptr()], dtype=torch.int64, device="cuda")
    shard16, replica16 = shard.to(torch.bfloat16), replica.to(torch.bfloat16)
    shadow_ptrs = torch.tensor([shard16.data_ptr()], dtype=torch.int64, device="cuda")
    check(lib.rb_dot_interaction_fwd_sharded_rep(shard_ptrs.data_ptr(), 1, total, ids.data_ptr(), _lib.RB_I64, offs.data_ptr(), dense.data_ptr(),
                                                 B, F, D, 0, 1, 1, out.data_ptr(), _lib.RB_F32, width, x_saved.data_ptr(),
                                                 shadow_ptrs.data_ptr() if use_shadow else None, small_base, replica.data_ptr(),
                                                 replica16.data_ptr(), ops._stream()), "rb_dot_interaction_fwd_sharded_rep")
    assert torch.equal(out, ref)                             # bf16 MMA operands either way: the same bits
    X = O.round_bf16(np.concatenate([full[(ids.cpu().numpy() + np.array(starts)[None])], dense.cpu().numpy()[:, None, :]], axis=1))
    np.testing.assert_array_equal(x_saved.float().cpu().numpy(), X)
    # backward
    dout = cu(rng.normal(0, 1e-2, size=(B, width)).astype(np.float32))
    dE_ref, dd_ref = ops.dot_interaction_bwd(dout, table=cu(O.round_bf16(full)), idx=ids, field_row_offset=offs, dense_vec=cu(O.round_bf16(dense.cpu().numpy())),
                                             tail=True)
    slot = np.full(F, -1, dtype=np.int32)
    slot[small] = np.arange(len(small), dtype=np.int32)
    dE = torch.full((B, F, D), float("nan"), device="cuda")
    dE_small = torch.full((B, len(small), D), float("nan"), device="cuda")
    d_dense = torch.empty(B, D, device="cuda")
    check(lib.rb_dot_interaction_bwd_sharded_split(shard_ptrs.data_ptr(), 1, total, ids.data_ptr(), _lib.RB_I64, offs.data_ptr(), dense.data_ptr(),
                                                   B, F, D, 0, 1, 1, dout.data_ptr(), _lib.RB_F32, width, dE.data_ptr(), d_dense.data_ptr(),
                                                   x_saved.data_ptr(), cu(slot).data_ptr(), dE_small.data_ptr(), len(small), ops._stream()),
          "rb_dot_interaction_bwd_sharded_split")
    assert torch.equal(dE[:, big], dE_ref[:, big])
    assert torch.equal(dE_small, dE_ref[:, small])
    assert torch.isnan(dE[:, small]).all()                   # the small fields' slots of dE are not written
    assert torch.equal(d_dense, dd_ref)
    ops.check_oob("cuda")
