"""CUDA kernels against the numpy oracle and the golden fixtures, through the C ABI (ops.py).

Bars (north star): index math and the un-pooled gather bit-exact; pooled outputs fp32 within
1e-6 relative (summation order equals the oracle's); interaction within the bf16-operand bound
and tight against the oracle's bf16-operand mode; updated rows bit-exact for duplicate-free
batches and within fp32 re-association tolerance otherwise."""
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


# ---- K1: gather ---------------------------------------------------------------------------------

@pytest.mark.parametrize("D", [16, 64, 18, 32, 128, 7])
@pytest.mark.parametrize("itype", [np.int64, np.int32])
def test_gather_bit_exact(ops, D, itype):
    rng = np.random.default_rng(D)
    V, B, F = 5000, 333, 26                       # ragged: n = 8658 is not a multiple of the chunk
    W = O.init_table(rng, V, D)
    idx = rng.integers(0, V, size=(B, F)).astype(itype)
    out = ops.gather_fwd(cu(W), cu(idx)).cpu().numpy()
    np.testing.assert_array_equal(out, O.embedding_lookup(W, idx))
    ops.check_oob("cuda")


def test_gather_empty_and_single(ops):
    W = cu(O.init_table(np.random.default_rng(0), 10, 16))
    out = ops.gather_fwd(W, torch.zeros(0, 26, dtype=torch.int64, device="cuda"))
    assert out.shape == (0, 26, 16)
    one = ops.gather_fwd(W, torch.tensor([[9]], device="cuda"))
    np.testing.assert_array_equal(one.cpu().numpy()[0, 0], W.cpu().numpy()[9])


def test_gather_out_of_range_zero_fills_and_flags(ops):
    W = O.init_table(np.random.default_rng(0), 10, 16)
    idx = np.array([[0, 10, -1, 3]], dtype=np.int64)
    out = ops.gather_fwd(cu(W), cu(idx)).cpu().numpy()
    np.testing.assert_array_equal(out[0, 1], 0)
    np.testing.assert_array_equal(out[0, 2], 0)
    np.testing.assert_array_equal(out[0, 3], W[3])
    with pytest.raises(IndexError):
        ops.check_oob("cuda")
    with pytest.raises(IndexError):
        O.embedding_lookup(W, idx)


def test_gather_multi_table_offsets_and_hash(ops):
    rng = np.random.default_rng(1)
    T, V, D, B = 26, 100, 16, 64
    W = O.init_table(rng, T * V, D)
    ids = rng.integers(-2 ** 62, 2 ** 62, size=(B, T), dtype=np.int64)
    off = np.arange(T, dtype=np.int64) * V
    out = ops.gather_fwd(cu(W), cu(ids), L=T, field_row_offset=cu(off), hash_mod=V).cpu().numpy()
    rows = O.id_to_row(ids, V) + off[None]
    np.testing.assert_array_equal(out, W[rows])


@pytest.mark.parametrize("itype", [np.int64, np.int32])
def test_per_table_id_check_flags_an_id_that_would_land_in_the_next_table(ops, itype):
    """the lookups range-check the final row against the total row count only; rb_check_indices is the per-table
    check keras.layers.Embedding makes on CPU (an id >= rows(f) is an InvalidArgument there, not a row of table f + 1)."""
    rng = np.random.default_rng(3)
    table_rows = np.array([7, 100, 3, 50], dtype=np.int64)
    B = 1031                                                            # not a multiple of the block
    ids = np.stack([rng.integers(0, r, size=B) for r in table_rows], axis=1).astype(itype)
    ops.check_oob("cuda")
    ops.check_indices(cu(ids), cu(table_rows))
    ops.check_oob("cuda")                                               # every id inside its own table: nothing raised
    for bad_col, bad_val in ((0, 7), (2, 3), (3, -1), (1, 100)):        # inside the 160-row tensor, outside its own table
        bad = ids.copy()
        bad[B - 1, bad_col] = bad_val
        ops.check_indices(cu(bad), cu(table_rows))
        with pytest.raises(IndexError):
            ops.check_oob("cuda")
    big = rng.integers(0, 2 ** 31 - 1, size=(B, 4)).astype(itype)        # folded first, like the lookups: id mod 3 fits every table
    ops.check_indices(cu(big), cu(table_rows), hash_mod=3)
    ops.check_oob("cuda")
    ops.check_indices(torch.zeros(0, 4, dtype=torch.int64, device="cuda"), cu(table_rows))      # empty batch
    ops.check_oob("cuda")


def test_embedding_validate_ids_raises_like_the_cpu_lookup(ops):
    from recommender_b200.layers import Embedding
    emb = Embedding(10, 16, num_tables=3, device="cuda")
    emb.validate_ids = True
    good = torch.tensor([[0, 9, 5], [3, 3, 3]], device="cuda")
    emb(good)
    ops.check_oob("cuda")
    emb(torch.tensor([[0, 10, 5]], device="cuda"))                      # row 20 of the 30-row tensor: table 2's row 0
    with pytest.raises(IndexError):
        ops.check_oob("cuda")


def test_hash_ids_bit_exact(ops):
    rng = np.random.default_rng(2)
    ids = rng.integers(-2 ** 63, 2 ** 63 - 1, size=4097, dtype=np.int64)
    for vocab, world in ((1_000_000, 8), (39_884_406, 4), (7, 2), (1000, 1)):
        rows, owner, local = ops.hash_ids(cu(ids), vocab, world)
        ref = O.id_to_row(ids, vocab)
        np.testing.assert_array_equal(rows.cpu().numpy(), ref)
        o, l = O.shard_of_row(ref, world)
        np.testing.assert_array_equal(owner.cpu().numpy(), o)
        np.testing.assert_array_equal(local.cpu().numpy(), l)


# ---- K1/K2/K11: pooled lookups --------------------------------------------------------------------

@pytest.mark.parametrize("mode", ["sum", "mean"])
@pytest.mark.parametrize("D,L", [(16, 26), (64, 26), (32, 100), (18, 5), (64, 1)])
def test_bag_pool(ops, mode, D, L):
    rng = np.random.default_rng(L * D)
    V, B = 3000, 257
    W = O.init_table(rng, V, D)
    idx = rng.integers(0, V, size=(B, L)).astype(np.int32)
    out, cnt = ops.bag_pool_fwd(cu(W), cu(idx), mode, want_count=True)
    ref = O.bag_pool(W, idx, mode)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=2e-6, atol=1e-7)
    np.testing.assert_array_equal(cnt.cpu().numpy(), np.full(B, L, np.float32))


def test_masked_mean_golden_and_all_pad_is_nan(ops, golden):
    g = golden("masked_mean")
    D = g["W_item"].shape[1]
    item, cat = cu(g["item"]), cu(g["cat"])
    out = torch.empty(item.shape[0], 2 * D, device="cuda")
    # dien/model.py:14-19: item || cat on the last axis; the item-derived mask gates both tables
    ops.bag_pool_fwd(cu(g["W_item"]), item, "masked_mean", out=out, out_stride=2 * D)
    ops.bag_pool_fwd(cu(g["W_cat"]), cat, "masked_mean", mask_idx=item, out=out[:, D:], out_stride=2 * D)
    np.testing.assert_allclose(out.cpu().numpy(), g["avg"], rtol=1e-5, atol=1e-7)
    W = cu(g["W_item"])
    allpad = torch.zeros(3, 100, dtype=torch.int32, device="cuda")
    res = ops.bag_pool_fwd(W, allpad, "masked_mean")
    assert torch.isnan(res).all()                                    # dien/layers.py:16 has no guard


def test_gather_fm_golden(ops, golden):
    g = golden("deepfm_small")
    E, s, fm = ops.gather_fm_fwd(cu(g["table"]), cu(g["cat"]))
    np.testing.assert_array_equal(E.cpu().numpy(), g["E"])
    np.testing.assert_allclose(s.cpu().numpy(), g["E"].sum(1, dtype=np.float32), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(fm.cpu().numpy(), g["fm"], rtol=2e-4, atol=2e-6)
    ref64 = O.fm_second_order_f64(g["E"])
    np.testing.assert_allclose(fm.cpu().numpy(), ref64, rtol=2e-4, atol=2e-6)


# ---- K3..K6, K10: DotInteraction ----------------------------------------------------------------------

BF16_REL = 2.0 ** -7      # two operands rounded to 8 significant bits, fp32 accumulate


@pytest.mark.parametrize("si", [False, True])
@pytest.mark.parametrize("sg", [False, True])
def test_dot_interaction_golden_all_modes(ops, golden, si, sg):
    g = golden("dot_interaction")
    tag = f"si{int(si)}_sg{int(sg)}"
    X = cu(g["X"])
    out = ops.dot_interaction_fwd(E=X, self_interaction=si, skip_gather=sg).cpu().numpy()
    assert out.shape == g[f"out_{tag}"].shape
    ref_bf = O.dot_interaction(g["X"], si, sg, operand_dtype="bf16")
    np.testing.assert_allclose(out, ref_bf, rtol=1e-5, atol=1e-5)                       # same arithmetic
    scale = np.abs(g[f"out_{tag}"]).max()
    assert np.abs(out - g[f"out_{tag}"]).max() <= BF16_REL * scale                      # vs the fp32 reference
    if sg:
        keep = O.keep_mask(27, si)
        assert (out.reshape(-1, 27, 27)[:, ~keep] == 0).all()                           # exact zeros (ctr/layers.py:37-38)
    dout = cu(g[f"dout_{tag}"])
    dX, _ = ops.dot_interaction_bwd(dout, E=X, self_interaction=si, skip_gather=sg)
    ref_dX_bf = O.dot_interaction_backward(g["X"], g[f"dout_{tag}"], si, sg, operand_dtype="bf16")
    np.testing.assert_allclose(dX.cpu().numpy(), ref_dX_bf, rtol=1e-4, atol=1e-4)
    scale = np.abs(g[f"dX_{tag}"]).max()
    assert np.abs(dX.cpu().numpy() - g[f"dX_{tag}"]).max() <= 2 * BF16_REL * scale


@pytest.mark.parametrize("name", ["dlrm_small", "dlrm_uniform"])
def test_fused_gather_interaction_golden(ops, golden, name):
    """ctr/model.py:49-55 in one kernel: rows straight from the table, bottom-MLP vector as feature 27,
    mask, zero fill and the '|| bmlp' tail."""
    g = golden(name)
    D = g["table"].shape[1]
    bmlp = np.ascontiguousarray(g["X"][:, 26])
    out = ops.dot_interaction_fwd(table=cu(g["table"]), idx=cu(g["cat"]), dense_vec=cu(bmlp), tail=True).cpu().numpy()
    assert out.shape == (g["cat"].shape[0], 729 + D)
    ref_bf = O.dot_interaction(g["X"], False, True, operand_dtype="bf16")
    np.testing.assert_allclose(out[:, :729], ref_bf, rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(out[:, 729:], bmlp)
    assert np.abs(out[:, :729] - g["inter"]).max() <= BF16_REL * np.abs(g["inter"]).max()
    # un-fused path (materialised E) gives the identical bits
    out2 = ops.dot_interaction_fwd(E=cu(np.ascontiguousarray(g["X"][:, :26])), dense_vec=cu(bmlp), tail=True).cpu().numpy()
    np.testing.assert_array_equal(out, out2)
    # backward: dX split into dE and d(bmlp) (+ the direct tail gradient)
    rng = np.random.default_rng(0)
    dtail = rng.normal(0, 1e-2, size=(out.shape[0], D)).astype(np.float32)
    dOut = np.concatenate([g["dinter"], dtail], axis=1)
    dE, d_dense = ops.dot_interaction_bwd(cu(dOut), table=cu(g["table"]), idx=cu(g["cat"]), dense_vec=cu(bmlp), tail=True)
    ref = O.dot_interaction_backward(g["X"], g["dinter"], False, True, operand_dtype="bf16")
    tol = dict(rtol=1e-4, atol=1e-4 * np.abs(ref).max())
    np.testing.assert_allclose(dE.cpu().numpy(), ref[:, :26], **tol)
    np.testing.assert_allclose(d_dense.cpu().numpy(), ref[:, 26] + dtail, **tol)
    assert np.abs(dE.cpu().numpy() - g["dX"][:, :26]).max() <= 2 * BF16_REL * np.abs(g["dX"]).max()


@pytest.mark.parametrize("D", [16, 32, 64, 128])
@pytest.mark.parametrize("F", [1, 8, 26, 31])
def test_dot_interaction_shapes(ops, D, F):
    rng = np.random.default_rng(D + F)
    B = 37                                            # not a multiple of the warps per CTA
    E = rng.normal(0, 0.3, size=(B, F, D)).astype(np.float32)
    dv = rng.normal(0, 0.3, size=(B, D)).astype(np.float32)
    X = np.concatenate([E, dv[:, None]], axis=1)
    out = ops.dot_interaction_fwd(E=cu(E), dense_vec=cu(dv), tail=True).cpu().numpy()
    Fp = F + 1
    ref = O.dot_interaction(X, False, True, operand_dtype="bf16")
    np.testing.assert_allclose(out[:, :Fp * Fp], ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(out[:, Fp * Fp:], dv)


def test_interaction_bf16_rows_for_the_top_mlp(ops, golden):
    """The bf16 form of the interaction row (what a bf16 top MLP consumes): the same values rounded
    to bf16, zero pad columns up to a multiple of 8, and a backward that takes the padded bf16 dOut."""
    g = golden("dlrm_uniform")
    D = g["table"].shape[1]
    B = g["cat"].shape[0]
    bmlp = np.ascontiguousarray(g["X"][:, 26])
    table, cat, dv = cu(g["table"]), cu(g["cat"]), cu(bmlp)
    ref32 = ops.dot_interaction_fwd(table=table, idx=cat, dense_vec=dv, tail=True)
    out = ops.dot_interaction_fwd(table=table, idx=cat, dense_vec=dv, tail=True, out_dtype=torch.bfloat16, pad_to=8)
    width = 729 + D
    stride = (width + 7) // 8 * 8
    assert out.dtype == torch.bfloat16 and out.shape == (B, stride)
    assert torch.equal(out[:, :width], ref32.to(torch.bfloat16))            # identical arithmetic, one final rounding
    assert (out[:, width:] == 0).all()                                       # GEMM-ready pad columns
    ones = ops.dot_interaction_fwd(table=table, idx=cat, dense_vec=dv, tail=True, out_dtype=torch.bfloat16, pad_to=8, ones_col=True)
    assert torch.equal(ones[:, :width], out[:, :width])
    if stride > width:                                                       # first pad column carries 1.0, the rest stay 0
        assert (ones[:, width] == 1).all() and (ones[:, width + 1:] == 0).all()
    # backward with the padded bf16 gradient == backward with the same values in fp32
    rng = np.random.default_rng(3)
    dpad = torch.zeros(B, stride, dtype=torch.bfloat16, device="cuda")
    dpad[:, :width] = cu(rng.normal(0, 1e-2, size=(B, width)).astype(np.float32)).to(torch.bfloat16)
    dE_b, dd_b = ops.dot_interaction_bwd(dpad, table=table, idx=cat, dense_vec=dv, tail=True)
    dE_f, dd_f = ops.dot_interaction_bwd(dpad[:, :width].float().contiguous(), table=table, idx=cat, dense_vec=dv, tail=True)
    assert torch.equal(dE_b, dE_f) and torch.equal(dd_b, dd_f)
    ref = O.dot_interaction_backward(g["X"], dpad[:, :729].float().cpu().numpy(), False, True, operand_dtype="bf16")
    np.testing.assert_allclose(dE_b.cpu().numpy(), ref[:, :26], rtol=1e-4, atol=1e-4 * np.abs(ref).max())


@pytest.mark.parametrize("B", [1, 7, 1184 * 2 + 5])
def test_interaction_ragged_batches_and_unaligned_rows(ops, B):
    """Persistent-grid edge cases: fewer samples than warps, a ragged last round, and fp32 rows whose
    start is only 4-byte aligned (width 793 is odd), written into a strided output."""
    rng = np.random.default_rng(B)
    V, F, D = 500, 26, 64
    W = O.init_table(rng, V, D)
    idx = rng.integers(0, V, size=(B, F)).astype(np.int64)
    dv = rng.normal(0, 0.1, size=(B, D)).astype(np.float32)
    X = np.concatenate([W[idx], dv[:, None]], axis=1)
    ref = O.dot_interaction(X, False, True, operand_dtype="bf16")
    big = torch.full((B, 801), -7.0, device="cuda")
    ops.dot_interaction_fwd(table=cu(W), idx=cu(idx), dense_vec=cu(dv), tail=True, out=big, out_stride=801)
    got = big.cpu().numpy()
    np.testing.assert_allclose(got[:, :729], ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(got[:, 729:793], dv)
    assert (got[:, 793:] == -7.0).all()                                      # nothing outside the row is touched
    dOut = rng.normal(0, 1e-2, size=(B, 793)).astype(np.float32)
    dE, dd = ops.dot_interaction_bwd(cu(dOut), table=cu(W), idx=cu(idx), dense_vec=cu(dv), tail=True)
    refd = O.dot_interaction_backward(X, np.ascontiguousarray(dOut[:, :729]), False, True, operand_dtype="bf16")
    tol = dict(rtol=1e-4, atol=1e-4 * np.abs(refd).max())
    np.testing.assert_allclose(dE.cpu().numpy(), refd[:, :26], **tol)
    np.testing.assert_allclose(dd.cpu().numpy(), refd[:, 26] + dOut[:, 729:], **tol)


def test_interaction_out_of_range_ids_read_zero_rows(ops):
    """TF's GPU gather writes zeros for out-of-range ids; the fused form must agree with
    interaction over the zero-filled E."""
    rng = np.random.default_rng(9)
    V, F, D, B = 50, 26, 32, 33
    W = O.init_table(rng, V, D)
    idx = rng.integers(0, V, size=(B, F)).astype(np.int64)
    idx[3, 5], idx[10, 0] = V, -1
    dv = rng.normal(0, 0.1, size=(B, D)).astype(np.float32)
    E = ops.gather_fwd(cu(W), cu(idx))
    torch.cuda.synchronize()
    ops.oob_flag("cuda").zero_()
    a = ops.dot_interaction_fwd(table=cu(W), idx=cu(idx), dense_vec=cu(dv), tail=True)
    b = ops.dot_interaction_fwd(E=E, dense_vec=cu(dv), tail=True)
    assert torch.equal(a, b)


# ---- K7..K9: backward scatter + sparse optimizers -----------------------------------------------------------

def _state(kind, W):
    if kind.startswith("adam"):
        return dict(m=np.zeros_like(W), v=np.zeros_like(W))
    if kind == "adagrad":
        return dict(acc=np.full_like(W, 0.1))
    return {}


def _run_update(ops, W, st, idx, dE, kind, step, L=None, **hp):
    from recommender_b200.ops import GradSource, LookupGroup
    Wt = cu(W)
    if kind.startswith("adam"):
        s0, s1 = cu(st["m"]), cu(st["v"])
    elif kind == "adagrad":
        s0, s1 = cu(st["acc"]), None
    else:
        s0 = s1 = None
    L = L or idx.shape[-1]
    ops.sparse_bwd_update(Wt, s0, s1, [LookupGroup(cu(idx), L, GradSource.per_position(cu(dE), L))], optimizer=kind,
                          step=step, **hp)
    res = dict(W=Wt.cpu().numpy())
    if kind.startswith("adam"):
        res.update(m=s0.cpu().numpy(), v=s1.cpu().numpy())
    elif kind == "adagrad":
        res.update(acc=s0.cpu().numpy())
    return res


@pytest.mark.parametrize("kind", ["adam_lazy", "adam_tf_dense", "adagrad", "sgd"])
@pytest.mark.parametrize("D", [16, 64])
def test_sparse_update_unique_rows_bit_exact(ops, kind, D):
    """No duplicate rows -> no summation-order freedom: every updated element must match numpy's
    correctly-rounded fp32 arithmetic bit for bit (the kernel uses explicit _rn ops, no FMA)."""
    rng = np.random.default_rng(D)
    V, B, F = 4000, 100, 26
    W = O.init_table(rng, V, D)
    idx = rng.permutation(V)[: B * F].reshape(B, F).astype(np.int64)
    st = _state(kind, W)
    hp = dict(lr=1e-2) if kind == "sgd" else {}
    for step in (1, 2):
        dE = rng.normal(0, 1e-3, size=(B, F, D)).astype(np.float32)
        got = _run_update(ops, W, st, idx, dE, kind, step, **hp)
        O.sparse_backward_update(W, st, idx, dE, kind, step, **hp)
        np.testing.assert_array_equal(got["W"], W)
        for k in st:
            np.testing.assert_array_equal(got[k], st[k])


@pytest.mark.parametrize("kind", ["adam_lazy", "adagrad"])
@pytest.mark.parametrize("dist", ["uniform", "zipf"])
def test_sparse_update_with_duplicates(ops, kind, dist):
    """Duplicate rows (incl. the OOV -> 0 hot row): the sums associate per 32-entry tile instead of
    strictly left to right, so allow fp32 re-association error on the summed gradient."""
    V, B, D = 2000, 512, 16
    cat, _, _ = O.synth_batch(B, V, seed=4, dist=dist)
    rng = np.random.default_rng(5)
    W = O.init_table(rng, V, D)
    st = _state(kind, W)
    for step in (1, 2, 3):
        dE = rng.normal(0, 1e-3, size=(B, 26, D)).astype(np.float32)
        got = _run_update(ops, W, st, cat, dE, kind, step)
        O.sparse_backward_update(W, st, cat, dE, kind, step)
        np.testing.assert_allclose(got["W"], W, rtol=0, atol=2e-6)          # |update| <= ~lr = 1e-3
        for k in st:
            np.testing.assert_allclose(got[k], st[k], rtol=2e-4, atol=1e-9)
        W, st = got["W"], {k: got[k] for k in st}                            # continue from the CUDA state


def test_sparse_update_is_deterministic(ops):
    V, B, D = 500, 2048, 64
    cat, _, _ = O.synth_batch(B, V, seed=9, dist="zipf")
    rng = np.random.default_rng(1)
    W = O.init_table(rng, V, D)
    dE = rng.normal(0, 1e-3, size=(B, 26, D)).astype(np.float32)
    st = _state("adam_lazy", W)
    a = _run_update(ops, W, st, cat, dE, "adam_lazy", 1)
    b = _run_update(ops, W, st, cat, dE, "adam_lazy", 1)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])


def test_dedup_matches_oracle(ops):
    from recommender_b200.ops import GradSource
    V, B, D = 300, 700, 32
    cat, _, _ = O.synth_batch(B, V, seed=3, dist="zipf")
    rng = np.random.default_rng(2)
    dE = rng.normal(0, 1.0, size=(B, 26, D)).astype(np.float32)
    rows, summed = ops.sparse_bwd_dedup(V, D, cu(cat), 26, GradSource.per_position(cu(dE), 26))
    rows, summed = rows.cpu().numpy(), summed.cpu().numpy()
    ref_rows, ref_sum = O.dedup_indexed_slices(*O.gather_backward(cat, dE))
    order = np.argsort(ref_rows)                     # the CUDA path emits rows ascending, TF first-occurrence
    np.testing.assert_array_equal(rows, ref_rows[order])
    _, ref64 = O.dedup_indexed_slices_f64(*O.gather_backward(cat, dE))
    counts = np.bincount(cat.reshape(-1), minlength=V)[rows]
    tol = 1e-6 * np.sqrt(counts)[:, None] * 4 + 1e-6
    assert (np.abs(summed - ref64) <= tol * np.maximum(1.0, np.abs(ref64))).all()
    np.testing.assert_allclose(summed, ref_sum[order], rtol=1e-3, atol=2e-5)


def test_hot_row_long_chain(ops):
    """One row receiving > 64 tiles of gradients (ESMM's 3-row tables, OOV id 0) takes the CTA-wide path."""
    from recommender_b200.ops import GradSource
    V, D, n = 3, 32, 40000
    rng = np.random.default_rng(6)
    idx = rng.integers(0, V, size=(n, 1)).astype(np.int32)
    dE = rng.normal(0, 1.0, size=(n, 1, D)).astype(np.float32)
    rows, summed = ops.sparse_bwd_dedup(V, D, cu(idx), 1, GradSource.per_position(cu(dE), 1))
    _, ref64 = O.dedup_indexed_slices_f64(*O.gather_backward(idx, dE))
    assert rows.cpu().numpy().tolist() == [0, 1, 2]
    np.testing.assert_allclose(summed.cpu().numpy(), ref64, rtol=1e-4, atol=2e-3)


def test_masked_mean_backward_golden(ops, golden):
    """config 4: bag-level gradient + mask + count -> table gradient, never materialising [B,L,D]."""
    from recommender_b200.ops import GradSource, LookupGroup
    g = golden("masked_mean")
    D = g["W_item"].shape[1]
    item, cat = cu(g["item"]), cu(g["cat"])
    davg = cu(g["davg"])
    _, count = ops.bag_pool_fwd(cu(g["W_item"]), item, "masked_mean", want_count=True)
    for W, idx, col, key in ((g["W_item"], item, 0, "dW_item"), (g["W_cat"], cat, D, "dW_cat")):
        Wt = cu(W)
        src = GradSource([davg[:, col:]], [davg.stride(0)], [0], scale="masked_mean", mask_idx=item, count=count)
        ops.sparse_bwd_update(Wt, None, None, [LookupGroup(idx, 100, src)], optimizer="sgd", lr=1.0)
        np.testing.assert_allclose(W - Wt.cpu().numpy(), g[key], rtol=1e-4, atol=1e-8)


def test_multi_consumer_and_multi_table_golden(ops, golden):
    """config 5: one table per feature, bag size 1, the consumers' gradients added left to right
    inside the scatter (esmm/esmm.py:15-24)."""
    from recommender_b200.ops import GradSource, LookupGroup
    g = golden("esmm_small")
    feats = [str(f) for f in g["feats"]]
    D = g[f"W_{feats[0]}"].shape[1]
    B = g[f"idx_{feats[0]}"].shape[0]
    width = D * len(feats)
    out = torch.empty(B, width, device="cuda")
    for k, f in enumerate(feats):                                  # concat on the last axis, written in place
        ops.gather_fwd(cu(g[f"W_{f}"]), cu(g[f"idx_{f}"][:, 0]), out=out[:, k * D:], out_stride=width)
    np.testing.assert_array_equal(out.cpu().numpy(), g["emb"])
    rng = np.random.default_rng(8)
    consumers = [rng.normal(0, 1e-2, size=(B, width)).astype(np.float32) for _ in range(10)]    # MMOE: 10 consumers
    total = O.multi_consumer_grad(consumers)
    cons_t = [cu(c) for c in consumers]
    for k, f in enumerate(feats):
        W = g[f"W_{f}"]
        Wt = cu(W)
        src = GradSource([c[:, k * D:] for c in cons_t], [width] * 10, [0] * 10)
        ops.sparse_bwd_update(Wt, None, None, [LookupGroup(cu(g[f"idx_{f}"]), 1, src)], optimizer="sgd", lr=1.0)
        ref = W.copy()
        O.sparse_backward_update(ref, {}, g[f"idx_{f}"], total[:, None, k * D:(k + 1) * D], "sgd", lr=1.0)
        np.testing.assert_allclose(Wt.cpu().numpy(), ref, rtol=0, atol=1e-6)


def test_two_uses_of_one_table_are_concatenated(ops):
    """dien/model.py:26-30: item_embedding serves the target item and the history; TF concatenates
    the two IndexedSlices before the duplicate-row sum, so Adam sees ONE summed gradient per row."""
    from recommender_b200.ops import GradSource, LookupGroup
    rng = np.random.default_rng(10)
    V, D, B, L = 50, 32, 64, 10
    W = O.init_table(rng, V, D)
    tgt = rng.integers(1, V, size=(B, 1)).astype(np.int32)
    his = rng.integers(0, V, size=(B, L)).astype(np.int32)
    d_tgt = rng.normal(0, 1e-3, size=(B, 1, D)).astype(np.float32)
    d_avg = rng.normal(0, 1e-3, size=(B, D)).astype(np.float32)
    mask = his != 0
    st = _state("adam_lazy", W)
    Wt, m, v = cu(W), cu(st["m"]), cu(st["v"])
    _, count = ops.bag_pool_fwd(Wt, cu(his), "masked_mean", want_count=True)
    groups = [LookupGroup(cu(tgt), 1, GradSource.per_position(cu(d_tgt), 1)),
              LookupGroup(cu(his), L, GradSource.per_bag([cu(d_avg)], scale="masked_mean", mask_idx=cu(his), count=count))]
    ops.sparse_bwd_update(Wt, m, v, groups, optimizer="adam_lazy", step=1)
    d_his = O.masked_mean_backward(d_avg, mask)
    ind, val = O.concat_indexed_slices([O.gather_backward(tgt, d_tgt), O.gather_backward(his, d_his)])
    rows, gsum = O.dedup_indexed_slices(ind, val)
    O.adam_lazy(W, st["m"], st["v"], rows, gsum, 1)
    np.testing.assert_allclose(Wt.cpu().numpy(), W, rtol=0, atol=2e-6)
    np.testing.assert_allclose(v.cpu().numpy(), st["v"], rtol=1e-3, atol=1e-12)


@pytest.mark.parametrize("optimizer", ["adam_lazy", "adam_tf_dense", "adagrad"])
def test_masked_pads_collapse_keeps_touched_rows(ops, optimizer):
    """config 4 over several steps: the one-call update collapses each run of masked pads into one zero pair
    (make_keys_kernel).  TF lists every pad in the IndexedSlices (dien/layers.py:13 multiplies, it does not drop), so
    the pad row stays *touched*: with lazy Adam its moments decay and the row moves although its gradient is zero.  The
    cat table is masked by the item ids (dien/model.py:25) and its pad row 0 also occurs at valid positions, so the row
    carries state when the pads touch it.  Interior (non-trailing) masks are in the batch."""
    from recommender_b200.ops import GradSource, LookupGroup
    rng = np.random.default_rng(21)
    V_item, V_cat, D, B, L = 300, 7, 32, 96, 100
    W = O.init_table(rng, V_cat, D)
    st = _state(optimizer, W)
    Wt = cu(W)
    s0 = cu(st["m"] if "m" in st else st["acc"]) if st else None
    s1 = cu(st["v"]) if "v" in st else None
    for step in (1, 2, 3):
        lens = rng.integers(1, L + 1, size=(B, 1))
        item = rng.integers(1, V_item, size=(B, L)).astype(np.int32)
        cat = rng.integers(0, V_cat, size=(B, L)).astype(np.int32)
        valid = np.arange(L)[None] < lens
        item, cat = item * valid, cat * valid                   # trailing zeros (dien/data_loader.py:44)
        item[3, 2:5] = 0                                        # interior pads; their cat rows differ from one another
        item[5, 1:] = 0                                         # one valid position, 99 masked ones with random cat rows
        if step == 2:
            cat[cat == 0] = 1                                   # row 0 is reached through pads only in this step
            cat = cat * valid
        mask = item != 0
        d_avg = rng.normal(0, 1e-3, size=(B, D)).astype(np.float32)
        count = cu(mask.sum(1).astype(np.float32))
        grp = LookupGroup(cu(cat), L, GradSource.per_bag([cu(d_avg)], scale="masked_mean", mask_idx=cu(item), count=count))
        ops.sparse_bwd_update(Wt, s0, s1, [grp], optimizer=optimizer, step=step)
        ops.check_oob("cuda")
        O.sparse_backward_update(W, st, cat, O.masked_mean_backward(d_avg, mask), optimizer, step=step)
        np.testing.assert_allclose(Wt.cpu().numpy(), W, rtol=0, atol=3e-6, err_msg=f"step {step}")
    if s1 is not None:
        np.testing.assert_allclose(s1.cpu().numpy(), st["v"], rtol=1e-3, atol=1e-12)


def test_bucket_by_owner(ops):
    rng = np.random.default_rng(11)
    n, V = 10007, 1000
    ids = rng.integers(0, V, size=n).astype(np.int64)
    for world in (1, 2, 4, 8, 3):
        local, perm, inv, counts = ops.bucket_by_owner(cu(ids), world)
        local, perm, inv, counts = (t.cpu().numpy() for t in (local, perm, inv, counts))
        owner, loc = O.shard_of_row(ids, world)
        ref_perm = np.argsort(owner, kind="stable")
        np.testing.assert_array_equal(perm, ref_perm)
        np.testing.assert_array_equal(local, loc[ref_perm])
        np.testing.assert_array_equal(inv[perm], np.arange(n))
        np.testing.assert_array_equal(counts, np.bincount(owner, minlength=world))


# ---- dense side: multi-tensor optimizer step, bias gradient --------------------------------------------------

@pytest.mark.parametrize("kind", ["adam_lazy", "adagrad", "sgd"])
def test_dense_opt_step_matches_keras_formulas(ops, kind):
    """rb_dense_opt_step over tensors of very different sizes (1 .. 406016 elements, 40 of them, so two
    launches) against the oracle's Keras `_resource_apply_dense` restatement — bit for bit."""
    rng = np.random.default_rng(11)
    sizes = [1, 7, 512, 2048, 2049, 793 * 512, 64] + [33] * 33
    P = [rng.normal(0, 0.1, size=n).astype(np.float32) for n in sizes]
    G = [rng.normal(0, 1e-2, size=n).astype(np.float32) for n in sizes]
    M = [rng.normal(0, 1e-3, size=n).astype(np.float32) for n in sizes]
    V = [np.abs(rng.normal(0, 1e-5, size=n)).astype(np.float32) for n in sizes]
    p, g, m, v = ([cu(a.copy()) for a in L] for L in (P, G, M, V))
    step = 3
    if kind == "adam_lazy":
        ops.dense_opt_step(p, g, m, v, optimizer=kind, step=step)
        for k in range(len(sizes)):
            rp, rm, rv = P[k].copy(), M[k].copy(), V[k].copy()
            O.adam_dense_param(rp, rm, rv, G[k], step)
            np.testing.assert_array_equal(p[k].cpu().numpy(), rp)
            np.testing.assert_array_equal(m[k].cpu().numpy(), rm)
            np.testing.assert_array_equal(v[k].cpu().numpy(), rv)
    elif kind == "adagrad":
        acc = [cu(np.full(n, 0.1, np.float32)) for n in sizes]
        ops.dense_opt_step(p, g, acc, None, optimizer=kind, lr=1e-3)
        for k in range(len(sizes)):
            a = np.full(sizes[k], 0.1, np.float32) + G[k] * G[k]
            ref = P[k] - (np.float32(1e-3) * G[k]) / (np.sqrt(a) + np.float32(1e-7))
            np.testing.assert_array_equal(acc[k].cpu().numpy(), a)
            np.testing.assert_array_equal(p[k].cpu().numpy(), ref.astype(np.float32))
    else:
        ops.dense_opt_step(p, g, None, None, optimizer=kind, lr=1e-2)
        for k in range(len(sizes)):
            np.testing.assert_array_equal(p[k].cpu().numpy(), P[k] - np.float32(1e-2) * G[k])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("rows,cols", [(65536, 512), (1000, 64), (63, 8), (4097, 256), (5, 2048)])
def test_colsum_bias_gradient(ops, dtype, rows, cols):
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    x = (torch.randn(rows, cols, device="cuda", generator=g) * 1e-2).to(dtype)
    got = ops.colsum(x)
    ref = x.double().sum(0)
    scale = x.double().abs().sum(0).max().item()
    assert (got.double() - ref).abs().max().item() <= 1e-6 * scale + 1e-12
    assert torch.equal(got, ops.colsum(x))                                   # deterministic


# ---- sharded table over peer memory: G emulated ranks on one GPU (p2p.LocalPeerLink) ------------------------------

@pytest.mark.parametrize("G", [2, 4, 8])
@pytest.mark.parametrize("T,shadow", [(1, True), (26, True), (26, False)])
def test_peer_memory_sharded_step_equals_unsharded(ops, G, T, shadow):
    """Row-wise sharding with the exchange done by the kernels (rb_dot_interaction_*_sharded read the rows from
    the owners' shards, rb_sparse_bwd_apply_p2p pulls gradient rows from every rank's dE): interaction
    outputs, dE and the updated table must equal the unsharded kernels bit for bit — same pairs, same
    (rank-major, position) order in the segmented reduction."""
    from recommender_b200.ops import GradSource, LookupGroup
    from recommender_b200.p2p import LocalPeerLink, P2PShardedEmbedding
    rng = np.random.default_rng(100 * G + T)
    V, D, B, F = (50021 if T == 1 else 2003), 32, 96, 26
    W = O.init_table(rng, V * T, D)
    registry = {}
    embs = [P2PShardedEmbedding(V, D, num_tables=T, link=LocalPeerLink(G, r, registry), device="cuda", capacity_factor=float(G),
                                bf16_shadow=shadow) for r in range(G)]
    for e in embs:
        e.load_full_table(torch.tensor(W))
    off = cu(np.arange(T, dtype=np.int64) * V) if T > 1 else None
    idx = [cu(np.where(rng.random((B, F)) < 0.3, 0, rng.integers(0, V, size=(B, F))).astype(np.int64)) for _ in range(G)]  # hot row 0
    dense = [cu(rng.normal(0, 0.1, size=(B, D)).astype(np.float32)) for _ in range(G)]
    dOut = [cu(rng.normal(0, 1e-2, size=(B, 27 * 27 + D)).astype(np.float32)) for _ in range(G)]
    Wd = cu(W)
    # lock-step phases: every rank routes, then every owner collects + sorts
    for r in range(G):
        embs[r].route(idx[r])
    for r in range(G):
        embs[r].collect_and_sort()
    groups = []
    for r in range(G):
        out = embs[r]._interaction_fwd(idx[r], dense[r], (False, True, True), torch.float32, 1)
        ref = ops.dot_interaction_fwd(table=Wd, idx=idx[r], field_row_offset=off, dense_vec=dense[r], tail=True)
        assert torch.equal(out, ref)
        d_dense = embs[r]._interaction_bwd(idx[r], dense[r], (False, True, True), dOut[r])
        dE_ref, dd_ref = ops.dot_interaction_bwd(dOut[r], table=Wd, idx=idx[r], field_row_offset=off, dense_vec=dense[r], tail=True)
        assert torch.equal(embs[r]._dE.view(B, F, D), dE_ref) and torch.equal(d_dense, dd_ref)
        groups.append(LookupGroup(idx[r], F, GradSource.per_position(dE_ref, F), field_row_offset=off))
    for r in range(G):
        embs[r]._routed_by_caller = True
        embs[r].apply_pending("adam_lazy", 3, 1e-3)
        embs[r].check_overflow()
    # reference: the same IndexedSlices concatenated in rank order, one unsharded update (<= 4 groups per call:
    # for G = 8 materialise the concatenation instead)
    m, v = torch.zeros_like(Wd), torch.zeros_like(Wd)
    if G <= 4:
        ops.sparse_bwd_update(Wd, m, v, groups, optimizer="adam_lazy", step=3)
    else:
        rows = torch.cat([(idx[r] + (off[None] if off is not None else 0)).reshape(-1) for r in range(G)])
        dE_all = torch.cat([g.grad.srcs[0].reshape(-1, D) for g in groups])
        ops.sparse_bwd_update(Wd, m, v, [LookupGroup(rows, 1, GradSource.per_position(dE_all, 1))], optimizer="adam_lazy", step=3)
    # rows hit once are bit-exact; duplicate rows are summed in the same (rank, position) order but the tile
    # boundaries of the segmented reduction fall elsewhere, so long runs may re-associate (fp32)
    got_W, got_m = torch.zeros_like(Wd), torch.zeros_like(Wd)
    for r in range(G):
        embs[r].scatter_into_full(got_W)
        embs[r].scatter_into_full(got_m, embs[r].opt_state["m"])
    np.testing.assert_allclose(got_W.cpu().numpy(), Wd.cpu().numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(got_m.cpu().numpy(), m.cpu().numpy(), rtol=1e-5, atol=1e-9)
    touched = (m.abs().sum(1) > 0)
    assert 0 < int(touched.sum()) < Wd.shape[0]
    assert torch.equal(got_m.abs().sum(1) > 0, touched)                                 # exactly the same rows moved
    if shadow:  # the bf16 shadow the other ranks read stays in step with the fp32 shard
        for e in embs:
            assert torch.equal(e._shadow_full[: e.local_rows], e.embeddings.to(torch.bfloat16))
    if T > 1:   # the table pitch spreads the hot id 0 of the 26 tables over the ranks
        owners = {int((t * embs[0].pitch) % G) for t in range(T)}
        assert len(owners) == min(G, T)


def test_peer_memory_capacity_overflow_is_loud(ops):
    """Every lookup of both emulated ranks addressed to owner 0 (id 0 of one shared table): twice the mean load against a
    capacity factor of 1.25.  The collect must raise the flag, the synchronising check and the non-blocking poll (pinned
    mirror of the flag, read by the next step) must both raise, and a run with room must not."""
    from recommender_b200.p2p import LocalPeerLink, P2PShardedEmbedding
    G, V, D, B, F = 2, 4096, 32, 1024, 26
    for factor, expect in ((1.25, True), (2.0, False)):
        registry = {}
        embs = [P2PShardedEmbedding(V, D, link=LocalPeerLink(G, r, registry), device="cuda", capacity_factor=factor) for r in range(G)]
        idx = [torch.zeros(B, F, dtype=torch.int64, device="cuda") for _ in range(G)]
        for r in range(G):
            embs[r].route(idx[r])
        for r in range(G):
            embs[r].collect_and_sort()
        torch.cuda.synchronize()
        if expect:
            with pytest.raises(RuntimeError, match="capacity"):
                embs[0].poll_overflow()
            embs[0]._overflow_host.fill_(1)
            with pytest.raises(RuntimeError, match="capacity"):
                embs[0].check_overflow()
        else:
            embs[0].poll_overflow()
            embs[0].check_overflow()
        embs[1].poll_overflow()          # owner 1 received nothing
        embs[1].check_overflow()


@pytest.mark.parametrize("G", [2, 8])
def test_peer_memory_sharding_of_unequal_tables(ops, G):
    """BASELINE config 3 in miniature: 26 tables with capped Criteo-Terabyte-like cardinalities from 3 rows to a few
    thousand, raw ids folded per table (cap_ids), row-wise sharded over G emulated ranks — against the unsharded
    Embedding(table_rows=...) on the same ids."""
    from recommender_b200.layers import Embedding
    from recommender_b200.ops import GradSource, LookupGroup
    from recommender_b200.p2p import LocalPeerLink, P2PShardedEmbedding
    rng = np.random.default_rng(G)
    cards = [3989, 391, 173, 75, 203, 3, 72, 16, 63, 3854, 2954, 404, 10, 23, 120, 155, 4, 976, 14, 3998, 2565, 3967, 586, 130, 108, 36]
    D, B, F = 32, 128, 26
    ref = Embedding(0, D, table_rows=cards, device="cuda")
    W = ref.embeddings.clone()
    registry = {}
    embs = [P2PShardedEmbedding(0, D, link=LocalPeerLink(G, r, registry), device="cuda", capacity_factor=float(G), table_rows=cards)
            for r in range(G)]
    for e in embs:
        e.load_full_table(W)
        assert int((e.full_row_ids() >= 0).sum()) > 0
    assert sum(int((e.full_row_ids() >= 0).sum()) for e in embs) == sum(cards)
    raw = [cu(rng.integers(0, 2 ** 40, size=(B, F)).astype(np.int64)) for _ in range(G)]
    idx = [ref.cap_ids(r_) for r_ in raw]
    for r in range(G):
        assert torch.equal(embs[r].cap_ids(raw[r]), idx[r])
    dense = [cu(rng.normal(0, 0.1, size=(B, D)).astype(np.float32)) for _ in range(G)]
    dOut = [cu(rng.normal(0, 1e-2, size=(B, 27 * 27 + D)).astype(np.float32)) for _ in range(G)]
    off = ref.row_offset_for(F)
    for r in range(G):
        embs[r].route(idx[r])
    for r in range(G):
        embs[r].collect_and_sort()
    groups = []
    for r in range(G):
        out = embs[r]._interaction_fwd(idx[r], dense[r], (False, True, True), torch.float32, 1)
        assert torch.equal(out, ops.dot_interaction_fwd(table=W, idx=idx[r], field_row_offset=off, dense_vec=dense[r], tail=True))
        embs[r]._interaction_bwd(idx[r], dense[r], (False, True, True), dOut[r])
        dE_ref, _ = ops.dot_interaction_bwd(dOut[r], table=W, idx=idx[r], field_row_offset=off, dense_vec=dense[r], tail=True)
        assert torch.equal(embs[r]._dE.view(B, F, D), dE_ref)
        groups.append((idx[r], dE_ref))
    for r in range(G):
        embs[r]._routed_by_caller = True
        embs[r].apply_pending("adam_lazy", 1, 1e-3)
        embs[r].check_overflow()
    rows = torch.cat([(i_ + off[None]).reshape(-1) for i_, _ in groups])
    dE_all = torch.cat([d_.reshape(-1, D) for _, d_ in groups])
    Wd, m, v = W.clone(), torch.zeros_like(W), torch.zeros_like(W)
    ops.sparse_bwd_update(Wd, m, v, [LookupGroup(rows, 1, GradSource.per_position(dE_all, 1))], optimizer="adam_lazy", step=1)
    got = torch.zeros_like(Wd)
    for e in embs:
        e.scatter_into_full(got)
    np.testing.assert_allclose(got.cpu().numpy(), Wd.cpu().numpy(), rtol=0, atol=2e-6)   # 3-row tables: runs of ~B*G/3 duplicates
    assert int(sum(int(e._n_valid.item()) for e in embs)) == G * B * F


@pytest.mark.parametrize("n", [1, 255, 65536, 100003])
@pytest.mark.parametrize("ltype", [torch.int64, torch.float32])
def test_bce_clipped_head(ops, n, ltype):
    """rb_bce_clipped against the oracle's Keras clipped-probability BCE and torch autograd of the same expression,
    including probabilities at and beyond the clip points."""
    rng = np.random.default_rng(n)
    p = rng.random(n).astype(np.float32)
    p[: min(n, 4)] = np.array([0.0, 1.0, 1e-9, 1.0 - 1e-9], np.float32)[: min(n, 4)]
    y = (rng.random(n) < 0.25)
    ref_loss, ref_d = O.bce_clipped(p, y.astype(np.int64))
    loss, dprob = ops.bce_clipped(cu(p), cu(y).to(ltype))
    assert abs(float(loss) - float(ref_loss)) <= 1e-6 * max(1.0, abs(float(ref_loss)))
    np.testing.assert_allclose(dprob.cpu().numpy(), ref_d, rtol=1e-5, atol=1e-12)
    pt = cu(p).requires_grad_()
    yt = cu(y).float()
    e = 1e-7
    pc = pt.clamp(e, 1 - e)
    (-(yt * torch.log(pc + e) + (1 - yt) * torch.log(1 - pc + e))).mean().backward()
    np.testing.assert_allclose(dprob.cpu().numpy(), pt.grad.cpu().numpy(), rtol=1e-5, atol=1e-12)


# ---- stand-in for compute-sanitizer (closed on this pool, profiles/r2_16_compute_sanitizer_closed.md) -------------------

CANARY = 0x5A


def _framed(shape, dtype, pad=4096):
    """A tensor of `shape` carved out of the middle of a byte buffer whose margins hold a canary pattern."""
    n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    buf = torch.full((pad + n + pad,), CANARY, dtype=torch.uint8, device="cuda")
    return buf, buf[pad:pad + n].view(dtype).reshape(shape)


def _intact(buf, n_inner, pad=4096):
    return bool((buf[:pad] == CANARY).all()) and bool((buf[pad + n_inner:] == CANARY).all())


def test_kernels_stay_inside_their_outputs_and_repeat_bit_for_bit(ops):
    """Every hot-path entry point writes into the interior of canary-framed buffers (the margins must survive) and gives
    bit-identical results when repeated on the same inputs (a race would show as run-to-run differences)."""
    from recommender_b200.ops import GradSource, LookupGroup
    g = torch.Generator(device="cuda").manual_seed(9)
    V, D, B, F = 5003, 64, 333, 26                 # odd sizes: ragged chunks, partial tiles
    W = torch.empty(V, D, device="cuda").uniform_(-0.05, 0.05, generator=g)
    idx = torch.randint(0, V, (B, F), device="cuda", generator=g)
    idx[torch.rand(B, F, device="cuda", generator=g) < 0.1] = 0
    dv = torch.randn(B, D, device="cuda", generator=g) * 0.1

    def thrice(fn):
        outs = [fn() for _ in range(3)]
        for o in outs[1:]:
            for a, b in zip(outs[0], o):
                assert torch.equal(a, b)
        return outs[0]

    # un-pooled gather into a framed [B, F, D]
    def gather():
        buf, E = _framed((B, F, D), torch.float32)
        ops.gather_fwd(W, idx, L=F, out=E, out_stride=D)
        torch.cuda.synchronize()
        assert _intact(buf, E.numel() * 4)
        return (E.clone(),)
    thrice(gather)

    # fused interaction forward (bf16 padded row with ones column) and backward
    width = 27 * 27 + D
    stride = (width + 7) // 8 * 8

    def ifwd():
        buf, out = _framed((B, stride), torch.bfloat16)
        ops.dot_interaction_fwd(table=W, idx=idx, dense_vec=dv, tail=True, out=out, out_stride=stride, out_dtype=torch.bfloat16, ones_col=True)
        torch.cuda.synchronize()
        assert _intact(buf, out.numel() * 2)
        return (out.clone(),)
    (row,) = thrice(ifwd)
    dout = (torch.randn(B, stride, device="cuda", generator=g) * 1e-2).to(torch.bfloat16)
    thrice(lambda: ops.dot_interaction_bwd(dout, table=W, idx=idx, dense_vec=dv, tail=True))

    # one-call sparse update (sort + segmented reduction + Adam) from identical state
    dE = torch.randn(B, F, D, device="cuda", generator=g) * 1e-3

    def update():
        bw, Wc = _framed((V, D), torch.float32)
        bm, m = _framed((V, D), torch.float32)
        bv, v = _framed((V, D), torch.float32)
        Wc.copy_(W)
        m.zero_()
        v.zero_()
        ops.sparse_bwd_update(Wc, m, v, [LookupGroup(idx, F, GradSource.per_position(dE, F))], optimizer="adam_lazy", step=2)
        torch.cuda.synchronize()
        assert _intact(bw, V * D * 4) and _intact(bm, V * D * 4) and _intact(bv, V * D * 4)
        return Wc.clone(), m.clone(), v.clone()
    thrice(update)

    # the three Dense products and the head, odd row counts
    rows, in_dim, units = 333, 800, 520
    x = (torch.randn(rows, in_dim, device="cuda", generator=g)).to(torch.bfloat16)
    w = (torch.randn(in_dim, units, device="cuda", generator=g) * in_dim ** -0.5).to(torch.bfloat16)
    dy = (torch.randn(rows, units, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    bias = torch.randn(units, device="cuda", generator=g)

    def dfwd():
        buf, y = _framed((rows, units), torch.bfloat16)
        ops.dense_fwd(x, w, bias, None, out=y)
        torch.cuda.synchronize()
        assert _intact(buf, y.numel() * 2)
        return (y.clone(),)
    thrice(dfwd)

    def dbwd_in():
        buf, dx = _framed((rows, in_dim), torch.bfloat16)
        ops.dense_bwd_input(dy, w, out=dx)
        torch.cuda.synchronize()
        assert _intact(buf, dx.numel() * 2)
        return (dx.clone(),)
    thrice(dbwd_in)

    def dbwd_w():
        buf, dw = _framed((in_dim, units), torch.float32)
        ops.dense_bwd_weight(x, dy, out=dw)
        torch.cuda.synchronize()
        assert _intact(buf, dw.numel() * 4)
        return (dw.clone(),)
    thrice(dbwd_w)
    wv = (torch.randn(in_dim, device="cuda", generator=g) * in_dim ** -0.5).to(torch.bfloat16)
    out = ops.dense_head_fwd(x, wv, bias[:1], "sigmoid")
    dout1 = torch.randn(rows, device="cuda", generator=g)
    thrice(lambda: tuple(t for t in ops.dense_head_bwd(dout1, out, "sigmoid", x, wv, want_dx_colsum=True)))


# ---- fused row update of the rows a step touches once (rb_dot_interaction_bwd_update + RB_APPLY_SKIP_SINGLETONS) -------------

@pytest.mark.parametrize("kind", ["adam_lazy", "adagrad", "sgd"])
@pytest.mark.parametrize("T,V,D,B,dist,dtype", [
    (26, 3000, 64, 1111, "uniform", torch.bfloat16),     # most rows touched once, ragged batch
    (26, 300, 64, 999, "uniform", torch.bfloat16),       # most rows touched several times
    (1, 20000, 64, 777, "zipf", torch.bfloat16),         # one shared table, hot rows
    (26, 1000, 16, 640, "zipf", torch.float32),
    (26, 1000, 128, 257, "uniform", torch.float32),
    (26, 50000, 32, 512, "uniform", torch.bfloat16),     # every row touched once
])
def test_fused_singleton_update_equals_the_unfused_chain(ops, kind, T, V, D, B, dist, dtype):
    """The same table and optimizer state as interaction backward -> sort -> duplicate-row sum -> optimizer row update, three
    steps in a row (the state carries over), for every row-sparse optimizer: bit for bit on every row a step touches at most
    twice (a sum of two terms has one association); rows with longer chains are summed tile by tile over the COMPACTED pair
    list, whose tile borders fall elsewhere — same terms, same order, another association of the fp32 additions."""
    from recommender_b200.ops import GradSource, LookupGroup
    rng = np.random.default_rng(B + D)
    F, rows = 26, V * T
    W0 = cu(O.init_table(rng, rows, D))
    off = torch.arange(T, device="cuda", dtype=torch.int64) * V if T > 1 else None
    width = 27 * 27 + D
    stride = (width + 7) // 8 * 8

    def state():
        if kind == "adam_lazy":
            return torch.zeros_like(W0), torch.zeros_like(W0)
        if kind == "adagrad":
            return torch.full_like(W0, 0.1), None
        return None, None

    Wa, Wb = W0.clone(), W0.clone()
    sa, sb = state(), state()
    ws_a, ws_b = ops.sparse_workspace(B * F, D, rows, "cuda"), ops.sparse_workspace(B * F, D, rows, "cuda")
    single = torch.empty(B * F, dtype=torch.uint8, device="cuda")
    hp = dict(optimizer=kind, lr=1e-2)
    n_single = 0
    exact_rows = torch.ones(rows, dtype=torch.bool, device="cuda")
    for step in (1, 2, 3):
        if dist == "uniform":
            idx = cu(rng.integers(0, V, size=(B, F)))
        else:
            idx = cu(np.minimum(rng.zipf(1.3, size=(B, F)) - 1, V - 1).astype(np.int64))
        dense = cu(rng.normal(0, 0.1, size=(B, D)).astype(np.float32))
        dout = cu(rng.normal(0, 1e-2, size=(B, stride)).astype(np.float32)).to(dtype)
        same_inputs = torch.equal(Wa, Wb)          # false once a long chain has moved a row by a different last bit
        # (a) the unfused chain
        dE_a, dd_a = ops.dot_interaction_bwd(dout, table=Wa, idx=idx, field_row_offset=off, dense_vec=dense, tail=True)
        grp = LookupGroup(idx, F, GradSource.per_position(dE_a, F), field_row_offset=off)
        sel = ops.sparse_bwd_prepare(rows, D, [grp], ws_a)
        ops.sparse_bwd_apply(Wa, sa[0], sa[1], [grp], ws_a, sel, step=step, **hp)
        # (b) fused
        grp0 = LookupGroup(idx, F, None, field_row_offset=off)
        sel = ops.sparse_bwd_prepare(rows, D, [grp0], ws_b)
        sel = ops.sparse_bwd_mark_singletons(rows, D, B * F, ws_b, sel, single)
        dE_b, dd_b = ops.dot_interaction_bwd_update(dout, table=Wb, idx=idx, single=single, state0=sb[0], state1=sb[1],
                                                    field_row_offset=off, dense_vec=dense, tail=True, step=step, **hp)
        grp = LookupGroup(idx, F, GradSource.per_position(dE_b, F), field_row_offset=off)
        ops.sparse_bwd_apply(Wb, sb[0], sb[1], [grp], ws_b, sel, step=step, skip_singletons=True, **hp)
        # the flags say what numpy says
        flat = (idx + (off[None] if off is not None else 0)).reshape(-1).cpu().numpy()
        _, inv, cnt = np.unique(flat, return_inverse=True, return_counts=True)
        np.testing.assert_array_equal(single.cpu().numpy(), (cnt[inv] == 1).astype(np.uint8))
        n_single += int(single.sum())
        keep = single.view(B, F) == 0
        if same_inputs:
            assert torch.equal(dd_a, dd_b)
            assert torch.equal(dE_a[keep], dE_b[keep])
        else:
            torch.testing.assert_close(dd_a, dd_b, rtol=1e-3, atol=1e-5)        # gradients of order 1e-2 through tables that differ by <= 2e-6
            torch.testing.assert_close(dE_a[keep], dE_b[keep], rtol=1e-3, atol=1e-5)
        short = torch.ones(rows, dtype=torch.bool, device="cuda")
        uq, c_ = np.unique(flat, return_counts=True)
        short[cu(uq[c_ >= 3])] = False
        exact_rows &= short                                   # a row stays comparable bit for bit until a long chain touches it
        for name, x, y in (("table", Wa, Wb), ("state0", sa[0], sb[0]), ("state1", sa[1], sb[1])):
            if x is None:
                continue
            if step == 1:      # later steps start from tables whose long-chain rows already differ in a last bit
                assert torch.equal(x[exact_rows], y[exact_rows]), f"step {step}: {name} differs on rows touched at most twice"
            # longer chains: |error of the sum| ~ 1e-9 here; Adam's step alpha * m / (sqrt(v) + eps) amplifies it where the
            # terms nearly cancel, bounded by a small fraction of one learning-rate step (lr = 1e-2)
            if same_inputs:
                if name == "table":
                    torch.testing.assert_close(x, y, rtol=0, atol=2e-6)
                else:
                    torch.testing.assert_close(x, y, rtol=1e-4, atol=1e-9)
            else:      # the steps' inputs already differ in last bits: the two runs drift apart like any two fp32 summation orders
                torch.testing.assert_close(x, y, rtol=5e-2, atol=2e-5 if name == "table" else 1e-6)
    assert not torch.equal(Wa, W0)
    if V >= 3000 and dist == "uniform":
        assert n_single > 0.5 * 3 * B * F
    ops.check_oob("cuda")
