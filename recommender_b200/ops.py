"""Thin Python calls over the C ABI (include/recsys_b200.h): one function per entry point.

torch tensors are only the carrier of device memory and of the CUDA stream; every computation
happens in librecsys_b200.so.  All calls are asynchronous on torch's current stream.  Nothing
here falls back to torch/CPU arithmetic: a non-CUDA tensor is an error.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import lib, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.RecsysError("recommender_b200 ops take CUDA tensors only (there is no CPU path)")


def _idx(idx: torch.Tensor):
    if idx.dtype == torch.int64:
        return _lib.RB_I64
    if idx.dtype == torch.int32:
        return _lib.RB_I32
    raise TypeError(f"indices must be int32 or int64, got {idx.dtype}")


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def version() -> int:
    return lib.rb_version()


def kernel_launches() -> int:
    """Kernels of librecsys_b200.so launched by this process so far."""
    return int(lib.rb_kernel_launches())


def adam_alpha_t(lr, beta_1, beta_2, step) -> float:
    return float(lib.rb_adam_alpha_t(lr, beta_1, beta_2, int(step)))


_oob_flags = {}


def oob_flag(device) -> torch.Tensor:
    """Per-device int32 flag the kernels raise when an index falls outside [0, rows)."""
    key = torch.device(device).index or 0
    if key not in _oob_flags:
        _oob_flags[key] = torch.zeros(1, dtype=torch.int32, device=f"cuda:{key}")
    return _oob_flags[key]


def check_oob(device) -> None:
    """Synchronising debug check (TF's CPU kernel raises InvalidArgument; the GPU one zero-fills)."""
    f = oob_flag(device)
    if int(f.item()) != 0:
        f.zero_()
        raise IndexError("embedding index out of range (InvalidArgument)")


# --------------------------------------------------------------------------------------------
# K1: lookups
# --------------------------------------------------------------------------------------------
def check_indices(idx, field_rows, *, hash_mod=0) -> None:
    """Raise the device's out-of-range flag (read by check_oob) when an id of column f lies outside [0, field_rows[f]): the
    per-table check keras.layers.Embedding makes on CPU, for T tables stored back to back (idx [..., T], field_rows int64[T])."""
    _need_cuda(idx, field_rows)
    idx = idx.contiguous()
    L = int(field_rows.numel())
    if idx.numel() % L != 0 or field_rows.dtype != torch.int64:
        raise ValueError("check_indices takes ids with T columns and int64 row counts [T]")
    check(lib.rb_check_indices(_ptr(idx), _idx(idx), idx.numel(), L, _ptr(field_rows), int(hash_mod), _ptr(oob_flag(idx.device)),
                               _stream()), "rb_check_indices")



def gather_fwd(table, idx, *, L=1, field_row_offset=None, hash_mod=0, out=None, out_stride=None):
    """out[p,:] = table[row(idx[p]),:]   (rb_gather_fwd).  idx any shape; returns [*idx.shape, D]."""
    _need_cuda(table, idx, field_row_offset, out)
    _f32c(table, "table")
    idx = idx.contiguous()
    rows, D = table.shape
    n = idx.numel()
    if out is None:
        out = torch.empty(*idx.shape, D, dtype=torch.float32, device=table.device)
        out_stride = D
    check(lib.rb_gather_fwd(_ptr(table), rows, D, _ptr(idx), _idx(idx), n, int(L), _ptr(field_row_offset), int(hash_mod),
                            _ptr(out), int(out_stride), _ptr(oob_flag(table.device)), _stream()), "rb_gather_fwd")
    return out


def bag_pool_fwd(table, idx, mode="sum", *, mask_idx=None, field_row_offset=None, hash_mod=0, out=None,
                 out_stride=None, want_count=False):
    """Pooled lookup over axis 1 of idx[B,L] (rb_bag_pool_fwd).  Returns out[B,D] (and count[B])."""
    _need_cuda(table, idx, mask_idx, field_row_offset, out)
    _f32c(table, "table")
    idx = idx.contiguous()
    B, L = idx.shape
    rows, D = table.shape
    if mask_idx is not None:
        mask_idx = mask_idx.contiguous()
        if mask_idx.dtype != idx.dtype or mask_idx.shape != idx.shape:
            raise ValueError("mask_idx must have the dtype and shape of idx")
    if out is None:
        out = torch.empty(B, D, dtype=torch.float32, device=table.device)
        out_stride = D
    count = torch.empty(B, dtype=torch.float32, device=table.device) if want_count else None
    check(lib.rb_bag_pool_fwd(_ptr(table), rows, D, _ptr(idx), _idx(idx), B, L, _ptr(field_row_offset), int(hash_mod),
                              _lib.POOL_ENUM[mode], _ptr(mask_idx), _ptr(out), int(out_stride), _ptr(count),
                              _ptr(oob_flag(table.device)), _stream()), "rb_bag_pool_fwd")
    return (out, count) if want_count else out


def gather_fm_fwd(table, idx, *, field_row_offset=None, hash_mod=0, want_E=True, want_s=True):
    """DeepFM front end (rb_gather_fm_fwd): returns (E[B,F,D] | None, s[B,D] | None, fm[B])."""
    _need_cuda(table, idx, field_row_offset)
    _f32c(table, "table")
    idx = idx.contiguous()
    B, F = idx.shape
    rows, D = table.shape
    dev = table.device
    E = torch.empty(B, F, D, dtype=torch.float32, device=dev) if want_E else None
    s = torch.empty(B, D, dtype=torch.float32, device=dev) if want_s else None
    fm = torch.empty(B, dtype=torch.float32, device=dev)
    check(lib.rb_gather_fm_fwd(_ptr(table), rows, D, _ptr(idx), _idx(idx), B, F, _ptr(field_row_offset), int(hash_mod),
                               _ptr(E), _ptr(s), _ptr(fm), _ptr(oob_flag(dev)), _stream()), "rb_gather_fm_fwd")
    return E, s, fm


def gather_fm_deep_fwd(table, idx, dense, ld_deep: int, *, field_row_offset=None, hash_mod=0):
    """DeepFM front end emitting the MLP's padded bf16 input row (rb_gather_fm_deep_fwd):
    returns (deep bf16 [B, ld_deep] = [flatten(E) | dense | 1 | 0...], s f32 [B, D], fm f32 [B])."""
    _need_cuda(table, idx, dense, field_row_offset)
    _f32c(table, "table")
    idx = idx.contiguous()
    B, F = idx.shape
    rows, D = table.shape
    dev = table.device
    if dense.dtype != torch.float32 or dense.dim() != 2 or dense.stride(1) != 1 or dense.shape[0] != B:
        raise TypeError("dense features must be float32 [B, num_dense] with unit inner stride")
    deep = torch.empty(B, ld_deep, dtype=torch.bfloat16, device=dev)
    s = torch.empty(B, D, dtype=torch.float32, device=dev)
    fm = torch.empty(B, dtype=torch.float32, device=dev)
    check(lib.rb_gather_fm_deep_fwd(_ptr(table), rows, D, _ptr(idx), _idx(idx), B, F, _ptr(field_row_offset), int(hash_mod),
                                    _ptr(dense), dense.shape[1], dense.stride(0), _ptr(deep), int(ld_deep), None, _ptr(s), _ptr(fm),
                                    _ptr(oob_flag(dev)), _stream()), "rb_gather_fm_deep_fwd")
    return deep, s, fm


# --------------------------------------------------------------------------------------------
# K3..K6, K10: DotInteraction
# --------------------------------------------------------------------------------------------

def interaction_ncols(Fp: int, self_interaction: bool, skip_gather: bool) -> int:
    if skip_gather:
        return Fp * Fp
    return Fp * (Fp + 1) // 2 if self_interaction else Fp * (Fp - 1) // 2


def _float_type(dtype) -> int:
    if dtype == torch.float32:
        return _lib.RB_F32
    if dtype == torch.bfloat16:
        return _lib.RB_BF16
    raise TypeError(f"interaction rows are float32 or bfloat16, got {dtype}")


def dot_interaction_fwd(*, E=None, table=None, idx=None, field_row_offset=None, dense_vec=None,
                        self_interaction=False, skip_gather=True, tail=False, out=None, out_stride=None,
                        out_dtype=torch.float32, pad_to=1, ones_col=False, row_cache=False, row_cache_hint=None):
    """rb_dot_interaction_fwd.  Either E[B,F,D] or (table, idx[B,F]) supplies the embedding rows.

    out_dtype=torch.float32 returns the reference layout [B, ncols(+D)].  out_dtype=torch.bfloat16
    returns [B, round_up(ncols(+D), pad_to)] in bf16 with zero pad columns (a GEMM-ready K operand); ones_col=True
    sets the FIRST pad column to 1.0 instead (the consumer's bias gradient then falls out of its dW GEMM)."""
    _need_cuda(E, table, idx, field_row_offset, dense_vec, out)
    if E is not None:
        _f32c(E, "E")
        B, F, D = E.shape
        rows, it, dev = 0, _lib.RB_I64, E.device
    else:
        _f32c(table, "table")
        idx = idx.contiguous()
        B, F = idx.shape
        rows, D = table.shape
        it, dev = _idx(idx), table.device
    if dense_vec is not None:
        _f32c(dense_vec, "dense_vec")
        if tuple(dense_vec.shape) != (B, D):
            raise ValueError(f"dense_vec must be [{B},{D}], got {tuple(dense_vec.shape)}")
    Fp = F + (dense_vec is not None)
    ncols = interaction_ncols(Fp, self_interaction, skip_gather)
    width = ncols + (D if tail else 0)
    if out is None:
        out_stride = (width + pad_to - 1) // pad_to * pad_to if out_dtype == torch.bfloat16 else width
        out = torch.empty(B, out_stride, dtype=out_dtype, device=dev)
    elif out.dtype != out_dtype:
        raise TypeError("out.dtype must equal out_dtype")
    check(lib.rb_dot_interaction_fwd(_ptr(E), _ptr(table), rows, _ptr(idx), it, _ptr(field_row_offset), _ptr(dense_vec),
                                     B, F, D, int(self_interaction), int(skip_gather), int(tail), _ptr(out),
                                     _lib.RB_BF16_ONES if (ones_col and out_dtype == torch.bfloat16 and out_stride > width)
                                     else _float_type(out_dtype), int(out_stride), _lib.ROW_CACHE_ENUM[row_cache], _ptr(row_cache_hint),
                                     _stream()), "rb_dot_interaction_fwd")
    return out


def dot_interaction_bwd(dOut, *, E=None, table=None, idx=None, field_row_offset=None, dense_vec=None,
                        self_interaction=False, skip_gather=True, tail=False, want_dE=True, row_cache=False, row_cache_hint=None):
    """rb_dot_interaction_bwd.  Returns (dE[B,F,D] | None, d_dense[B,D] | None)."""
    _need_cuda(dOut, E, table, idx, field_row_offset, dense_vec)
    if E is not None:
        _f32c(E, "E")
        B, F, D = E.shape
        rows, it, dev = 0, _lib.RB_I64, E.device
    else:
        _f32c(table, "table")
        idx = idx.contiguous()
        B, F = idx.shape
        rows, D = table.shape
        it, dev = _idx(idx), table.device
    if dOut.stride(-1) != 1 or dOut.dim() != 2 or dOut.shape[0] != B:
        raise ValueError("dOut must be [B, cols] (float32 or bfloat16) with unit inner stride")
    dE = torch.empty(B, F, D, dtype=torch.float32, device=dev) if want_dE else None
    d_dense = torch.empty(B, D, dtype=torch.float32, device=dev) if dense_vec is not None else None
    check(lib.rb_dot_interaction_bwd(_ptr(E), _ptr(table), rows, _ptr(idx), it, _ptr(field_row_offset), _ptr(dense_vec),
                                     B, F, D, int(self_interaction), int(skip_gather), int(tail), _ptr(dOut),
                                     _float_type(dOut.dtype), int(dOut.stride(0)), _ptr(dE), _ptr(d_dense), _lib.ROW_CACHE_ENUM[row_cache],
                                     _ptr(row_cache_hint), _stream()), "rb_dot_interaction_bwd")
    return dE, d_dense


def dot_interaction_bwd_update(dOut, *, table, idx, single, state0, state1, field_row_offset=None, dense_vec=None,
                               self_interaction=False, skip_gather=True, tail=False, optimizer="adam_lazy", step=1, lr=1e-3,
                               beta_1=0.9, beta_2=0.999, epsilon=1e-7, alpha_dev=None, row_cache=False, row_cache_hint=None,
                               dE=None):
    """rb_dot_interaction_bwd_update: the backward of the fused lookup + interaction that also applies the optimizer to
    the rows flagged in `single` (uint8 [B*F], from sparse_bwd_mark_singletons) — their dE rows stay unwritten.
    Returns (dE, d_dense)."""
    _need_cuda(dOut, table, idx, single, state0, state1, field_row_offset, dense_vec)
    _f32c(table, "table")
    idx = idx.contiguous()
    B, F = idx.shape
    rows, D = table.shape
    if dOut.stride(-1) != 1 or dOut.dim() != 2 or dOut.shape[0] != B:
        raise ValueError("dOut must be [B, cols] (float32 or bfloat16) with unit inner stride")
    if single.dtype != torch.uint8 or single.numel() != B * F or not single.is_contiguous():
        raise ValueError("single must be a contiguous uint8 tensor with one flag per lookup position")
    if dE is None:
        dE = torch.empty(B, F, D, dtype=torch.float32, device=table.device)
    d_dense = torch.empty(B, D, dtype=torch.float32, device=table.device) if dense_vec is not None else None
    opt = _opt_params(optimizer, step, lr, beta_1, beta_2, epsilon, alpha_dev)
    check(lib.rb_dot_interaction_bwd_update(_ptr(table), rows, _ptr(idx), _idx(idx), _ptr(field_row_offset), _ptr(dense_vec), B, F, D,
                                            int(self_interaction), int(skip_gather), int(tail), _ptr(dOut), _float_type(dOut.dtype),
                                            int(dOut.stride(0)), _ptr(dE), _ptr(d_dense), _ptr(single), _ptr(state0), _ptr(state1),
                                            C.byref(opt), _lib.ROW_CACHE_ENUM[row_cache], _ptr(row_cache_hint), _stream()),
          "rb_dot_interaction_bwd_update")
    return dE, d_dense


# --------------------------------------------------------------------------------------------
# K7..K9: backward scatter + sparse optimizer
# --------------------------------------------------------------------------------------------

class GradSource:
    """Python mirror of rb_grad_source (see the header for the addressing rule)."""

    def __init__(self, srcs: Sequence[torch.Tensor], bag_strides: Sequence[int], pos_strides: Sequence[int],
                 scale="none", mask_idx=None, count=None, fm_g=None, fm_s=None):
        if not (1 <= len(srcs) <= _lib.RB_MAX_GRAD_SOURCES):
            raise ValueError(f"1..{_lib.RB_MAX_GRAD_SOURCES} gradient sources, got {len(srcs)}")
        for s in srcs:
            if s.dtype != torch.float32:
                raise TypeError("gradient sources must be float32")
        self.srcs, self.bag_strides, self.pos_strides = list(srcs), list(bag_strides), list(pos_strides)
        self.scale, self.mask_idx, self.count, self.fm_g, self.fm_s = scale, mask_idx, count, fm_g, fm_s

    @classmethod
    def per_position(cls, dE: torch.Tensor, L: int):
        """dE[..., L, D] contiguous: one gradient row per lookup position (un-pooled lookup)."""
        _f32c(dE, "dE")
        D = dE.shape[-1]
        return cls([dE], [L * D], [D])

    @classmethod
    def per_bag(cls, douts: Sequence[torch.Tensor], scale="none", mask_idx=None, count=None, col_offset=0, D=None):
        """Bag-level gradients dout[k][B, ld_k]: every position of bag b receives
        dout[k][b, col_offset : col_offset + D] (sum / mean / masked-mean pooling, or bag size 1)."""
        return cls([d[:, col_offset:] if col_offset else d for d in douts], [d.stride(0) for d in douts],
                   [0] * len(douts), scale=scale, mask_idx=mask_idx, count=count)

    def to_c(self) -> _lib.RbGradSource:
        g = _lib.RbGradSource()
        g.num_src = len(self.srcs)
        g.scale_mode = _lib.SCALE_ENUM[self.scale]
        for k, (s, bs, ps) in enumerate(zip(self.srcs, self.bag_strides, self.pos_strides)):
            g.src[k] = s.data_ptr()
            g.bag_stride[k] = int(bs)
            g.pos_stride[k] = int(ps)
        g.mask_idx = _ptr(self.mask_idx)
        g.count = _ptr(self.count)
        g.fm_g = _ptr(self.fm_g)
        g.fm_s = _ptr(self.fm_s)
        return g


class LookupGroup:
    """One use of a table inside a step: its index array, bag length and gradient source."""

    def __init__(self, idx: torch.Tensor, L: int, grad: Optional[GradSource], field_row_offset=None, hash_mod=0):
        self.idx = idx.contiguous()
        self.L, self.grad, self.field_row_offset, self.hash_mod = int(L), grad, field_row_offset, int(hash_mod)

    @property
    def n(self) -> int:
        return self.idx.numel()


_workspaces = {}
_retired = []


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Caller-owned scratch (the library never allocates): grown geometrically, reused per device."""
    key = torch.device(device).index or 0
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _retired.append(ws)      # a captured CUDA graph may hold its address: outgrown buffers are kept, never freed
        ws = torch.empty(max(nbytes, 1 << 20) * 5 // 4, dtype=torch.uint8, device=f"cuda:{key}")
        _workspaces[key] = ws
    return ws


def _opt_params(optimizer: str, step: int, lr: float, beta_1=0.9, beta_2=0.999, epsilon=1e-7, alpha_dev=None) -> _lib.RbOptParams:
    return _lib.RbOptParams(_lib.OPTIMIZER_ENUM[optimizer], int(step), float(lr), float(beta_1), float(beta_2), float(epsilon),
                            _ptr(alpha_dev))


def _groups_c(groups, with_grad=True):
    arr = (_lib.RbLookupGroup * len(groups))()
    for k, g in enumerate(groups):
        arr[k].idx = g.idx.data_ptr()
        arr[k].idx_type = _idx(g.idx)
        arr[k].L = g.L
        arr[k].n = g.n
        arr[k].field_row_offset = _ptr(g.field_row_offset)
        arr[k].hash_mod = g.hash_mod
        if with_grad:
            arr[k].grad = g.grad.to_c()
    return arr


def sparse_workspace(n: int, D: int, rows: int, device) -> torch.Tensor:
    """A private workspace for the prepare -> apply pair (the shared per-device one may be reused in between)."""
    nbytes = lib.rb_sparse_bwd_update_workspace_bytes(n, D, rows)
    if nbytes == 0:
        raise _lib.RecsysError("rb_sparse_bwd_update_workspace_bytes rejected the problem size")
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def sparse_bwd_prepare(rows: int, D: int, groups: Sequence["LookupGroup"], ws: torch.Tensor, hot_rows_flag=None) -> int:
    """Phase 1 (rb_sparse_bwd_prepare): keys + stable radix sort of the groups' ids into `ws`, on the
    current stream.  Returns the selector to hand to sparse_bwd_apply.  Group gradients are not read."""
    for g in groups:
        _need_cuda(g.idx, g.field_row_offset)
    sel = C.c_int32(0)
    check(lib.rb_sparse_bwd_prepare(int(rows), int(D), _groups_c(groups, with_grad=False), len(groups), _ptr(ws), ws.numel(),
                                    _ptr(oob_flag(ws.device)), C.byref(sel), _ptr(hot_rows_flag), _stream()), "rb_sparse_bwd_prepare")
    return int(sel.value)


def sparse_bwd_mark_singletons(rows: int, D: int, n: int, ws: torch.Tensor, sel: int, single: torch.Tensor) -> int:
    """rb_sparse_bwd_mark_singletons: single[p] = 1 iff the row of lookup position p occurs once among the pairs sorted by
    sparse_bwd_prepare(ws) — the rows dot_interaction_bwd_update may update on its own; the other rows' pairs are compacted.
    Returns the selector to hand to sparse_bwd_apply(skip_singletons=True)."""
    _need_cuda(ws, single)
    if single.dtype != torch.uint8 or single.numel() < n:
        raise ValueError("single must be uint8 with one element per lookup position")
    s = C.c_int32(int(sel))
    check(lib.rb_sparse_bwd_mark_singletons(int(rows), int(D), int(n), _ptr(ws), ws.numel(), C.byref(s), _ptr(single), _stream()),
          "rb_sparse_bwd_mark_singletons")
    return int(s.value)


def sparse_bwd_apply(table, state0, state1, groups: Sequence["LookupGroup"], ws: torch.Tensor, sel: int, *, optimizer="adam_lazy",
                     step=1, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, alpha_dev=None, skip_singletons=False) -> None:
    """Phase 2 (rb_sparse_bwd_apply): segmented reduction + optimizer row update over pairs sorted by
    sparse_bwd_prepare with the same groups and workspace.  skip_singletons: rows that occur once were already updated by
    dot_interaction_bwd_update (rb_sparse_bwd_apply_ex, RB_APPLY_SKIP_SINGLETONS)."""
    _need_cuda(table, state0, state1)
    for g in groups:
        _need_cuda(g.idx, g.field_row_offset, *g.grad.srcs, g.grad.mask_idx, g.grad.count, g.grad.fm_g, g.grad.fm_s)
    _f32c(table, "table")
    rows, D = table.shape
    opt = _opt_params(optimizer, step, lr, beta_1, beta_2, epsilon, alpha_dev)
    if skip_singletons:
        check(lib.rb_sparse_bwd_apply_ex(_ptr(table), _ptr(state0), _ptr(state1), rows, D, _groups_c(groups), len(groups), C.byref(opt),
                                         _ptr(ws), ws.numel(), int(sel), 1, _stream()), "rb_sparse_bwd_apply_ex")
        return
    check(lib.rb_sparse_bwd_apply(_ptr(table), _ptr(state0), _ptr(state1), rows, D, _groups_c(groups), len(groups), C.byref(opt),
                                  _ptr(ws), ws.numel(), int(sel), _stream()), "rb_sparse_bwd_apply")


def sparse_bwd_update(table, state0, state1, groups: Sequence[LookupGroup], *, optimizer="adam_lazy", step=1,
                      lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, alpha_dev=None) -> None:
    """IndexedSlices -> duplicate-row sum -> optimizer row update, in place on table/state
    (rb_sparse_bwd_update_groups; one group = rb_sparse_bwd_update)."""
    _need_cuda(table, state0, state1)
    for g in groups:
        _need_cuda(g.idx, g.field_row_offset, *g.grad.srcs, g.grad.mask_idx, g.grad.count, g.grad.fm_g, g.grad.fm_s)
    _f32c(table, "table")
    rows, D = table.shape
    if not (1 <= len(groups) <= _lib.RB_MAX_LOOKUP_GROUPS):
        raise ValueError(f"1..{_lib.RB_MAX_LOOKUP_GROUPS} lookup groups per table and step, got {len(groups)}")
    n = sum(g.n for g in groups)
    nbytes = lib.rb_sparse_bwd_update_workspace_bytes(n, D, rows)
    if nbytes == 0:
        raise _lib.RecsysError("rb_sparse_bwd_update_workspace_bytes rejected the problem size")
    ws = _workspace(nbytes, table.device)
    opt = _opt_params(optimizer, step, lr, beta_1, beta_2, epsilon, alpha_dev)
    arr = _groups_c(groups)
    check(lib.rb_sparse_bwd_update_groups(_ptr(table), _ptr(state0), _ptr(state1), rows, D, arr, len(groups),
                                          C.byref(opt), _ptr(ws), ws.numel(), _ptr(oob_flag(table.device)), _stream()),
          "rb_sparse_bwd_update_groups")


def sparse_bwd_dedup(rows: int, D: int, idx, L: int, grad: GradSource, *, field_row_offset=None, hash_mod=0):
    """The deduplicated IndexedSlices itself (rb_sparse_bwd_dedup): (unique rows ascending, summed grads)."""
    _need_cuda(idx, field_row_offset)
    idx = idx.contiguous()
    n = idx.numel()
    dev = idx.device
    uniq_rows = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    uniq_grad = torch.empty(max(n, 1), D, dtype=torch.float32, device=dev)
    num = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = _workspace(max(lib.rb_sparse_bwd_update_workspace_bytes(n, D, rows), 256), dev)
    g = grad.to_c()
    check(lib.rb_sparse_bwd_dedup(rows, D, _ptr(idx), _idx(idx), n, int(L), _ptr(field_row_offset), int(hash_mod),
                                  C.byref(g), _ptr(uniq_rows), _ptr(uniq_grad), _ptr(num), _ptr(ws), ws.numel(),
                                  _ptr(oob_flag(dev)), _stream()), "rb_sparse_bwd_dedup")
    u = int(num.item())
    return uniq_rows[:u], uniq_grad[:u]


# --------------------------------------------------------------------------------------------
# dense side of the step
# --------------------------------------------------------------------------------------------

def dense_opt_step(params, grads, state0, state1, *, optimizer="adam_lazy", step=1, lr=1e-3, beta_1=0.9, beta_2=0.999,
                   epsilon=1e-7, alpha_dev=None) -> None:
    """Keras `_resource_apply_dense` for a list of fp32 tensors, RB_MAX_DENSE_TENSORS per launch
    (rb_dense_opt_step).  state0/state1: lists (Adam m, v; Adagrad acc, None) or None (SGD)."""
    opt = _opt_params(optimizer, step, lr, beta_1, beta_2, epsilon, alpha_dev)
    n = len(params)
    for lo in range(0, n, _lib.RB_MAX_DENSE_TENSORS):
        hi = min(n, lo + _lib.RB_MAX_DENSE_TENSORS)
        arr = (_lib.RbDenseSlot * (hi - lo))()
        for k in range(lo, hi):
            p, g = params[k], grads[k]
            _need_cuda(p, g)
            if p.dtype != torch.float32 or g.dtype != torch.float32 or not p.is_contiguous() or not g.is_contiguous():
                raise TypeError("dense_opt_step takes contiguous float32 parameters and gradients")
            arr[k - lo].param = p.data_ptr()
            arr[k - lo].grad = g.data_ptr()
            arr[k - lo].state0 = None if state0 is None else state0[k].data_ptr()
            arr[k - lo].state1 = None if state1 is None else state1[k].data_ptr()
            arr[k - lo].n = p.numel()
            sh = getattr(p, "_rb_shadow", None)      # (bf16 tensor whose first n elements mirror p, version it was made at)
            arr[k - lo].shadow_bf16 = sh[0].data_ptr() if sh is not None and sh[1] == p._version else None
        check(lib.rb_dense_opt_step(arr, hi - lo, C.byref(opt), _stream()), "rb_dense_opt_step")


def mm_f32_out(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a @ b with fp32 accumulation AND an fp32 result from bf16 operands: the weight-gradient GEMM of the cuBLASLt A/B path
    (layers._LinearBF16Fn; the default path is rb_dense_bwd_weight)."""
    _need_cuda(a, b)
    return torch.mm(a, b, out_dtype=torch.float32)


def colsum(x: torch.Tensor, out=None, ws=None) -> torch.Tensor:
    """fp32 column sums of a [rows, cols] float32 / bfloat16 matrix (rb_colsum): a Dense bias gradient."""
    _need_cuda(x)
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("colsum takes a [rows, cols] matrix with unit inner stride")
    rows, cols = x.shape
    if out is None:
        out = torch.empty(cols, dtype=torch.float32, device=x.device)
    if ws is None:
        ws = _workspace(max(lib.rb_colsum_workspace_bytes(rows, cols), 256), x.device)
    check(lib.rb_colsum(_ptr(x), _float_type(x.dtype), rows, cols, int(x.stride(0)), _ptr(out), _ptr(ws), ws.numel(), _stream()),
          "rb_colsum")
    return out


def bce_clipped(prob: torch.Tensor, label: torch.Tensor, want_grad: bool = True):
    """Keras binary_crossentropy on probabilities, batch mean (rb_bce_clipped).  Returns (loss f32[], dloss/dprob f32[n] | None)."""
    _need_cuda(prob, label)
    prob = prob.contiguous()
    label = label.contiguous()
    if prob.dtype != torch.float32 or label.dtype not in (torch.float32, torch.int64) or label.numel() != prob.numel():
        raise TypeError("bce_clipped takes float32 probabilities and float32 / int64 labels of the same size")
    n = prob.numel()
    loss = torch.empty((), dtype=torch.float32, device=prob.device)
    dprob = torch.empty_like(prob) if want_grad else None
    ws = _workspace(max(lib.rb_bce_workspace_bytes(n), 256), prob.device)
    check(lib.rb_bce_clipped(_ptr(prob), _ptr(label), 1 if label.dtype == torch.int64 else 0, n, _ptr(loss), _ptr(dprob), _ptr(ws),
                             ws.numel(), _stream()), "rb_bce_clipped")
    return loss, dprob


# --------------------------------------------------------------------------------------------
# Dense layers of the towers on tcgen05 tensor cores (csrc/mlp.cu; ctr/layers.py:5-14)
# --------------------------------------------------------------------------------------------

def _bf16m(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
        raise TypeError(f"{name} must be a bfloat16 matrix with unit inner stride, got {t.dtype} {tuple(t.shape)} {t.stride()}")
    return t


def dense_fwd(x, w, bias=None, activation=None, out_dtype=torch.bfloat16, out=None):
    """y = activation(x @ w + bias) (rb_dense_fwd).  x bf16 [rows, in_dim], w bf16 [in_dim, units], bias f32 [units];
    y bf16 or f32 [rows, units]."""
    _need_cuda(x, w, bias, out)
    _bf16m(x, "x"), _bf16m(w, "w")
    rows, in_dim = x.shape
    units = w.shape[1]
    if w.shape[0] != in_dim:
        raise ValueError(f"x has {in_dim} features, w has {w.shape[0]} rows")
    if out is None:
        out = torch.empty(rows, units, dtype=out_dtype, device=x.device)
    check(lib.rb_dense_fwd(_ptr(x), rows, in_dim, x.stride(0), _ptr(w), units, w.stride(0), _ptr(bias), _lib.ACT_ENUM[activation],
                           _ptr(out), _float_type(out.dtype), out.stride(0), _stream()), "rb_dense_fwd")
    return out


def dense_bwd_input(dy, w, out=None):
    """dx = dy @ w.T (rb_dense_bwd_input).  dy bf16 [rows, units], w bf16 [in_dim, units] -> dx bf16 [rows, in_dim]."""
    _need_cuda(dy, w, out)
    _bf16m(dy, "dy"), _bf16m(w, "w")
    rows, units = dy.shape
    in_dim = w.shape[0]
    if out is None:
        out = torch.empty(rows, in_dim, dtype=torch.bfloat16, device=dy.device)
    check(lib.rb_dense_bwd_input(_ptr(dy), rows, units, dy.stride(0), _ptr(w), in_dim, w.stride(0), _ptr(out), out.stride(0), _stream()),
          "rb_dense_bwd_input")
    return out


def dense_bwd_weight_workspace_bytes(rows: int, in_dim: int, units: int) -> int:
    nbytes = lib.rb_dense_bwd_weight_workspace_bytes(rows, in_dim, units)
    if nbytes == 0:
        raise _lib.RecsysError(f"rb_dense_bwd_weight rejected the problem size ({rows}, {in_dim}, {units})")
    return int(nbytes)


def dense_bwd_weight(x, dy, out=None, ws=None):
    """dW = x.T @ dy in fp32 (rb_dense_bwd_weight), deterministic split over the batch.  x bf16 [rows, in_dim], dy bf16 [rows, units].
    `ws`: a private workspace (a call on a stream other than the one the shared per-device workspace serves)."""
    _need_cuda(x, dy, out)
    _bf16m(x, "x"), _bf16m(dy, "dy")
    rows, in_dim = x.shape
    units = dy.shape[1]
    if out is None:
        out = torch.empty(in_dim, units, dtype=torch.float32, device=x.device)
    if ws is None:
        ws = _workspace(dense_bwd_weight_workspace_bytes(rows, in_dim, units), x.device)
    check(lib.rb_dense_bwd_weight(_ptr(x), rows, in_dim, x.stride(0), _ptr(dy), units, dy.stride(0), _ptr(out), out.stride(0), _ptr(ws),
                                  ws.numel(), _stream()), "rb_dense_bwd_weight")
    return out


def dense_head_fwd(x, w, bias=None, activation=None):
    """out[r] = activation(x[r] . w + bias) for a Dense(1) layer (rb_dense_head_fwd).  x bf16 [rows, in_dim], w bf16 [in_dim]."""
    _need_cuda(x, w, bias)
    _bf16m(x, "x")
    rows, in_dim = x.shape
    out = torch.empty(rows, dtype=torch.float32, device=x.device)
    check(lib.rb_dense_head_fwd(_ptr(x), rows, in_dim, x.stride(0), _ptr(w), _ptr(bias), _lib.ACT_ENUM[activation], _ptr(out), _stream()),
          "rb_dense_head_fwd")
    return out


def dense_head_bwd(dout, out, activation, x, w, want_dx=True, want_dx_colsum=False):
    """Backward of the Dense(1) head (rb_dense_head_bwd): returns (dx bf16 [rows, in_dim] | None, dW f32 [in_dim], db f32 [1])
    and, with want_dx_colsum, the f32 column sums of dx (the bias gradient of the layer below) as a fourth value."""
    _need_cuda(dout, out, x, w)
    _bf16m(x, "x")
    rows, in_dim = x.shape
    dx = torch.empty(rows, in_dim, dtype=torch.bfloat16, device=x.device) if want_dx else None
    dw = torch.empty(in_dim, dtype=torch.float32, device=x.device)
    db = torch.empty(1, dtype=torch.float32, device=x.device)
    cs = torch.empty(in_dim, dtype=torch.float32, device=x.device) if want_dx_colsum else None
    ws = _workspace(max(lib.rb_dense_head_bwd_workspace_bytes(rows, in_dim), 256), x.device)
    check(lib.rb_dense_head_bwd(_ptr(_f32c(dout, "dout")), _ptr(out), _lib.ACT_ENUM[activation], _ptr(x), rows, in_dim, x.stride(0), _ptr(w),
                                _ptr(dx), in_dim, _ptr(dw), _ptr(db), _ptr(cs), _ptr(ws), ws.numel(), _stream()), "rb_dense_head_bwd")
    return (dx, dw, db, cs) if want_dx_colsum else (dx, dw, db)


def dense_head_bce(x, w, bias, label, want_dx=True, want_dx_colsum=False):
    """Dense(1) + sigmoid, its clipped binary cross-entropy against `label` and the head's backward (seeded with d loss = 1) in one
    pass over x (rb_dense_head_bce).  Returns (prob f32[rows], loss f32[1], dx bf16 | None, dW f32[in_dim], db f32[1], dx_colsum | None)."""
    _need_cuda(x, w, bias, label)
    _bf16m(x, "x")
    rows, in_dim = x.shape
    if label.dtype not in (torch.float32, torch.int64) or label.numel() != rows or not label.is_contiguous():
        raise TypeError("label must be a contiguous float32 / int64 vector with one entry per row")
    dev = x.device
    prob = torch.empty(rows, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    dx = torch.empty(rows, in_dim, dtype=torch.bfloat16, device=dev) if want_dx else None
    dw = torch.empty(in_dim, dtype=torch.float32, device=dev)
    db = torch.empty(1, dtype=torch.float32, device=dev)
    cs = torch.empty(in_dim, dtype=torch.float32, device=dev) if want_dx_colsum else None
    ws = _workspace(max(lib.rb_dense_head_bce_workspace_bytes(rows, in_dim), 256), dev)
    check(lib.rb_dense_head_bce(_ptr(x), rows, in_dim, x.stride(0), _ptr(w), _ptr(bias), _ptr(label), 1 if label.dtype == torch.int64 else 0,
                                _ptr(prob), _ptr(loss), _ptr(dx), in_dim, _ptr(dw), _ptr(db), _ptr(cs), _ptr(ws), ws.numel(), _stream()),
          "rb_dense_head_bce")
    return prob, loss, dx, dw, db, cs


def dense_head_bce_ok(in_dim: int) -> bool:
    return in_dim in (8, 16, 32, 64, 128, 256)


def dense_act_bwd(dy, y, activation):
    """bf16(dy * activation'(y)) (rb_dense_act_bwd); dy, y f32 of the same contiguous shape."""
    _need_cuda(dy, y)
    _f32c(dy, "dy")
    if y is not None:
        _f32c(y, "y")
    out = torch.empty(dy.shape, dtype=torch.bfloat16, device=dy.device)
    check(lib.rb_dense_act_bwd(_ptr(dy), _ptr(y), _lib.ACT_ENUM[activation], dy.numel(), _ptr(out), _stream()), "rb_dense_act_bwd")
    return out


def dense_act_bwd_bf16(dy, y, activation):
    """bf16(dy * activation'(y)) for a hidden layer kept in bf16 (rb_dense_act_bwd_bf16); dy, y bf16 of the same contiguous shape."""
    _need_cuda(dy, y)
    if dy.dtype != torch.bfloat16 or y.dtype != torch.bfloat16 or not dy.is_contiguous() or not y.is_contiguous() or dy.shape != y.shape:
        raise TypeError("act_bwd_bf16 takes two contiguous bfloat16 tensors of one shape")
    out = torch.empty_like(dy)
    check(lib.rb_dense_act_bwd_bf16(_ptr(dy), _ptr(y), _lib.ACT_ENUM[activation], dy.numel(), _ptr(out), _stream()), "rb_dense_act_bwd_bf16")
    return out


# --------------------------------------------------------------------------------------------
# DIN attention pooling (csrc/din.cu)
# --------------------------------------------------------------------------------------------

class DinHistory:
    """Python mirror of rb_din_history: the behaviour history as ids into an item table and (optionally) a category
    table, whose rows concatenate to the E-wide history row (dien/model.py:14-19); mask = `mask` != 0, default item id != 0."""

    def __init__(self, table0, idx0, table1=None, idx1=None, mask=None):
        _need_cuda(table0, idx0, table1, idx1, mask)
        _f32c(table0, "table0")
        self.idx0 = idx0.contiguous()
        if self.idx0.dim() != 2:
            raise ValueError("history ids must be [B, L]")
        self.table0, self.table1 = table0, table1
        self.idx1 = None if idx1 is None else idx1.contiguous()
        if (table1 is None) != (idx1 is None):
            raise ValueError("give the second table together with its ids")
        if table1 is not None:
            _f32c(table1, "table1")
            if self.idx1.shape != self.idx0.shape or self.idx1.dtype != self.idx0.dtype:
                raise ValueError("both id arrays must share shape and dtype")
        self.mask = None if mask is None else mask.contiguous()
        if self.mask is not None and (self.mask.shape != self.idx0.shape or self.mask.dtype not in (torch.int32, torch.int64)):
            raise ValueError("mask must be an int32 / int64 array of the ids' shape (nonzero = valid)")
        self.B, self.L = self.idx0.shape
        self.E = table0.shape[1] + (0 if table1 is None else table1.shape[1])

    def to_c(self) -> _lib.RbDinHistory:
        h = _lib.RbDinHistory()
        h.table0, h.rows0, h.D0, h.idx0 = self.table0.data_ptr(), self.table0.shape[0], self.table0.shape[1], self.idx0.data_ptr()
        if self.table1 is not None:
            h.table1, h.rows1, h.D1, h.idx1 = self.table1.data_ptr(), self.table1.shape[0], self.table1.shape[1], self.idx1.data_ptr()
        else:
            h.table1, h.rows1, h.D1, h.idx1 = None, 0, 0, None
        h.idx_type = _idx(self.idx0)
        h.mask = _ptr(self.mask)
        h.mask_type = _idx(self.mask) if self.mask is not None else _lib.RB_I64
        h.B, h.L = self.B, self.L
        return h


def din_offsets(h: DinHistory) -> torch.Tensor:
    """int32 [B + 1]: exclusive scan of the per-sample valid counts (rb_din_offsets); offsets[B] = number of feature rows."""
    off = torch.empty(h.B + 1, dtype=torch.int32, device=h.idx0.device)
    ws = _workspace(max(lib.rb_din_workspace_bytes(max(h.B, 1)), 256), h.idx0.device)
    c = h.to_c()
    check(lib.rb_din_offsets(C.byref(c), _ptr(off), _ptr(ws), ws.numel(), _stream()), "rb_din_offsets")
    return off


def din_build_features(h: DinHistory, target, offsets, P: int, ldx: int) -> torch.Tensor:
    """bf16 [max(P, 1), ldx]: [t | h | t - h | t * h | 0...] per valid position (rb_din_build_features)."""
    _need_cuda(target, offsets)
    _f32c(target, "target")
    x = torch.empty(max(P, 1), ldx, dtype=torch.bfloat16, device=target.device)
    c = h.to_c()
    check(lib.rb_din_build_features(C.byref(c), _ptr(target), _ptr(offsets), _ptr(x), ldx, _stream()), "rb_din_build_features")
    return x


def din_pool_fwd(h: DinHistory, offsets, w) -> torch.Tensor:
    """rep f32 [B, E] = sum_p w[p] * h_p (rb_din_pool_fwd)."""
    _need_cuda(offsets, w)
    rep = torch.empty(h.B, h.E, dtype=torch.float32, device=w.device)
    c = h.to_c()
    check(lib.rb_din_pool_fwd(C.byref(c), _ptr(offsets), _ptr(_f32c(w, "w")), _ptr(rep), _stream()), "rb_din_pool_fwd")
    return rep


def din_pool_bwd_weights(h: DinHistory, offsets, d_rep, P: int) -> torch.Tensor:
    """dw f32 [max(P, 1)] = <d_rep[b], h_p> (rb_din_pool_bwd_weights)."""
    _need_cuda(offsets, d_rep)
    dw = torch.empty(max(P, 1), dtype=torch.float32, device=d_rep.device)
    c = h.to_c()
    check(lib.rb_din_pool_bwd_weights(C.byref(c), _ptr(offsets), _ptr(_f32c(d_rep, "d_rep")), _ptr(dw), _stream()), "rb_din_pool_bwd_weights")
    return dw


def din_feature_bwd(h: DinHistory, target, offsets, dx, w, d_rep, zero_masked: bool = False):
    """(dh f32 [B, L, E] — rows of valid positions; masked rows zero-filled only when zero_masked — , d_target f32 [B, E])
    (rb_din_feature_bwd)."""
    _need_cuda(target, offsets, dx, w, d_rep)
    if dx.dtype != torch.bfloat16 or dx.dim() != 2 or dx.stride(1) != 1:
        raise TypeError("dx must be a bfloat16 matrix with unit inner stride")
    dh = torch.empty(h.B, h.L, h.E, dtype=torch.float32, device=dx.device)
    dt = torch.empty(h.B, h.E, dtype=torch.float32, device=dx.device)
    c = h.to_c()
    check(lib.rb_din_feature_bwd(C.byref(c), _ptr(_f32c(target, "target")), _ptr(offsets), _ptr(dx), dx.stride(0), _ptr(_f32c(w, "w")),
                                 _ptr(_f32c(d_rep, "d_rep")), _ptr(dh), _ptr(dt), int(bool(zero_masked)), _stream()), "rb_din_feature_bwd")
    return dh, dt


def dense_pack_input(x, ld: int, ones_col: bool = True):
    """bf16 [rows, ld] = [x | 1 | 0...] from f32 x [rows, in_dim] (rb_dense_pack_input)."""
    _need_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise TypeError("pack_input takes a float32 matrix with unit inner stride")
    rows, in_dim = x.shape
    out = torch.empty(rows, ld, dtype=torch.bfloat16, device=x.device)
    check(lib.rb_dense_pack_input(_ptr(x), rows, in_dim, x.stride(0), _ptr(out), ld, int(bool(ones_col)), _stream()), "rb_dense_pack_input")
    return out


# --------------------------------------------------------------------------------------------
# id -> row map, sharding
# --------------------------------------------------------------------------------------------

def hash_ids(ids, vocab: int, world: int = 1):
    """rows = uint64(id) mod vocab; owner = row mod world; local = row div world (rb_hash_ids)."""
    _need_cuda(ids)
    ids = ids.contiguous()
    n = ids.numel()
    rows = torch.empty(ids.shape, dtype=torch.int64, device=ids.device)
    owner = torch.empty(ids.shape, dtype=torch.int32, device=ids.device)
    local = torch.empty(ids.shape, dtype=torch.int64, device=ids.device)
    check(lib.rb_hash_ids(_ptr(ids), _idx(ids), n, int(vocab), int(world), _ptr(rows), _ptr(owner), _ptr(local), _stream()),
          "rb_hash_ids")
    return rows, owner, local


def bucket_by_owner(idx, world: int, *, L=1, field_row_offset=None, hash_mod=0, skip_from_row=None):
    """Stable partition of the lookups by owner rank = row mod world (rb_bucket_by_owner / rb_bucket_by_owner_skip).

    Returns (local_rows int64[n] in bucket order, perm int32[n]: bucket slot -> lookup position,
    inv_perm int32[n]: lookup position -> bucket slot, counts int64[world]).  skip_from_row: lookups whose row is >= it belong
    to no owner (rows of replicated tables): not counted, no slot, inv_perm = -1; only the first sum(counts) slots are written."""
    _need_cuda(idx, field_row_offset)
    idx = idx.contiguous()
    n = idx.numel()
    dev = idx.device
    local_rows = torch.empty(n, dtype=torch.int64, device=dev)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    inv_perm = torch.empty(n, dtype=torch.int32, device=dev)
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    ws = _workspace(max(lib.rb_bucket_by_owner_workspace_bytes(n, world), 256), dev)
    if skip_from_row is not None:
        check(lib.rb_bucket_by_owner_skip(_ptr(idx), _idx(idx), n, int(L), _ptr(field_row_offset), int(hash_mod), int(world), int(skip_from_row),
                                          _ptr(local_rows), _ptr(perm), _ptr(inv_perm), _ptr(counts), _ptr(ws), ws.numel(), _stream()),
              "rb_bucket_by_owner_skip")
        return local_rows, perm, inv_perm, counts
    check(lib.rb_bucket_by_owner(_ptr(idx), _idx(idx), n, int(L), _ptr(field_row_offset), int(hash_mod), int(world),
                                 _ptr(local_rows), _ptr(perm), _ptr(inv_perm), _ptr(counts), _ptr(ws), ws.numel(), _stream()),
          "rb_bucket_by_owner")
    return local_rows, perm, inv_perm, counts
