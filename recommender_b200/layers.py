"""Host-side mirror of the reference's layer call surface (ctr/layers.py, keras.layers.Embedding).

Same class names, constructor arguments and call semantics as the reference so that the model
code in model.py reads like ctr/model.py; every hot-path computation is a call into
librecsys_b200.so (ops.py).  torch supplies device memory, streams and the autograd tape only.

Training contract.  An `Embedding` owns its table as a plain CUDA tensor (not an autograd
leaf).  The backward of every lookup records a *lookup group* (index array + where its gradient
rows live) on the layer; `optimizers.*.apply_gradients()` then runs ONE fused
sort -> duplicate-row sum -> optimizer-row-update call per table over the concatenation of the
step's groups — the IndexedSlices concatenation + `_deduplicate_indexed_slices` +
`_resource_apply_sparse` chain of Keras, never materialising a dense
[V, D] gradient.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import ops
from .ops import GradSource, LookupGroup


def _stream_priority(env: str, default: int = 0) -> int:
    """CUDA stream priority of a helper stream (0 = default, negative = higher), overridable for scheduling experiments."""
    import os
    v = os.environ.get(env)
    return default if v is None else int(v)


def _default_device(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("recommender_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


# ----------------------------------------------------------------------------------------------
# autograd glue: forward = one C-ABI call; backward = one C-ABI call or a recorded lookup group
# ----------------------------------------------------------------------------------------------

class _GatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, emb, idx):
        L = idx.shape[-1] if idx.dim() >= 1 else 1
        out = ops.gather_fwd(emb.embeddings, idx, L=L, field_row_offset=emb.row_offset_for(L), hash_mod=emb.hash_mod)
        ctx.emb, ctx.idx, ctx.L = emb, idx, L
        return out

    @staticmethod
    def backward(ctx, dE):
        emb = ctx.emb
        dE = dE.contiguous()
        emb._record(LookupGroup(ctx.idx, ctx.L, GradSource.per_position(dE, ctx.L),
                                field_row_offset=emb.row_offset_for(ctx.L), hash_mod=emb.hash_mod))
        return None, None, None


class _BagPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, emb, idx, mode, mask_idx):
        out, count = ops.bag_pool_fwd(emb.embeddings, idx, mode, mask_idx=mask_idx, hash_mod=emb.hash_mod, want_count=True)
        ctx.emb, ctx.idx, ctx.mode, ctx.mask_idx, ctx.count = emb, idx, mode, mask_idx, count
        return out

    @staticmethod
    def backward(ctx, dout):
        emb = ctx.emb
        dout = dout.contiguous()
        scale = {"sum": "none", "mean": "mean", "masked_mean": "masked_mean"}[ctx.mode]
        grad = GradSource.per_bag([dout], scale=scale, mask_idx=ctx.mask_idx if ctx.mask_idx is not None else ctx.idx,
                                  count=ctx.count)
        emb._record(LookupGroup(ctx.idx, ctx.idx.shape[1], grad, hash_mod=emb.hash_mod))
        return None, None, None, None, None


class _GatherFMFn(torch.autograd.Function):
    """ctr/model.py:19-23 in one pass: E, and fm = 0.5*sum_d((sum_f E)^2 - sum_f E^2)."""

    @staticmethod
    def forward(ctx, anchor, emb, idx):
        F = idx.shape[1]
        E, s, fm = ops.gather_fm_fwd(emb.embeddings, idx, field_row_offset=emb.row_offset_for(F), hash_mod=emb.hash_mod)
        ctx.emb, ctx.idx, ctx.s = emb, idx, s
        return E, fm

    @staticmethod
    def backward(ctx, dE, dfm):
        emb = ctx.emb
        B, F = ctx.idx.shape
        D = emb.output_dim
        if dE is None:
            src, bs, ps = torch.zeros(D, dtype=torch.float32, device=ctx.idx.device), 0, 0
        else:
            src, bs, ps = dE.contiguous(), F * D, D
        grad = GradSource([src], [bs], [ps], fm_g=None if dfm is None else dfm.contiguous(), fm_s=ctx.s)
        emb._record(LookupGroup(ctx.idx, F, grad, field_row_offset=emb.row_offset_for(F), hash_mod=emb.hash_mod))
        return None, None, None


class _GatherFMDeepFn(torch.autograd.Function):
    """ctr/model.py:19-26 in one pass: fm, and the deep input [flatten(E) | int_features | 1 | 0...] as the bf16 K operand
    of the MLP's first layer; E itself is never written.  Backward: the MLP's dx (bf16) carries the gradient of E through
    the reshape / concat; the FM term is added inside the scatter (fm_g, fm_s)."""

    @staticmethod
    def forward(ctx, anchor, emb, idx, dense, ld):
        F = idx.shape[1]
        deep, s, fm = ops.gather_fm_deep_fwd(emb.embeddings, idx, dense, ld, field_row_offset=emb.row_offset_for(F), hash_mod=emb.hash_mod)
        ctx.emb, ctx.idx, ctx.s = emb, idx, s
        ctx.mark_non_differentiable(s)
        return deep, fm

    @staticmethod
    def backward(ctx, d_deep, dfm):
        emb = ctx.emb
        B, F = ctx.idx.shape
        D = emb.output_dim
        if d_deep is None:
            src, bs, ps = torch.zeros(D, dtype=torch.float32, device=ctx.idx.device), 0, 0
        else:
            src, bs, ps = d_deep[:, : F * D].float().contiguous(), F * D, D
        grad = GradSource([src], [bs], [ps], fm_g=None if dfm is None else dfm.contiguous(), fm_s=ctx.s)
        emb._record(LookupGroup(ctx.idx, F, grad, field_row_offset=emb.row_offset_for(F), hash_mod=emb.hash_mod))
        return None, None, None, None, None


class _InteractFn(torch.autograd.Function):
    """Lookup + DLRM concat + DotInteraction + '|| bmlp' tail in one kernel (ctr/model.py:49-55)."""

    @staticmethod
    def forward(ctx, anchor, emb, idx, dense_vec, self_interaction, skip_gather, tail, out_dtype, pad_to, ones_col=False):
        F = idx.shape[1]
        dense_vec = dense_vec.contiguous()
        out = ops.dot_interaction_fwd(table=emb.embeddings, idx=idx, field_row_offset=emb.row_offset_for(F),
                                      dense_vec=dense_vec, self_interaction=self_interaction, skip_gather=skip_gather,
                                      tail=tail, out_dtype=out_dtype, pad_to=pad_to, ones_col=ones_col,
                                      row_cache=emb.row_cache, row_cache_hint=emb._hot_rows)
        ctx.emb, ctx.idx, ctx.flags = emb, idx, (self_interaction, skip_gather, tail)
        ctx.save_for_backward(dense_vec)
        return out

    @staticmethod
    def backward(ctx, dOut):
        (dense_vec,) = ctx.saved_tensors
        emb, idx = ctx.emb, ctx.idx
        F = idx.shape[1]
        si, sg, tail = ctx.flags
        if dOut.stride(-1) != 1:
            dOut = dOut.contiguous()
        fused = emb._fused_update_args(idx)
        if fused is not None:
            # rows this step touches once get their optimizer update right here (rb_dot_interaction_bwd_update); the sorted
            # reduction launched by apply_pending() then passes over them
            dE, d_dense = ops.dot_interaction_bwd_update(dOut, table=emb.embeddings, idx=idx, field_row_offset=emb.row_offset_for(F),
                                                         dense_vec=dense_vec, self_interaction=si, skip_gather=sg, tail=tail,
                                                         row_cache=emb.row_cache, row_cache_hint=emb._hot_rows, **fused)
            emb._fused_done = True
        else:
            dE, d_dense = ops.dot_interaction_bwd(dOut, table=emb.embeddings, idx=idx, field_row_offset=emb.row_offset_for(F),
                                                  dense_vec=dense_vec, self_interaction=si, skip_gather=sg, tail=tail,
                                                  row_cache=emb.row_cache, row_cache_hint=emb._hot_rows)
        emb._record(LookupGroup(idx, F, GradSource.per_position(dE, F), field_row_offset=emb.row_offset_for(F),
                                hash_mod=emb.hash_mod))
        emb._grad_ready = torch.cuda.current_stream().record_event()    # dE is complete here: the row update may start
        for hook in emb.after_grad_hooks:
            hook(emb._grad_ready)
        return None, None, None, d_dense, None, None, None, None, None, None


class _DotInteractionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, self_interaction, skip_gather):
        X = X.contiguous()
        ctx.save_for_backward(X)
        ctx.flags = (self_interaction, skip_gather)
        return ops.dot_interaction_fwd(E=X, self_interaction=self_interaction, skip_gather=skip_gather)

    @staticmethod
    def backward(ctx, dOut):
        (X,) = ctx.saved_tensors
        si, sg = ctx.flags
        if dOut.stride(-1) != 1:
            dOut = dOut.contiguous()
        dX, _ = ops.dot_interaction_bwd(dOut, E=X, self_interaction=si, skip_gather=sg)
        return dX, None, None


# ----------------------------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------------------------

class Embedding(nn.Module):
    """Stand-in for keras.layers.Embedding(input_dim, output_dim, mask_zero=False)
    (ctr/model.py:10,42; dien/model.py:11-12; esmm/esmm.py:10-11).

    `embeddings` f32[input_dim, output_dim] ~ U(-0.05, 0.05) (the Keras default initialiser).
    `num_tables` > 1 stores that many input_dim-row tables back to back and routes position f of
    a [B, num_tables] index array to table f (the T = 26 form of BASELINE config 2).  `table_rows`
    gives every table its own row count instead (BASELINE config 3: capped Criteo-Terabyte
    cardinalities from 3 to 40M rows); `cap_ids(ids)` folds raw ids into each table's range.
    `hash_mod` folds ids with uint64 mod before the lookup (capped tables).
    """

    def __init__(self, input_dim: int, output_dim: int, mask_zero: bool = False, *, num_tables: int = 1,
                 hash_mod: int = 0, table_rows: Optional[Sequence[int]] = None, device=None,
                 generator: Optional[torch.Generator] = None):
        super().__init__()
        dev = _default_device(device)
        self.input_dim, self.output_dim, self.mask_zero = int(input_dim), int(output_dim), bool(mask_zero)
        self.num_tables, self.hash_mod = int(num_tables), int(hash_mod)
        self.table_rows = None if table_rows is None else [int(r) for r in table_rows]
        if self.table_rows is not None:
            if min(self.table_rows) < 1:
                raise ValueError("every table needs at least one row")
            self.num_tables = len(self.table_rows)
        rows = self.input_dim * self.num_tables if self.table_rows is None else sum(self.table_rows)
        w = torch.empty(rows, self.output_dim, dtype=torch.float32, device=dev)
        w.uniform_(-0.05, 0.05, generator=generator)
        self.register_buffer("embeddings", w)
        if self.table_rows is not None:
            starts = torch.tensor([0] + self.table_rows[:-1], dtype=torch.int64).cumsum(0)
            self.register_buffer("_row_offset", starts.to(dev), persistent=False)
            self.register_buffer("_table_rows_dev", torch.tensor(self.table_rows, dtype=torch.int64, device=dev), persistent=False)
        else:
            self.register_buffer("_row_offset", torch.arange(self.num_tables, dtype=torch.int64, device=dev) * self.input_dim
                                 if self.num_tables > 1 else None, persistent=False)
        # makes autograd call the lookups' backward although the table itself is not a leaf
        self._anchor = torch.zeros((), dtype=torch.float32, device=dev, requires_grad=True)
        self.pending: List[LookupGroup] = []
        self.opt_state = {}
        # The (row, position) sort of the backward depends on the ids only: it is started on a side stream
        # when the lookup runs and overlaps the forward / MLPs (rb_sparse_bwd_prepare / _apply).
        self.presort = True
        import os
        self.validate_ids = os.environ.get("RB_VALIDATE_IDS", "0") == "1"     # per-table id check before every lookup (check_ids)
        self.presort_at = os.environ.get("RB_PRESORT_AT", "start")      # "start": DLRM starts the sort before its bottom MLP
        # How the fused lookups copy table rows (rb_row_cache): "auto" lets the device decide per step from the hot-row census
        # the pre-sort takes of the ids (rows with >= 64 lookups holding more than a quarter of them: through L1)
        self.row_cache = "auto"
        self._hot_rows = torch.zeros(1, dtype=torch.int32, device=dev)
        self._side_stream: Optional[torch.cuda.Stream] = None
        self._sorted = None          # (idx tensor, L, selector, done event) of the sort in flight
        self._sort_ws: Optional[torch.Tensor] = None
        # The row update only needs the sorted pairs and dE: it is launched on the side stream as soon as the
        # interaction backward has written dE, so it overlaps the rest of the backward (bottom MLP) and the dense
        # optimizer step; join() brings the streams back together.
        # Fused row update (optimizers.*.fuse_sparse_updates): the optimizer whose row update the fused lookup's backward
        # applies to the rows a step touches once; None = every row goes through apply_pending()
        self.fused_optimizer = None
        self._single: Optional[torch.Tensor] = None       # uint8 [n]: flags written by the pre-sort (rb_sparse_bwd_mark_singletons)
        self._fused_done = False
        self._grad_ready: Optional[torch.cuda.Event] = None
        self.after_grad_hooks: list = []      # called with the event above once the fused lookup's backward is queued
        self._apply_done: Optional[torch.cuda.Event] = None
        self._inflight = None

    # -- helpers used by the autograd functions
    def cap_ids(self, ids: torch.Tensor) -> torch.Tensor:
        """row-in-table = id mod rows(table) for non-negative raw ids [B, num_tables] (the MLPerf-DLRM capping of the
        Criteo-Terabyte cardinalities)."""
        if self.table_rows is None:
            return ids % self.input_dim
        return ids % self._table_rows_dev[None].to(ids.dtype)

    def row_offset_for(self, L: int):
        if self.num_tables == 1 and self.table_rows is None:
            return None
        if L != self.num_tables:
            raise ValueError(f"a {self.num_tables}-table embedding takes [B, {self.num_tables}] indices, got last dim {L}")
        return self._row_offset

    def check_ids(self, idx: torch.Tensor) -> None:
        """Per-table range check of raw ids [..., num_tables] (rb_check_indices): raises the device's out-of-range flag, which
        ops.check_oob turns into the IndexError keras.layers.Embedding gives on CPU.  The lookups check the final row against
        the total row count only — with several tables in one tensor an id >= rows(f) would otherwise read table f + 1.
        Called by every lookup when `validate_ids` is set (RB_VALIDATE_IDS=1); one extra pass over the ids."""
        if self.num_tables == 1 and self.table_rows is None:
            return                                   # one table: the kernels' own check is the per-table check
        if getattr(self, "_table_rows_dev", None) is None:
            self._table_rows_dev = torch.full((self.num_tables,), self.input_dim, dtype=torch.int64, device=self.embeddings.device)
        ops.check_indices(idx, self._table_rows_dev, hash_mod=self.hash_mod)

    def _record(self, group: LookupGroup) -> None:
        self.pending.append(group)

    def _presort(self, idx: torch.Tensor, L: int, field_row_offset) -> None:
        """Start keys + radix sort for this lookup on the side stream (first lookup of the step only;
        a table used several times per step falls back to the one-call path in apply_pending)."""
        if not self.embeddings.is_cuda:
            return
        if not (self.presort and torch.is_grad_enabled()) or self._sorted is not None or self.pending:
            if self._sorted is not None and self._sorted[0] is not idx:
                self._sorted = (None, 0, 0, self._sorted[3])       # a second use: the pre-sort no longer covers the step
            return
        n = idx.numel()
        if n == 0:
            return
        rows, D = self.embeddings.shape
        need = ops.lib.rb_sparse_bwd_update_workspace_bytes(n, D, rows)
        if self._sort_ws is None or self._sort_ws.numel() < need:
            self._sort_ws = ops.sparse_workspace(n, D, rows, self.embeddings.device)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.embeddings.device, priority=_stream_priority("RB_PRIO_SIDE"))
        main, side = torch.cuda.current_stream(), self._side_stream
        side.wait_stream(main)            # ids are ready; the previous step's apply has released the workspace
        with torch.cuda.stream(side):
            sel = ops.sparse_bwd_prepare(rows, D, [LookupGroup(idx, L, None, field_row_offset=field_row_offset, hash_mod=self.hash_mod)],
                                         self._sort_ws, hot_rows_flag=self._hot_rows)
            if self.fused_optimizer is not None and self.hash_mod == 0 and D % 4 == 0:
                if self._single is None or self._single.numel() < n:
                    self._single = torch.empty(n, dtype=torch.uint8, device=self.embeddings.device)
                self._sel_compact = ops.sparse_bwd_mark_singletons(rows, D, n, self._sort_ws, sel, self._single)
                self._single_for = idx.data_ptr()
            done = side.record_event()
        if not torch.cuda.is_current_stream_capturing():
            idx.record_stream(side)
        self._sorted = (idx, L, sel, done)

    def start_presort(self, idx: torch.Tensor) -> torch.Tensor:
        if self.presort_at == "after_lookup":
            return idx.contiguous()
        return self._start_presort(idx)

    def _start_presort(self, idx: torch.Tensor) -> torch.Tensor:
        """Start the backward's (row, position) sort for `idx` NOW on the side stream — a model calls this before work that
        does not depend on the ids (DLRM's bottom MLP) so that the sort runs under it instead of beside the lookup.  Returns
        the (contiguous) index tensor the later lookup must be given."""
        idx = idx.contiguous()
        if self.validate_ids:
            self.check_ids(idx)
        L = idx.shape[-1] if idx.dim() >= 1 else 1
        self._presort(idx, L, self.row_offset_for(L))
        return idx

    # -- the Keras call surface
    def forward(self, idx: torch.Tensor) -> torch.Tensor:
        """E[..., :] = embeddings[idx[...], :]   (ctr/model.py:19, :49)."""
        idx = idx.contiguous()
        if self.validate_ids:
            self.check_ids(idx)
        L = idx.shape[-1] if idx.dim() >= 1 else 1
        self._presort(idx, L, self.row_offset_for(L) if self.num_tables > 1 else None)
        return _GatherFn.apply(self._anchor, self, idx)

    def compute_mask(self, x, mask=None):
        return (x != 0) if self.mask_zero else None   # dien/model.py:25

    # -- fused forms of the call sites around the lookup
    def pooled(self, idx: torch.Tensor, mode: str = "sum", mask_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
        """sum / mean / masked mean over axis 1 of the looked-up rows without materialising them
        (ctr/model.py:21; dien/layers.py:5-17 with mask = mask_idx != 0, default idx != 0)."""
        return _BagPoolFn.apply(self._anchor, self, idx, mode, mask_idx)

    def lookup_fm(self, idx: torch.Tensor):
        """(E[B,F,D], fm[B]) of ctr/model.py:19-23 in one pass over the rows."""
        idx = idx.contiguous()
        if self.validate_ids:
            self.check_ids(idx)
        self._presort(idx, idx.shape[1], self.row_offset_for(idx.shape[1]))
        return _GatherFMFn.apply(self._anchor, self, idx)

    def lookup_fm_deep(self, idx: torch.Tensor, dense: torch.Tensor, ld: int):
        """(deep bf16 [B, ld], fm [B]) of ctr/model.py:19-26: the lookup, the FM second order and the MLP's padded input row
        [flatten(E) | dense | 1 | 0...] in one pass over the rows (rb_gather_fm_deep_fwd)."""
        idx = idx.contiguous()
        if self.validate_ids:
            self.check_ids(idx)
        self._presort(idx, idx.shape[1], self.row_offset_for(idx.shape[1]))
        return _GatherFMDeepFn.apply(self._anchor, self, idx, dense.float().contiguous(), int(ld))

    def interact(self, idx: torch.Tensor, dense_vec: torch.Tensor, self_interaction=False, skip_gather=True, tail=True,
                 out_dtype=torch.float32, pad_to=1, ones_col=False):
        """ctr/model.py:49-55 fused: [DotInteraction([E ; dense_vec]) || dense_vec] with E read
        straight from the table, so [B,F,D] and [B,F+1,D] never exist in HBM.  out_dtype=bfloat16
        emits the row in bf16, zero-padded to a multiple of `pad_to` columns (the K operand of a
        bf16 top MLP); its gradient then comes back in the same padded bf16 form."""
        idx = idx.contiguous()
        if self.validate_ids:
            self.check_ids(idx)
        late = self.presort_at == "after_lookup"
        if not late:
            self._presort(idx, idx.shape[1], self.row_offset_for(idx.shape[1]))
        out = _InteractFn.apply(self._anchor, self, idx, dense_vec.float(), self_interaction, skip_gather, tail, out_dtype, pad_to,
                                ones_col)
        if late:        # the sort's HBM passes start once the lookup kernel is queued ahead of them
            self._presort(idx, idx.shape[1], self.row_offset_for(idx.shape[1]))
        return out

    def _ensure_state(self, kind: str, initial_accumulator_value=0.1):
        """(state0, state1, created-just-now) of optimizer `kind` for this table."""
        st = self.opt_state
        fresh = False
        if kind in ("adam_lazy", "adam_tf_dense"):
            if "m" not in st:
                st["m"], st["v"] = torch.zeros_like(self.embeddings), torch.zeros_like(self.embeddings)
                fresh = True
            return st["m"], st["v"], fresh
        if kind == "adagrad":
            if "acc" not in st:
                st["acc"] = torch.full_like(self.embeddings, initial_accumulator_value)
                fresh = True
            return st["acc"], None, fresh
        if kind == "sgd":
            return None, None, False
        raise ValueError(kind)

    def _fused_update_args(self, idx: torch.Tensor):
        """Keyword arguments of ops.dot_interaction_bwd_update when this backward may apply the optimizer to the rows the step
        touches once, else None: an optimizer was armed (fuse_sparse_updates), its update is row-sparse, this lookup is the
        one the pre-sort covers and no other use of the table has recorded a gradient in this step."""
        opt = self.fused_optimizer
        srt = self._sorted
        if (opt is None or srt is None or srt[0] is None or self.pending or self._single is None
                or srt[0].data_ptr() != idx.data_ptr() or srt[0].numel() != idx.numel()
                or getattr(self, "_single_for", None) != idx.data_ptr()):
            return None
        kind = opt.sparse_kind
        if kind not in ("sgd", "adagrad", "adam_lazy"):
            return None
        kw = dict(opt._sparse_kwargs())
        s0, s1, _ = self._ensure_state(kind, kw.pop("initial_accumulator_value", 0.1))
        torch.cuda.current_stream().wait_event(srt[3])          # the flags are written by the pre-sort's stream
        step = opt.iterations if opt._prepared else opt.iterations + 1
        return dict(single=self._single[: idx.numel()], state0=s0, state1=s1, optimizer=kind, step=step, **kw)

    def join(self) -> None:
        """Wait (on the current stream) for a row update launched on the side stream."""
        if self._apply_done is not None:
            torch.cuda.current_stream().wait_event(self._apply_done)
            self._apply_done = None
        self._inflight = None

    # -- optimizer side (called by optimizers.*.apply_gradients)
    def apply_pending(self, kind: str, step: int, lr: float, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
                      initial_accumulator_value=0.1, alpha_dev=None) -> int:
        """Runs the fused backward scatter + row update over this step's lookup groups."""
        if not self.pending:
            if kind == "adam_tf_dense":   # Keras moves every row every step, gradient or not
                raise NotImplementedError("adam_tf_dense with no lookup in the step")
            return 0
        # fresh_state: state tensors created just now on the current stream: the side stream must see them
        s0, s1, fresh_state = self._ensure_state(kind, initial_accumulator_value)
        groups, self.pending = self.pending, []
        skip, self._fused_done = self._fused_done, False
        if skip and len(groups) != 1:
            raise RuntimeError("the fused row update (fuse_sparse_updates) serves tables looked up once per step; this step recorded "
                               f"{len(groups)} uses after rows were already updated in the backward")
        sorted_, self._sorted = self._sorted, None
        n = sum(g.n for g in groups)
        if sorted_ is not None:
            torch.cuda.current_stream().wait_event(sorted_[3])     # also orders later reuse of the sort workspace
        grad_ready, self._grad_ready = self._grad_ready, None
        if (sorted_ is not None and sorted_[0] is not None and len(groups) == 1 and groups[0].L == sorted_[1]
                and groups[0].idx.data_ptr() == sorted_[0].data_ptr() and groups[0].n == sorted_[0].numel()):
            if grad_ready is not None and self._side_stream is not None:
                side = self._side_stream
                if fresh_state:
                    side.wait_stream(torch.cuda.current_stream())
                else:
                    side.wait_event(grad_ready)
                with torch.cuda.stream(side):
                    ops.sparse_bwd_apply(self.embeddings, s0, s1, groups, self._sort_ws, self._sel_compact if skip else sorted_[2],
                                         optimizer=kind, step=step, lr=lr, beta_1=beta_1, beta_2=beta_2, epsilon=epsilon,
                                         alpha_dev=alpha_dev, skip_singletons=skip)
                    self._apply_done = side.record_event()
                self._inflight = groups          # keeps dE alive until join(): the allocator must not hand it out meanwhile
                return n
            ops.sparse_bwd_apply(self.embeddings, s0, s1, groups, self._sort_ws, self._sel_compact if skip else sorted_[2],
                                 optimizer=kind, step=step, lr=lr, beta_1=beta_1, beta_2=beta_2, epsilon=epsilon, alpha_dev=alpha_dev,
                                 skip_singletons=skip)
        else:
            if skip:
                raise RuntimeError("fused row update: the pre-sorted pairs of this step are gone")
            ops.sparse_bwd_update(self.embeddings, s0, s1, groups, optimizer=kind, step=step, lr=lr, beta_1=beta_1,
                                  beta_2=beta_2, epsilon=epsilon, alpha_dev=alpha_dev)
        return n


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _LinearBF16Fn(torch.autograd.Function):
    """y = x @ W + b on bf16 tensor cores (cuBLASLt through torch), fp32 master weights.

    x is bf16 [B, Kp] with Kp = in_dim rounded up to 8 (zero pad columns) so that cuBLAS never
    falls back to its unaligned legacy kernels; W [in_dim, out] and b [out] are the fp32
    parameters.  Backward: dx = dy W^T (bf16, padded), dW = x^T dy, db = 1^T dy as a skinny GEMM
    (a column reduction over 65536 rows is ~6x slower as an elementwise reduce kernel)."""

    @staticmethod
    def _bf16_shadow(p: torch.Tensor, rows: Optional[int] = None) -> torch.Tensor:
        """bf16 copy of a parameter ([rows, out] zero-padded below for a kernel).  It is cached on the parameter and
        kept in step by the fused optimizer step (rb_dense_opt_step writes it together with the fp32 value); any
        torch-side in-place change of the parameter bumps its version and forces a rebuild here."""
        sh = getattr(p, "_rb_shadow", None)
        shape = tuple(p.shape) if rows is None else (rows, p.shape[1])
        if sh is None or sh[1] != p._version or tuple(sh[0].shape) != shape:
            with torch.no_grad():
                t = torch.zeros(shape, dtype=torch.bfloat16, device=p.device)
                t.reshape(-1)[: p.numel()].copy_(p.detach().reshape(-1))
            p._rb_shadow = sh = (t, p._version)
        return sh[0]

    @staticmethod
    def forward(ctx, x, W, b, need_dx, ones_col=False):
        """ones_col: the caller set pad column `in_dim` of x to 1.0 (the matching row of the padded kernel is zero,
        so the output is unchanged); then row `in_dim` of x^T dy IS the bias gradient and no column sum runs."""
        in_dim, out = W.shape
        Kp = x.shape[1]
        Wp = _LinearBF16Fn._bf16_shadow(W, Kp)
        y = torch.addmm(_LinearBF16Fn._bf16_shadow(b), x, Wp)
        ctx.save_for_backward(x, Wp)
        ctx.in_dim, ctx.need_dx, ctx.ones_col = in_dim, need_dx, bool(ones_col) and Kp > in_dim
        return y

    @staticmethod
    def backward(ctx, dy):
        x, Wp = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.mm(dy, Wp.t()) if ctx.need_dx else None
        dW_full = ops.mm_f32_out(x.t(), dy)                                       # fp32 accumulate AND fp32 result
        dW = dW_full[: ctx.in_dim]
        if ctx.ones_col:
            return dx, dW, dW_full[ctx.in_dim].clone(), None, None               # bias gradient came with the GEMM
        if dy.shape[1] % 8 == 0 and dy.shape[1] <= 2048:
            db = ops.colsum(dy)                       # rb_colsum: deterministic fp32 column sums
        else:
            ones = torch.ones(1, dy.shape[0], dtype=dy.dtype, device=dy.device)
            db = ops.mm_f32_out(ones, dy).reshape(-1)
        return dx, dW, db, None, None


class _DenseStackFn(torch.autograd.Function):
    """The whole MLP of ctr/layers.py:5-14 on the tcgen05 Dense kernels (csrc/mlp.cu): hidden layers linear, bf16 operands with
    fp32 accumulation, fp32 master weights, the last layer's activation applied on the fp32 accumulator.

    x is either the padded bf16 K operand [B, Kp] (what the fused interaction kernel emits; `ones_col` says its column
    `in_dim` holds 1.0) or the raw f32 [B, in_dim] features, which are packed here (rb_dense_pack_input).  One C-ABI call
    per product: y = x W + b (rb_dense_fwd), dx = dy W^T (rb_dense_bwd_input), dW = x^T dy (rb_dense_bwd_weight; with the
    ones column its row `in_dim` is the bias gradient), and the Dense(1) head as a row dot product (rb_dense_head_*).

    Backward: only the dx chain is on the step's critical path (it feeds the interaction backward and, through it, the
    sparse row update).  The weight and bias gradients run on the module's own stream beside it and are written straight
    into the parameters' .grad; the streams meet again when the backward pass ends (`MLP._join_wgrad`, queued as an
    autograd-engine callback), so readers of .grad on the caller's stream see finished values."""

    @staticmethod
    def forward(ctx, x, mlp, Kp, ones_col, need_dx, *params):
        Ws, bs = params[0::2], params[1::2]
        n = len(Ws)
        in_dim, final_activation = mlp.in_dim, mlp.final_activation
        raw = x.dtype != torch.bfloat16
        label, mlp._bce_label = mlp._bce_label, None          # set by MLP.forward(bce_label=...): fuse the loss into the head
        fused_head = None
        if raw:
            ones_col = Kp > in_dim
            x = ops.dense_pack_input(x.float(), Kp, ones_col)
        acts = [x]
        shadows = []
        h = x
        for i, (W, b) in enumerate(zip(Ws, bs)):
            last = i == n - 1
            Wp = _LinearBF16Fn._bf16_shadow(W, Kp if i == 0 else None)
            shadows.append(Wp)
            if last and W.shape[1] == 1 and label is not None:
                # Dense(1, sigmoid) + clipped BCE + the head's backward in one pass over h (rb_dense_head_bce): `out` is the LOSS
                want_dx = n > 1 or need_dx
                want_cs = n > 1 and not (n == 2 and bool(ones_col) and Kp > in_dim)
                prob, out, dxh, dwh, dbh, csh = ops.dense_head_bce(h, Wp.reshape(-1), b, label, want_dx=want_dx, want_dx_colsum=want_cs)
                fused_head = (dxh, dwh, dbh, csh)
                mlp.last_prob = prob
            elif last and W.shape[1] == 1:
                out = ops.dense_head_fwd(h, Wp.reshape(-1), b, final_activation).reshape(-1, 1)
            elif last:
                out = ops.dense_fwd(h, Wp, b, final_activation, torch.float32)
            else:
                h = ops.dense_fwd(h, Wp, b, None, torch.bfloat16)
                acts.append(h)
        ctx.save_for_backward(out, *acts, *shadows)
        ctx.mlp, ctx.n, ctx.ones_col, ctx.need_dx, ctx.raw = mlp, n, bool(ones_col) and Kp > in_dim, need_dx, raw
        ctx.fused_head = fused_head
        return out

    @staticmethod
    def backward(ctx, dout):
        n, mlp = ctx.n, ctx.mlp
        in_dim, act = mlp.in_dim, mlp.final_activation
        saved = ctx.saved_tensors
        out, acts, shadows = saved[0], saved[1:1 + n], saved[1 + n:]
        Ws, bs = list(mlp.kernels), list(mlp.biases)
        dout = dout.contiguous().float()
        main = torch.cuda.current_stream()
        wg, ws = mlp._wgrad_begin(acts, shadows)

        def assign(p, g):           # runs inside the wgrad stream context
            if p.grad is None:
                p.grad = g
            else:
                p.grad.add_(g)

        def wgrad(i, x_i, dy_i, have_db=None):
            """dW_i (and db_i) on the wgrad stream, from operands the main stream has just produced."""
            W, b = Ws[i], bs[i]
            dW_full = torch.empty(x_i.shape[1], W.shape[1], dtype=torch.float32, device=x_i.device)
            need_db = not (i == 0 and ctx.ones_col) and have_db is None
            db = torch.empty(W.shape[1], dtype=torch.float32, device=x_i.device) if need_db else have_db
            mlp._wgrad_keep.extend((x_i, dy_i, dW_full, db))

            def launch():
                ops.dense_bwd_weight(x_i, dy_i, out=dW_full, ws=ws)
                if need_db:
                    ops.colsum(dy_i, out=db, ws=ws)
                if i == 0:
                    assign(W, dW_full[:in_dim])
                    assign(b, dW_full[in_dim] if db is None else db)
                else:
                    assign(W, dW_full)
                    assign(b, db)

            if mlp.defer_wgrad:
                mlp._wgrad_deferred.append(launch)      # launched by flush_wgrad(): after the caller's critical kernels are queued
                return
            wg.wait_event(main.record_event())
            with torch.cuda.stream(wg):
                launch()

        h, Wp = acts[n - 1], shadows[n - 1]
        want_dx = n > 1 or ctx.need_dx
        head_cs = None
        if Wp.shape[1] == 1:
            # the head also sums the columns of the dx it writes: the bias gradient of the layer below, without re-reading dx
            want_cs = n > 1 and not (n == 2 and ctx.ones_col)
            if ctx.fused_head is not None:        # computed with the loss in the forward (d loss = 1: `dout` is the seed of backward())
                dy, dw, db, head_cs = ctx.fused_head
            else:
                res = ops.dense_head_bwd(dout.reshape(-1), out.reshape(-1), act, h, Wp.reshape(-1), want_dx=want_dx, want_dx_colsum=want_cs)
                dy, dw, db = res[:3]
                head_cs = res[3] if want_cs else None
            dw = dw[:in_dim] if n == 1 else dw
            for p_, g_ in ((Ws[n - 1], dw.reshape(-1, 1)), (bs[n - 1], db)):
                if p_.grad is None:
                    p_.grad = g_
                else:
                    p_.grad.add_(g_)
        else:
            dyp = ops.dense_act_bwd(dout, out, act)
            dy = ops.dense_bwd_input(dyp, Wp) if want_dx else None
            wgrad(n - 1, h, dyp)
        for i in range(n - 2, -1, -1):
            dy_i = dy
            dy = ops.dense_bwd_input(dy_i, shadows[i]) if (i > 0 or ctx.need_dx) else None
            wgrad(i, acts[i], dy_i, head_cs if i == n - 2 else None)
        mlp._wgrad_end()
        dx = dy
        if dx is not None and ctx.raw:
            dx = dx[:, :in_dim].float()
        return (dx if ctx.need_dx else None, None, None, None, None, *([None] * (2 * n)))


class _CollapsedAffineFn(torch.autograd.Function):
    """The affine part of an MLP whose hidden layers are linear (ctr/layers.py:8 builds them without activation):

        z = ((x W1 + b1) W2 + b2) ... Wn + bn  =  x A_n + c_n,     A_k = W1 ... Wk,   c_k = c_{k-1} W_k + b_k.

    Opt-in (MLP(collapse_linear=True)); the default path runs the layers one GEMM at a time like the reference.  Only
    three passes touch the batch: z = x A_n + c_n, G = x^T dz (+ s = 1^T dz) and dx = dz A_n^T.  Every layer's own
    gradient follows from G and s in weight space, with S_k = W_{k+1} ... W_n:

        dW_k = (A_{k-1}^T G + c_{k-1} s^T) S_k^T,      db_k = S_k s            (A_0 = I, c_0 = 0, S_n = I)

    which is H_{k-1}^T (dz S_k^T) with H_{k-1} = x A_{k-1} + 1 c_{k-1}^T written out.  Same function, same parameters,
    same gradients in exact arithmetic; in floating point the products are associated differently (weight-space
    products in fp32).  x may be bf16, zero-padded to Kp columns, with column in_dim set to 1.0 (ones_col): the padded
    rows of A_n are zero, so the pads do not contribute, and row in_dim of x^T dz is s."""

    @staticmethod
    def forward(ctx, x, in_dim, need_dx, ones_col, *params):
        Ws, bs = params[0::2], params[1::2]
        A, c = [Ws[0]], [bs[0]]
        for W, b in zip(Ws[1:], bs[1:]):
            A.append(A[-1] @ W)
            c.append(torch.addmv(b, W.t(), c[-1]))
        Kp = x.shape[1]
        Ap = A[-1]
        if Kp != in_dim:
            Ap = torch.zeros(Kp, Ap.shape[1], dtype=Ap.dtype, device=Ap.device)
            Ap[:in_dim] = A[-1]
        Ax = Ap.to(x.dtype)
        z = torch.addmm(c[-1].to(x.dtype), x, Ax)
        ctx.save_for_backward(x, Ax, *A[:-1], *c[:-1], *Ws)
        ctx.n, ctx.in_dim, ctx.need_dx, ctx.ones_col = len(Ws), in_dim, need_dx, bool(ones_col) and Kp > in_dim
        return z

    @staticmethod
    def backward(ctx, dz):
        n = ctx.n
        saved = ctx.saved_tensors
        x, Ax = saved[0], saved[1]
        A, c, Ws = saved[2:2 + n - 1], saved[2 + n - 1:2 + 2 * (n - 1)], saved[2 + 2 * (n - 1):]
        dz = dz.contiguous()
        bf16 = dz.dtype == torch.bfloat16
        G_full = ops.mm_f32_out(x.t(), dz) if bf16 else torch.mm(x.t(), dz)      # fp32 accumulate AND fp32 result
        G = G_full[: ctx.in_dim]
        if ctx.ones_col:
            s = G_full[ctx.in_dim]
        elif bf16 and dz.shape[1] % 8 == 0 and dz.shape[1] <= 2048:
            s = ops.colsum(dz)                                                    # rb_colsum: deterministic fp32 column sums
        else:
            s = dz.float().sum(0)
        dx = torch.mm(dz, Ax.t()) if ctx.need_dx else None
        grads = [None] * (2 * n)
        S = None                                                                # S_k, starting from S_n = I
        for k in range(n - 1, -1, -1):
            left = G if k == 0 else torch.addmm(torch.outer(c[k - 1], s), A[k - 1].t(), G)
            grads[2 * k] = left if S is None else left @ S.t()
            grads[2 * k + 1] = s.clone() if S is None else S @ s
            S = Ws[k] if S is None else Ws[k] @ S
        return (dx, None, None, None, *grads)


class MLP(nn.Module):
    """ctr/layers.py:5-14: Dense layers whose HIDDEN layers are linear; only the last layer has
    `final_activation` (None | 'relu' | 'sigmoid').  Kernels are [in, units] Glorot-uniform, biases
    zero (Keras defaults), built on the first call like Keras does.  Dense and data-parallel:
    it runs on cuBLAS through torch and is not part of the sparse hot path.

    compute_dtype=torch.bfloat16 runs the GEMMs on bf16 tensor cores with fp32 master weights and
    fp32 accumulation (activations between layers are bf16; the final activation is applied in
    fp32).  In that mode the input may arrive already as bf16 with its feature axis zero-padded
    to a multiple of 8 (what the fused interaction kernel emits)."""

    def __init__(self, units: Sequence[int], final_activation=None, *, compute_dtype: Optional[torch.dtype] = None,
                 generator: Optional[torch.Generator] = None, collapse_linear: bool = False):
        super().__init__()
        self.collapse_linear = bool(collapse_linear)     # opt-in: evaluate the linear stack as ONE affine map (_CollapsedAffineFn)
        self.backend = "tcgen05"                          # bf16 mode: csrc/mlp.cu kernels; "cublas" = torch.addmm / mm
        self._wgrad_stream: Optional[torch.cuda.Stream] = None
        self._wgrad_ws: Optional[torch.Tensor] = None
        self._wgrad_done: Optional[torch.cuda.Event] = None
        self._wgrad_keep: list = []
        self._wgrad_join_queued = False
        self._bce_label = None                            # one-shot: the next tcgen05 forward fuses Dense(1, sigmoid) + clipped BCE
        self.last_prob: Optional[torch.Tensor] = None     # probabilities of the last fused-loss forward
        self.defer_wgrad = False                          # hold the weight-gradient products back until flush_wgrad()
        self._wgrad_deferred: list = []
        if final_activation not in (None, "relu", "sigmoid"):
            raise ValueError(final_activation)
        if compute_dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError("compute_dtype must be None/float32 or bfloat16")
        self.units, self.final_activation = list(units), final_activation
        self.compute_dtype = None if compute_dtype == torch.float32 else compute_dtype
        self._generator = generator
        self.in_dim: Optional[int] = None
        self.kernels = nn.ParameterList()
        self.biases = nn.ParameterList()

    def build(self, in_dim: int, device) -> None:
        self.in_dim = int(in_dim)
        for u in self.units:
            lim = math.sqrt(6.0 / (in_dim + u))
            w = torch.empty(in_dim, u, dtype=torch.float32, device=device).uniform_(-lim, lim, generator=self._generator)
            self.kernels.append(nn.Parameter(w))
            self.biases.append(nn.Parameter(torch.zeros(u, dtype=torch.float32, device=device)))
            in_dim = u

    def load_arrays(self, layers, device) -> None:
        """Adopt [(kernel[in,units], bias[units])...] (numpy or tensors): the oracle owns the init in
        parity tests because TF's RNG streams cannot be matched."""
        self.kernels = nn.ParameterList(nn.Parameter(torch.as_tensor(W, dtype=torch.float32).to(device).contiguous()) for W, _ in layers)
        self.biases = nn.ParameterList(nn.Parameter(torch.as_tensor(b, dtype=torch.float32).to(device).contiguous()) for _, b in layers)
        self.in_dim = int(self.kernels[0].shape[0])

    # -- weight gradients beside the dx chain (see _DenseStackFn) -------------------------------------------------------
    def _wgrad_begin(self, acts, shadows):
        dev = acts[0].device
        if self._wgrad_stream is None:
            self._wgrad_stream = torch.cuda.Stream(device=dev, priority=_stream_priority("RB_PRIO_WGRAD", -1))
        rows = acts[0].shape[0]
        need = 256
        for a, w in zip(acts, shadows):
            if w.shape[1] > 1:
                need = max(need, ops.dense_bwd_weight_workspace_bytes(rows, a.shape[1], w.shape[1]),
                           int(ops.lib.rb_colsum_workspace_bytes(rows, w.shape[1])))
        if self._wgrad_ws is None or self._wgrad_ws.numel() < need:
            self._wgrad_ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self._wgrad_keep.append(self._wgrad_ws)
        return self._wgrad_stream, self._wgrad_ws

    def _wgrad_end(self) -> None:
        self._wgrad_done = self._wgrad_stream.record_event()
        if not self._wgrad_join_queued:
            self._wgrad_join_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._join_wgrad)

    def flush_wgrad(self, after: Optional[torch.cuda.Event] = None) -> None:
        """Launch the weight-gradient products held back by `defer_wgrad` on the wgrad stream, ordered after `after` (default:
        everything queued on the current stream so far).  DLRM defers the top tower's products until the interaction backward
        is queued: the persistent Dense CTAs would otherwise occupy the SMs just when that kernel becomes ready (r2_19
        timeline: an 18 us hole in front of it), and under the HBM-bound row update they cost next to nothing."""
        if not self._wgrad_deferred:
            return
        todo, self._wgrad_deferred = self._wgrad_deferred, []
        wg = self._wgrad_stream
        wg.wait_event(after if after is not None else torch.cuda.current_stream().record_event())
        with torch.cuda.stream(wg):
            for launch in todo:
                launch()
        self._wgrad_done = wg.record_event()

    def _join_wgrad(self) -> None:
        """The caller's stream waits for the weight gradients; operands the wgrad stream was reading may be freed after that."""
        self._wgrad_join_queued = False
        self.flush_wgrad()
        if self._wgrad_done is not None:
            torch.cuda.current_stream().wait_event(self._wgrad_done)
            self._wgrad_done = None
        self._wgrad_keep.clear()

    def _tcgen05_ok(self, x: torch.Tensor) -> bool:
        """The hand-written tcgen05 Dense kernels serve CUDA inputs when every hidden width is a multiple of 8 (16-byte row
        strides for TMA) and the last layer is either such a width or the Dense(1) head; `backend='cublas'` keeps the
        torch / cuBLASLt path (the round-1 implementation, used as the A/B baseline in bench.py --mlp-backend)."""
        if self.backend != "tcgen05" or not x.is_cuda or x.dim() != 2:
            return False
        us = [int(k.shape[1]) for k in self.kernels]
        if any(u % 8 for u in us[:-1]) or (us[-1] != 1 and us[-1] % 8):
            return False
        return not (us[-1] == 1 and int(self.kernels[-1].shape[0]) % 8)

    def padded_in_dim(self) -> int:
        """Feature count the bf16 path wants its input padded to (multiple of 8)."""
        return _round_up(self.in_dim, 8)

    def _activate(self, x: torch.Tensor) -> torch.Tensor:
        x = x.float()
        if self.final_activation == "relu":
            x = torch.relu(x)
        elif self.final_activation == "sigmoid":
            x = torch.sigmoid(x)
        return x

    def can_fuse_bce(self, x: torch.Tensor) -> bool:
        """The last layer is Dense(1, sigmoid) on the tcgen05 path with a width rb_dense_head_bce serves."""
        return (self.compute_dtype == torch.bfloat16 and len(self.kernels) >= 1 and self.final_activation == "sigmoid"
                and int(self.kernels[-1].shape[1]) == 1 and ops.dense_head_bce_ok(int(self.kernels[-1].shape[0]))
                and not (self.collapse_linear and len(self.kernels) >= 2) and self._tcgen05_ok(x))

    def forward(self, x: torch.Tensor, ones_col: bool = False, bce_label: Optional[torch.Tensor] = None) -> torch.Tensor:
        """ones_col=True: x arrives bf16, padded, with pad column `in_dim` already set to 1.0 (the fused interaction
        kernel does that), which lets the first layer read its bias gradient off the weight-gradient GEMM.
        bce_label (only when can_fuse_bce(x)): returns the LOSS f32[1] = mean clipped binary cross-entropy of the probabilities
        against the labels instead of the probabilities — head, loss and the head's backward in one kernel; the backward pass
        must be seeded with 1 (loss.backward())."""
        if len(self.kernels) == 0:
            self.build(x.shape[-1], x.device)
        if bce_label is not None:
            if not self.can_fuse_bce(x):
                raise RuntimeError("bce_label needs the bf16 tcgen05 path and a Dense(1, sigmoid) head of 8..256 inputs")
            self._bce_label = bce_label.contiguous()
        collapse = self.collapse_linear and len(self.kernels) >= 2
        if self.compute_dtype is None:
            if collapse:
                flat = [t for Wb in zip(self.kernels, self.biases) for t in Wb]
                return self._activate(_CollapsedAffineFn.apply(x.float(), self.in_dim, x.requires_grad, False, *flat))
            for i, (W, b) in enumerate(zip(self.kernels, self.biases)):
                x = torch.addmm(b, x.float(), W)
            return self._activate(x)
        # bf16 tensor-core path
        Kp = self.padded_in_dim()
        need_dx = x.requires_grad
        collapse_ = self.collapse_linear and len(self.kernels) >= 2
        if not collapse_ and self._tcgen05_ok(x) and (x.dtype != torch.bfloat16 or x.shape[-1] != Kp):
            if x.shape[-1] != self.in_dim:
                raise ValueError(f"MLP built for {self.in_dim} input features, got {x.shape[-1]}")
            flat = [t for Wb in zip(self.kernels, self.biases) for t in Wb]
            return _DenseStackFn.apply(x.float(), self, Kp, False, need_dx, *flat)   # packed inside
        if x.dtype != torch.bfloat16 or x.shape[-1] != Kp:
            if x.shape[-1] != self.in_dim:
                raise ValueError(f"MLP built for {self.in_dim} input features, got {x.shape[-1]}")
            xp = torch.zeros(x.shape[0], Kp, dtype=torch.bfloat16, device=x.device) if Kp != self.in_dim else None
            if xp is None:
                x = x.to(torch.bfloat16)
            else:
                xp[:, : self.in_dim] = x              # autograd-aware pad (dense inputs carry no gradient in the models)
                xp[:, self.in_dim] = 1.0              # ones column: the bias gradient falls out of the dW GEMM
                x = xp
                ones_col = True
        if collapse:
            flat = [t for Wb in zip(self.kernels, self.biases) for t in Wb]
            return self._activate(_CollapsedAffineFn.apply(x, self.in_dim, need_dx, ones_col, *flat))
        if self._tcgen05_ok(x):
            flat = [t for Wb in zip(self.kernels, self.biases) for t in Wb]
            return _DenseStackFn.apply(x, self, Kp, ones_col, need_dx, *flat)
        for i, (W, b) in enumerate(zip(self.kernels, self.biases)):
            x = _LinearBF16Fn.apply(x, W, b, need_dx or i > 0, ones_col and i == 0)
        return self._activate(x)


class DotInteraction(nn.Module):
    """ctr/layers.py:17-43.  X f32[B,F',D] -> f32[B,F'^2] (skip_gather=True: kept entries in place,
    zeros elsewhere) or the compact [B, F'(F'-1)/2] / [B, F'(F'+1)/2] (skip_gather=False).
    self_interaction=False keeps the strict upper triangle (j > i); True keeps the lower triangle
    including the diagonal (j <= i) — the reference's `upper_matrix` really is band_part(.,-1,0)."""

    def __init__(self, self_interaction: bool, skip_gather: bool):
        super().__init__()
        self.self_interaction, self.skip_gather = bool(self_interaction), bool(skip_gather)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        return _DotInteractionFn.apply(inputs, self.self_interaction, self.skip_gather)


def compute_his_average(embedding: Embedding, idx: torch.Tensor, mask_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dien/layers.py:5-17 fused with its lookup (dien/model.py:14-19,25-31): masked mean over the
    history axis with mask = (mask_idx != 0); an all-pad history yields NaN, as in the reference."""
    return embedding.pooled(idx, "masked_mean", mask_idx)
