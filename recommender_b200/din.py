"""DIN's attention pooling over a behaviour history with the reference's call surface.

`LocalActivationUnit` mirrors dien/layers.py:34-59 (three Dense layers 4E -> 80 -> 40 -> 1, sigmoid / sigmoid / none, masked
weights, weights^T . history) and `DIN` the part of dien/model.py:36-53 that belongs to the embedding path: the two shared
tables, the mask, the target / history lookups and the unit, up to `concat([target, history_representation])`.  The
BatchNorm-MLP head of the reference (dien/layers.py:20-31) is dense, data-parallel work outside the path; pass
any torch module as `head`.

The B200 form (csrc/din.cu): the history [B, L, E] and the 4E-wide feature tensor are never materialised for the padded
positions.  Valid positions are numbered on the device, get one bf16 feature row each, and the three Dense layers run on
the tcgen05 GEMMs of csrc/mlp.cu over that compact [P, 4E] matrix; the weighted sum and both halves of the backward
re-read the fp32 rows straight from the tables.  The number of valid positions P sizes the GEMMs, so one int32 is read
back per call (the only host synchronisation on this path; a fixed-capacity variant would make the step graph-capturable).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
from torch import nn

from . import ops
from .layers import Embedding, _LinearBF16Fn, _round_up
from .ops import GradSource, LookupGroup


class _LocalActivationFn(torch.autograd.Function):
    """forward: rep[B, E]; backward: d_target, the six parameter gradients and — recorded on the embedding layers as lookup
    groups, or returned as the dense gradient of a materialised history — the gradient rows of the history."""

    @staticmethod
    def forward(ctx, target, history, spec, *params):
        W1, b1, W2, b2, W3, b3 = params
        target = target.float().contiguous()
        if history is not None:      # the reference's call surface: a materialised [B, L, E] history, ids = its row numbers
            B, L, E = history.shape
            hist2d = history.float().contiguous().reshape(B * L, E)
            idx = torch.arange(B * L, device=history.device, dtype=torch.int64).reshape(B, L)
            h = ops.DinHistory(hist2d, idx, mask=spec["mask"])
        else:
            h = ops.DinHistory(spec["item_emb"].embeddings, spec["item"], spec["cat_emb"].embeddings if spec["cat_emb"] is not None else None,
                               spec["cat"], mask=spec["mask"])
        E = h.E
        if target.shape != (h.B, E):
            raise ValueError(f"target must be [B, E] = [{h.B}, {E}], got {tuple(target.shape)}")
        if W1.shape[0] != 4 * E:
            raise ValueError(f"the first attention layer reads 4E = {4 * E} columns, its kernel has {W1.shape[0]}")
        off = ops.din_offsets(h)
        P = int(off[-1].item())                      # sizes the GEMMs below: the one host read-back of this path
        Kp = _round_up(4 * E, 8)
        ctx.h, ctx.off, ctx.P, ctx.spec, ctx.materialised = h, off, P, spec, history is not None
        if P == 0:
            ctx.save_for_backward(target)
            return torch.zeros(h.B, E, dtype=torch.float32, device=target.device)
        X = ops.din_build_features(h, target, off, P, Kp)[:P]                                    # dien/layers.py:47-48
        W1p, W2p, W3p = _LinearBF16Fn._bf16_shadow(W1, Kp), _LinearBF16Fn._bf16_shadow(W2), _LinearBF16Fn._bf16_shadow(W3)
        a1 = ops.dense_fwd(X, W1p, b1, "sigmoid", torch.bfloat16)                                # :49
        a2 = ops.dense_fwd(a1, W2p, b2, "sigmoid", torch.bfloat16)                               # :50
        w = ops.dense_head_fwd(a2, W3p.reshape(-1), b3, None)                                    # :51 (masked positions have no row: :52-54)
        rep = ops.din_pool_fwd(h, off, w)                                                        # :55-56
        ctx.save_for_backward(target, X, a1, a2, w, W1p, W2p, W3p)
        return rep

    @staticmethod
    def backward(ctx, d_rep):
        h, off, P, spec = ctx.h, ctx.off, ctx.P, ctx.spec
        E = h.E
        dev = d_rep.device
        if P == 0:
            (target,) = ctx.saved_tensors
            d_hist = torch.zeros(h.B, h.L, E, dtype=torch.float32, device=dev) if ctx.materialised else None
            return (torch.zeros_like(target), d_hist, None, *([None] * 6))
        target, X, a1, a2, w, W1p, W2p, W3p = ctx.saved_tensors
        d_rep = d_rep.float().contiguous()
        dw = ops.din_pool_bwd_weights(h, off, d_rep, P)[:P]
        da2, dW3, db3 = ops.dense_head_bwd(dw, None, None, a2, W3p.reshape(-1), want_dx=True)
        dz2 = ops.dense_act_bwd_bf16(da2, a2, "sigmoid")
        dW2, db2 = ops.dense_bwd_weight(a1, dz2), ops.colsum(dz2)
        da1 = ops.dense_bwd_input(dz2, W2p)
        dz1 = ops.dense_act_bwd_bf16(da1, a1, "sigmoid")
        dW1, db1 = ops.dense_bwd_weight(X, dz1)[: 4 * E], ops.colsum(dz1)
        dX = ops.dense_bwd_input(dz1, W1p)
        dh, dt = ops.din_feature_bwd(h, target, off, dX, w, d_rep, zero_masked=ctx.materialised)
        d_hist = None
        if ctx.materialised:
            d_hist = dh
        else:
            # the IndexedSlices of the two history lookups: row (b, l) of table k is columns [c0, c0 + D_k) of dh[b, l, :];
            # masked positions still count as touched rows with zero gradients (like compute_his_average's)
            mask = h.mask if h.mask is not None else h.idx0
            D0 = h.table0.shape[1]
            for emb, idx, c0 in ((spec["item_emb"], h.idx0, 0), (spec["cat_emb"], h.idx1, D0)):
                if emb is None:
                    continue
                src = dh if c0 == 0 else dh[:, :, c0:]
                emb._record(LookupGroup(idx, h.L, GradSource([src], [h.L * E], [E], scale="masked", mask_idx=mask), hash_mod=emb.hash_mod))
        return (dt, d_hist, None, dW1, db1, dW2, db2, dW3.reshape(-1, 1), db3)


class LocalActivationUnit(nn.Module):
    """dien/layers.py:34-59.  `unit((target, history), mask=mask)` takes the reference's tensors (target [B, 1, E] or [B, E],
    history [B, L, E], mask bool / int [B, L]); `unit.attend(...)` is the fused form that reads the history rows from the
    embedding tables by id.  Kernels Glorot-uniform, biases zero (Keras defaults), built on the first call."""

    UNITS = (80, 40, 1)      # dien/layers.py:37-39

    def __init__(self, generator: Optional[torch.Generator] = None):
        super().__init__()
        self._generator = generator
        self.kernels = nn.ParameterList()
        self.biases = nn.ParameterList()

    def build(self, in_dim: int, device) -> None:
        for u in self.UNITS:
            lim = math.sqrt(6.0 / (in_dim + u))
            self.kernels.append(nn.Parameter(torch.empty(in_dim, u, dtype=torch.float32, device=device).uniform_(-lim, lim, generator=self._generator)))
            self.biases.append(nn.Parameter(torch.zeros(u, dtype=torch.float32, device=device)))
            in_dim = u

    def load_arrays(self, layers, device) -> None:
        """Adopt [(kernel, bias)] x 3 (parity tests: the oracle owns the init)."""
        self.kernels = nn.ParameterList(nn.Parameter(torch.as_tensor(W, dtype=torch.float32).to(device).contiguous()) for W, _ in layers)
        self.biases = nn.ParameterList(nn.Parameter(torch.as_tensor(b, dtype=torch.float32).to(device).contiguous()) for _, b in layers)

    def _params(self, E: int, device):
        if len(self.kernels) == 0:
            self.build(4 * E, device)
        return [t for Wb in zip(self.kernels, self.biases) for t in Wb]

    @staticmethod
    def _int_mask(mask: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if mask is None:
            return None
        return mask if mask.dtype in (torch.int32, torch.int64) else mask.to(torch.int32)

    def forward(self, inputs, mask=None):
        target, history = inputs                                                   # dien/layers.py:46
        if not history.is_cuda:
            raise RuntimeError("recommender_b200 needs CUDA tensors (there is no CPU path)")
        if target.dim() == 3:
            target = target.squeeze(1)
        if mask is None:
            raise ValueError("LocalActivationUnit needs the history mask (dien/model.py:43)")
        spec = dict(mask=self._int_mask(mask).to(torch.int64))
        return _LocalActivationFn.apply(target, history, spec, *self._params(history.shape[-1], history.device))

    def attend(self, target: torch.Tensor, item_emb: Embedding, his_item: torch.Tensor, cat_emb: Optional[Embedding] = None,
               his_cat: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """history_representation [B, E] for history rows [item_emb[his_item] | cat_emb[his_cat]] (dien/model.py:46-49), read
        from the tables by the kernels; mask defaults to his_item != 0 (mask_zero, dien/model.py:43)."""
        E = item_emb.output_dim + (cat_emb.output_dim if cat_emb is not None else 0)
        m = self._int_mask(mask)
        if m is not None and m.dtype != his_item.dtype:
            m = m.to(his_item.dtype)
        spec = dict(item_emb=item_emb, cat_emb=cat_emb, item=his_item, cat=his_cat, mask=m)
        if target.dim() == 3:
            target = target.squeeze(1)
        return _LocalActivationFn.apply(target, None, spec, *self._params(E, his_item.device))


class DIN(nn.Module):
    """dien/model.py:6-20, 36-53 on the CUDA path: item / category tables (mask_zero), target and history lookups, attention
    pooling, `concat([target_embedding, history_representation])`; `head` (any module, e.g. the caller's BatchNorm-MLP)
    maps that [B, 2E] tensor to the probability.  Inputs: the reference's dict ('target_item', 'target_cat' [B, 1];
    'pos_his_item', 'pos_his_cat' [B, L])."""

    def __init__(self, item_vocab_size: int, item_embedding_size: int, cat_vocab_size: int, cat_embedding_size: int,
                 head: Optional[nn.Module] = None, device=None, generator: Optional[torch.Generator] = None):
        super().__init__()
        self.item_embedding = Embedding(item_vocab_size, item_embedding_size, mask_zero=True, device=device, generator=generator)   # :11
        self.cat_embedding = Embedding(cat_vocab_size, cat_embedding_size, mask_zero=True, device=device, generator=generator)     # :12
        # each table is looked up twice per step (target and history): the one-call update concatenates the groups
        self.item_embedding.presort = self.cat_embedding.presort = False
        self.local_activation_unit = LocalActivationUnit(generator=generator)                                                   # :39
        self.head = head

    def compute_flat_embedding(self, inputs: Sequence[torch.Tensor]) -> torch.Tensor:
        item, cat = inputs                                                              # :15
        return torch.cat([self.item_embedding(item), self.cat_embedding(cat)], dim=-1)  # :16-19

    def embed(self, inputs) -> torch.Tensor:
        target_embedding = self.compute_flat_embedding((inputs["target_item"], inputs["target_cat"]))      # :44-45  [B, 1, E]
        target_embedding = target_embedding.squeeze(1)                                                     # :50
        history_representation = self.local_activation_unit.attend(                                        # :43, :46-49
            target_embedding, self.item_embedding, inputs["pos_his_item"], self.cat_embedding, inputs["pos_his_cat"])
        return torch.cat([target_embedding, history_representation], dim=-1)                               # :51

    def forward(self, inputs, training=False, mask=None):
        embedding = self.embed(inputs)
        return embedding if self.head is None else self.head(embedding)                                    # :52
