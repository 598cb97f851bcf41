"""DeepFM and DLRM with the reference's call surface (ctr/model.py), on the CUDA hot path.

Constructor arguments, the input dict ({'cat_features', 'int_features'}) and the f32[B] output
are those of ctr/model.py:6-58.  `fused=True` (default) routes the lookup + interaction through
the single-pass kernels; `fused=False` replays the reference's op sequence layer by layer
(Embedding -> concat -> DotInteraction -> concat) on the same kernels, un-fused.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import nn

from . import ops
from .layers import MLP, DotInteraction, Embedding


class DeepFM(nn.Module):
    """ctr/model.py:6-31.  One shared table; FM second order (no first-order term, no bias) + an MLP
    on [flatten(E) || int_features]; sigmoid(fm + mlp)."""

    def __init__(self, embedding_size: int, vocab_size: int, num_int_fea: int, num_cat_fea: int, mlp_units: Sequence[int],
                 *, num_tables: int = 1, fused: bool = True, device=None, compute_dtype: Optional[torch.dtype] = None,
                 generator: Optional[torch.Generator] = None, collapse_linear: bool = False):
        super().__init__()
        self.embedding_layer = Embedding(vocab_size, embedding_size, num_tables=num_tables, device=device, generator=generator)  # :10
        self.mlp = MLP(mlp_units, None, compute_dtype=compute_dtype, generator=generator, collapse_linear=collapse_linear)     # :11
        self.num_int_fea, self.num_cat_fea, self.fused = num_int_fea, num_cat_fea, fused

    def logits(self, inputs):
        int_features = inputs["int_features"].reshape(-1, self.num_int_fea)            # :17
        cat_features = inputs["cat_features"].reshape(-1, self.num_cat_fea)            # :18
        if self.fused and self.mlp.compute_dtype == torch.bfloat16 and cat_features.is_cuda:
            # :19-27 with the deep input written by the gather itself as the MLP's bf16 K operand: no E tensor, no reshape,
            # no concat, no pad copy
            D = self.embedding_layer.output_dim
            if len(self.mlp.kernels) == 0:
                self.mlp.build(self.num_cat_fea * D + self.num_int_fea, cat_features.device)
            Kp = self.mlp.padded_in_dim()
            deep_input, interaction = self.embedding_layer.lookup_fm_deep(cat_features, int_features, Kp)       # :19-26
            dense_output = self.mlp(deep_input, ones_col=Kp > self.mlp.in_dim)                                    # :27
            return interaction + dense_output.squeeze(1)                                                          # :28-29
        if self.fused:
            cat_embedding, interaction = self.embedding_layer.lookup_fm(cat_features)  # :19-23 in one pass
        else:
            cat_embedding = self.embedding_layer(cat_features)                          # :19
            sum_square = torch.square(cat_embedding.sum(dim=1))                         # :21
            square_sum = torch.square(cat_embedding).sum(dim=1)                         # :22
            interaction = 0.5 * (sum_square - square_sum).sum(dim=1)                    # :23
        deep_cat_input = cat_embedding.reshape(-1, self.num_cat_fea * cat_embedding.shape[2])   # :25
        deep_input = torch.cat([deep_cat_input, int_features], dim=1)                   # :26
        dense_output = self.mlp(deep_input)                                             # :27
        return interaction + dense_output.squeeze(1)                                    # :28-29

    def forward(self, inputs, training=None, mask=None):
        return torch.sigmoid(self.logits(inputs))                                       # :30


class DLRM(nn.Module):
    """ctr/model.py:34-58.  Shared table, bottom MLP (last activation relu) whose output joins the 26
    embeddings as feature 27, DotInteraction(False, True), [interaction || bottom] -> top MLP (sigmoid)."""

    def __init__(self, bottom_mlp_units: Sequence[int], top_mlp_units: Sequence[int], embedding_size: int, vocab_size: int,
                 num_cat_fea: int, num_int_fea: int, *, num_tables: int = 1, fused: bool = True, device=None,
                 compute_dtype: Optional[torch.dtype] = None, generator: Optional[torch.Generator] = None,
                 collapse_linear: bool = False, table_rows: Optional[Sequence[int]] = None):
        super().__init__()
        if bottom_mlp_units[-1] != embedding_size:
            # ctr/model.py:52,55: the concat and the shape-asserting reshape need equal widths
            raise ValueError("bottom_mlp_units[-1] must equal embedding_size")
        # collapse_linear (opt-in): the hidden layers are linear (ctr/layers.py:8), so each tower is one affine map
        # followed by its last activation; layers._CollapsedAffineFn evaluates it that way, gradients per layer intact
        self.bottom_mlp = MLP(bottom_mlp_units, "relu", compute_dtype=compute_dtype, generator=generator,
                              collapse_linear=collapse_linear)                                                 # :38
        self.top_mlp = MLP(top_mlp_units, "sigmoid", compute_dtype=compute_dtype, generator=generator,
                           collapse_linear=collapse_linear)                                                    # :39
        # table_rows: one row count per field (BASELINE config 3's capped cardinalities) instead of num_tables x vocab_size
        self.embedding_layer = Embedding(vocab_size, embedding_size, num_tables=num_tables, device=device, generator=generator,
                                         table_rows=table_rows)                                                             # :42
        self.interaction = DotInteraction(False, True)                                                          # :43
        if fused:
            # the top tower's weight gradients wait until the interaction backward is queued (MLP.flush_wgrad)
            self.top_mlp.defer_wgrad = True
            self.embedding_layer.after_grad_hooks.append(self.top_mlp.flush_wgrad)
        self.num_cat_fea, self.num_int_fea, self.embedding_size, self.fused = num_cat_fea, num_int_fea, embedding_size, fused

    def forward_bce(self, inputs, label: torch.Tensor) -> torch.Tensor:
        """bce_clipped(self(inputs), label) — ctr/model.py:45-58 + the loss of ctr/train.py:85-87 — with the last Dense(1, sigmoid),
        the loss and the head's backward in ONE kernel (rb_dense_head_bce) when the top tower runs on the bf16 tcgen05 path; the
        probabilities are left in `self.top_mlp.last_prob`.  Same numbers as the two-call form (the loss up to the order its terms
        are added); for training steps that call `.backward()` on the returned loss itself."""
        if label.dtype not in (torch.float32, torch.int64):
            label = label.to(torch.float32)
        return self.forward(inputs, _bce_label=label)

    def forward(self, inputs, training=None, mask=None, _bce_label=None):
        int_features = inputs["int_features"].reshape(-1, self.num_int_fea)            # :47
        cat_features = inputs["cat_features"].reshape(-1, self.num_cat_fea)            # :48
        if _bce_label is not None and not (self.fused and self.top_mlp.compute_dtype == torch.bfloat16):
            return bce_clipped(self.forward(inputs), _bce_label)
        if self.fused and torch.is_grad_enabled():
            # the backward's sort of (row, position) pairs depends on the ids only: start it under the bottom MLP
            cat_features = self.embedding_layer.start_presort(cat_features)
        bmlp_output = self.bottom_mlp(int_features)                                     # :50
        if self.fused:
            width = (self.num_cat_fea + 1) ** 2 + self.embedding_size                                     # :55
            if self.top_mlp.compute_dtype == torch.bfloat16:
                # the fused kernel emits the top MLP's K operand directly: bf16, zero-padded to 8 columns
                if len(self.top_mlp.kernels) == 0:
                    self.top_mlp.build(width, bmlp_output.device)
                ones = width % 8 != 0          # a spare pad column exists: it carries 1.0 so that db comes with the dW GEMM
                tmlp_input = self.embedding_layer.interact(cat_features, bmlp_output, False, True, True,
                                                           out_dtype=torch.bfloat16, pad_to=8, ones_col=ones)   # :49,:51-55
                if _bce_label is not None:
                    if self.top_mlp.can_fuse_bce(tmlp_input):
                        return self.top_mlp(tmlp_input, ones_col=ones, bce_label=_bce_label).reshape(())  # :56-57 + the loss
                    return bce_clipped(self.top_mlp(tmlp_input, ones_col=ones).squeeze(1), _bce_label)
                return self.top_mlp(tmlp_input, ones_col=ones).squeeze(1)                               # :56-57
            tmlp_input = self.embedding_layer.interact(cat_features, bmlp_output, False, True, True)   # :49,:51-55
        else:
            cat_embedding = self.embedding_layer(cat_features)                          # :49
            interaction_input = torch.cat([cat_embedding, bmlp_output.unsqueeze(1)], dim=1)            # :51-52
            interaction_output = self.interaction(interaction_input)                    # :53
            tmlp_input = torch.cat([interaction_output, bmlp_output], dim=1)            # :54
        tmlp_input = tmlp_input.reshape(-1, (self.num_cat_fea + 1) ** 2 + self.embedding_size)          # :55
        output = self.top_mlp(tmlp_input)                                               # :56
        return output.squeeze(1)                                                        # :57


class _BCEClippedFn(torch.autograd.Function):
    """Loss and d loss / d prob from one kernel pair (rb_bce_clipped) instead of ~25 elementwise launches."""

    @staticmethod
    def forward(ctx, prob, label):
        loss, dprob = ops.bce_clipped(prob, label, want_grad=True)
        ctx.save_for_backward(dprob)
        ctx.shape = prob.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (dprob,) = ctx.saved_tensors
        return (dprob * g).reshape(ctx.shape), None


def bce_clipped(prob: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """Keras binary_crossentropy on probabilities, batch mean — what compile(loss=BinaryCrossentropy)
    evaluates for DLRM (ctr/train.py:85-87)."""
    if label.dtype not in (torch.float32, torch.int64):
        label = label.to(torch.float32)
    return _BCEClippedFn.apply(prob.float(), label)          # rb_bce_clipped: loss and d loss / d prob in one pass


def bce_logits(logit: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """sigmoid_cross_entropy_with_logits, batch mean — the form Keras recovers for DeepFM whose last op
    is Sigmoid (ctr/model.py:30)."""
    y = label.to(logit.dtype)
    return (logit.clamp(min=0) - logit * y + torch.log1p(torch.exp(-logit.abs()))).mean()
