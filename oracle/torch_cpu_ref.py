"""The reference's DLRM training step restated on torch-CPU ops (TEST / BASELINE INFRASTRUCTURE).

Why this exists.  The reference's CPU path is TensorFlow-CPU executing ctr/model.py + Keras Adam
(ctr/train.py:71-97).  TensorFlow is absent here and un-installable, so `bench.py --impl reference`
and the `cpu_baseline` leg time THIS restatement: the same op sequence TF would run — gather,
concat, batched matmul, mask/select, MLPs, IndexedSlices dedup (unique + segment sum), Adam —
each as the multi-threaded torch-CPU kernel that corresponds to TF's Eigen kernel, on all host
cores.  It is checked against the numpy oracle in tests/test_torch_cpu_ref.py.  Only bench.py and
tests/ may import it (see oracle/__init__.py).
"""
from __future__ import annotations

import math

import torch


def glorot(rng: torch.Generator, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(fan_in, fan_out, generator=rng) * 2 - 1) * lim


class DLRMRef:
    """ctr/model.py:34-58 + Keras Adam on torch-CPU tensors, hand-derived backward."""

    def __init__(self, bottom_units, top_units, D, vocab_size, num_cat=26, num_int=13, num_tables=1, seed=4,
                 adam="tf_dense"):
        g = torch.Generator().manual_seed(seed)
        rows = vocab_size * num_tables
        self.table = (torch.rand(rows, D, generator=g) - 0.5) * 0.1           # U(-0.05, 0.05)
        self.offsets = torch.arange(num_tables, dtype=torch.int64) * vocab_size if num_tables > 1 else None
        self.D, self.F = D, num_cat
        self.bottom, d = [], num_int
        for u in bottom_units:
            self.bottom.append([glorot(g, d, u), torch.zeros(u)])
            d = u
        self.top, d = [], (num_cat + 1) ** 2 + D
        for u in top_units:
            self.top.append([glorot(g, d, u), torch.zeros(u)])
            d = u
        self.m = torch.zeros_like(self.table)
        self.v = torch.zeros_like(self.table)
        self.dense_state = {}
        self.adam = adam
        self.step = 0
        Fp = num_cat + 1
        self.keep = ~torch.tril(torch.ones(Fp, Fp, dtype=torch.bool))         # ctr/layers.py:32-33

    @staticmethod
    def _mlp_fwd(x, layers, act):
        acts = [x]
        for i, (W, b) in enumerate(layers):
            x = torch.addmm(b, x, W)                                          # hidden layers linear (ctr/layers.py:8)
            if i == len(layers) - 1:
                x = torch.relu(x) if act == "relu" else torch.sigmoid(x)
            acts.append(x)
        return x, acts

    @staticmethod
    def _mlp_bwd(dy, acts, layers, act):
        y = acts[-1]
        dy = dy * (y > 0) if act == "relu" else dy * y * (1 - y)
        grads = []
        for i in range(len(layers) - 1, -1, -1):
            grads.append((acts[i].t() @ dy, dy.sum(0)))
            dy = dy @ layers[i][0].t()
        return dy, grads[::-1]

    def train_step(self, cat, dense_x, label):
        F, D = self.F, self.D
        B = cat.shape[0]
        rows = cat if self.offsets is None else cat + self.offsets[None]
        flat = rows.reshape(-1)
        E = self.table.index_select(0, flat).reshape(B, F, D)                 # ctr/model.py:49
        bmlp, bacts = self._mlp_fwd(dense_x, self.bottom, "relu")             # :50
        X = torch.cat([E, bmlp[:, None, :]], dim=1)                           # :51-52
        Z = torch.bmm(X, X.transpose(1, 2))                                   # ctr/layers.py:25
        inter = torch.where(self.keep[None], Z, torch.zeros((), dtype=Z.dtype)).reshape(B, -1)   # :36-38
        tin = torch.cat([inter, bmlp], dim=1)                                 # ctr/model.py:54
        out, tacts = self._mlp_fwd(tin, self.top, "sigmoid")                  # :56
        prob = out[:, 0]
        eps = 1e-7                                                            # SURVEY A.5 (probability form)
        y = label.float()
        p = prob.clamp(eps, 1 - eps)
        loss = (-(y * torch.log(p + eps) + (1 - y) * torch.log(1 - p + eps))).mean()
        inside = (prob >= eps) & (prob <= 1 - eps)
        dprob = ((-(y / (p + eps)) + (1 - y) / (1 - p + eps)) * inside) / B
        dtin, tgrads = self._mlp_bwd(dprob[:, None], tacts, self.top, "sigmoid")
        Fp = F + 1
        G = torch.where(self.keep[None], dtin[:, : Fp * Fp].reshape(B, Fp, Fp), torch.zeros((), dtype=Z.dtype))
        dX = torch.bmm(G + G.transpose(1, 2), X)
        dbmlp = dX[:, F] + dtin[:, Fp * Fp:]
        _, bgrads = self._mlp_bwd(dbmlp, bacts, self.bottom, "relu")
        self.step += 1
        t = self.step
        b1, b2, lr, e = 0.9, 0.999, 1e-3, 1e-7
        alpha = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        # IndexedSlices -> _deduplicate_indexed_slices (A.1, A.2)
        uniq, inverse = torch.unique(flat, return_inverse=True)
        g = torch.zeros(uniq.numel(), D).index_add_(0, inverse, dX[:, :F].reshape(-1, D))
        if self.adam == "tf_dense":                                           # A.3: every row, every step
            self.m.mul_(b1)
            self.m.index_add_(0, uniq, g, alpha=1 - b1)
            self.v.mul_(b2)
            self.v.index_add_(0, uniq, g * g, alpha=1 - b2)
            self.table.addcdiv_(self.m, self.v.sqrt().add_(e), value=-alpha)
        else:                                                                 # lazy: touched rows only
            m = self.m.index_select(0, uniq).mul_(b1).add_(g, alpha=1 - b1)
            v = self.v.index_select(0, uniq).mul_(b2).addcmul_(g, g, value=1 - b2)
            self.m.index_copy_(0, uniq, m)
            self.v.index_copy_(0, uniq, v)
            self.table.index_copy_(0, uniq, self.table.index_select(0, uniq).addcdiv_(m, v.sqrt().add_(e), value=-alpha))
        for name, layers, grads in (("b", self.bottom, bgrads), ("t", self.top, tgrads)):
            for i, (layer, gr) in enumerate(zip(layers, grads)):
                for j in (0, 1):
                    st = self.dense_state.setdefault((name, i, j), [torch.zeros_like(layer[j]), torch.zeros_like(layer[j])])
                    st[0].mul_(b1).add_(gr[j], alpha=1 - b1)
                    st[1].mul_(b2).addcmul_(gr[j], gr[j], value=1 - b2)
                    layer[j].addcdiv_(st[0], st[1].sqrt().add_(e), value=-alpha)
        return float(loss), prob
