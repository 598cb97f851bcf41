"""Mirrors of the Keras optimizers the reference builds (`Adam()` at ctr/train.py:80,84).

One optimizer object drives both kinds of variables of a model:
  * dense parameters (the MLPs) — the Keras formula (`_resource_apply_dense`; its epsilon
    placement differs from torch.optim.Adam, which is why torch's own Adam is not used) for ALL tensors in one launch
    (rb_dense_opt_step);
  * embedding tables — ONE fused CUDA call per table (`Embedding.apply_pending`): sort of the
    (row, position) pairs, duplicate-row sum, optimizer row update (A.1-A.4).
"""
from __future__ import annotations

import torch

from . import ops


def _split(model_or_vars):
    if isinstance(model_or_vars, torch.nn.Module):
        # layers.Embedding and the sharded embeddings (sharded.py's local shard, p2p.P2PShardedEmbedding)
        embs = [m for m in model_or_vars.modules() if hasattr(m, "apply_pending")]
        dense = [p for p in model_or_vars.parameters() if p.requires_grad]
        return embs, dense
    embs, dense = [], []
    for v in model_or_vars:
        (embs if hasattr(v, "apply_pending") else dense).append(v)
    return embs, dense


class _Optimizer:
    sparse_kind = "sgd"

    def __init__(self):
        self.iterations = 0
        self._state = {}
        self._prepared = False      # prepare_step() already advanced `iterations` for the step being applied

    def prepare_step(self) -> None:
        """Host half of a step whose device half is a captured CUDA graph (graph.GraphedTrainStep): advances
        `iterations` and refreshes every step-dependent scalar the kernels read from device memory.
        Call it before each replay; apply_gradients() then must not advance the counter again."""
        self.iterations += 1
        self._prepared = True
        self._refresh_device_scalars()

    def _refresh_device_scalars(self) -> None:
        pass

    def fuse_sparse_updates(self, model, enable: bool = True) -> int:
        """Opt in to the fused row update: from now on the backward of a table's fused lookup (DLRM's lookup + interaction)
        applies THIS optimizer's row update to the rows the step touches exactly once, where their gradient rows are produced
        (rb_dot_interaction_bwd_update); apply_gradients() then covers the remaining rows.  Tables, states and results are
        bit-identical to the unfused order — but the table changes during loss.backward(), so only a caller that always
        follows backward() with apply_gradients() (a training step) may arm it.  Returns the number of tables armed."""
        n = 0
        for m in model.modules():
            if hasattr(m, "fused_optimizer") and hasattr(m, "_fused_update_args"):
                m.fused_optimizer = self if (enable and self.sparse_kind in ("sgd", "adagrad", "adam_lazy")) else None
                n += m.fused_optimizer is not None
        return n

    def _sparse_kwargs(self):
        raise NotImplementedError

    def _dense_step(self, params, grads, step):
        raise NotImplementedError

    @torch.no_grad()
    def apply_gradients(self, model_or_vars) -> None:
        """One optimizer step over every dense parameter with a .grad and every Embedding with
        recorded lookups; then clears them (the reference's `apply_gradients(zip(grads, vars))`,
        dien/train.py:22)."""
        embs, dense = _split(model_or_vars)
        if hasattr(model_or_vars, "reduce_dense_grads"):
            model_or_vars.reduce_dense_grads()      # data-parallel replicas (sharded.ShardedDLRM, p2p.P2PShardedDLRM)
        step = self.iterations if self._prepared else self.iterations + 1
        # sparse tables first: their row updates run on side streams (layers.Embedding, p2p.P2PShardedEmbedding) and overlap
        # the replicas' all-reduce and the dense step below; join() brings the streams back together
        for e in embs:
            e.apply_pending(self.sparse_kind, step, **self._sparse_kwargs())
        if hasattr(model_or_vars, "wait_dense_grads"):
            model_or_vars.wait_dense_grads()
        params = [p for p in dense if p.grad is not None]
        if params:
            self._dense_step(params, [p.grad for p in params], step)
            if hasattr(model_or_vars, "reset_dense_grads"):
                model_or_vars.reset_dense_grads()      # gradients are views of one flat buffer: zero it, keep the views
            else:
                for p in params:
                    p.grad = None
        for e in embs:
            if hasattr(e, "join"):
                e.join()
        self.iterations = step
        if not torch.cuda.is_available() or not torch.cuda.is_current_stream_capturing():
            self._prepared = False      # a captured body is replayed: every replay is preceded by prepare_step()


class Adam(_Optimizer):
    """tf.keras.optimizers.Adam(learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7).

    sparse='lazy' (default) updates only the rows a step touches; sparse='tf_dense' reproduces
    Keras' `_resource_apply_sparse` exactly (m, v decay and var moves on EVERY row each step).
    Both agree on step 1 from zero state."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, sparse="lazy"):
        super().__init__()
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.sparse_kind = {"lazy": "adam_lazy", "tf_dense": "adam_tf_dense"}[sparse]
        self._alpha_dev = None      # device f32[1] holding alpha_t of the current step (graph replays)
        self._alpha_host = None

    def enable_device_scalars(self, device) -> None:
        """alpha_t lives in device memory from now on (rb_opt_params.alpha_t_dev), refreshed by prepare_step()."""
        if self._alpha_dev is None:
            self._alpha_dev = torch.zeros(1, dtype=torch.float32, device=device)
            # A ring of pinned slots, one per step in flight: the copy below is asynchronous and a replay loop that does
            # not read anything back runs many steps ahead of the GPU; a single slot would be overwritten with a later
            # step's alpha before the DMA of an earlier step has read it.
            self._alpha_host = torch.zeros(self._ALPHA_SLOTS, dtype=torch.float32).pin_memory()

    _ALPHA_SLOTS = 4096      # far beyond the launches the driver queues ahead

    def _refresh_device_scalars(self) -> None:
        if self._alpha_dev is not None:
            slot = self.iterations % self._ALPHA_SLOTS
            self._alpha_host[slot] = ops.adam_alpha_t(self.learning_rate, self.beta_1, self.beta_2, self.iterations)
            self._alpha_dev.copy_(self._alpha_host[slot:slot + 1], non_blocking=True)

    def _sparse_kwargs(self):
        return dict(lr=self.learning_rate, beta_1=self.beta_1, beta_2=self.beta_2, epsilon=self.epsilon,
                    alpha_dev=self._alpha_dev if self._prepared else None)

    def _dense_step(self, params, grads, step):
        ms, vs = [], []
        for p in params:
            st = self._state.get(p)
            if st is None:
                st = self._state[p] = (torch.zeros_like(p), torch.zeros_like(p))
            ms.append(st[0])
            vs.append(st[1])
        # one launch for every dense tensor (rb_dense_opt_step)
        ops.dense_opt_step(params, grads, ms, vs, optimizer="adam_lazy", step=step, lr=self.learning_rate,
                           beta_1=self.beta_1, beta_2=self.beta_2, epsilon=self.epsilon,
                           alpha_dev=self._alpha_dev if self._prepared else None)


class Adagrad(_Optimizer):
    """tf.keras.optimizers.Adagrad(learning_rate=1e-3, initial_accumulator_value=0.1, epsilon=1e-7)
    (named by the north star)."""
    sparse_kind = "adagrad"

    def __init__(self, learning_rate=1e-3, initial_accumulator_value=0.1, epsilon=1e-7):
        super().__init__()
        self.learning_rate, self.initial_accumulator_value, self.epsilon = learning_rate, initial_accumulator_value, epsilon

    def _sparse_kwargs(self):
        return dict(lr=self.learning_rate, epsilon=self.epsilon, initial_accumulator_value=self.initial_accumulator_value)

    def _dense_step(self, params, grads, step):
        accs = []
        for p in params:
            st = self._state.get(p)
            if st is None:
                st = self._state[p] = torch.full_like(p, self.initial_accumulator_value)
            accs.append(st)
        ops.dense_opt_step(params, grads, accs, None, optimizer="adagrad", lr=self.learning_rate, epsilon=self.epsilon)


class SGD(_Optimizer):
    """tf.keras.optimizers.SGD(learning_rate) — the commented-out option at ctr/train.py:79."""
    sparse_kind = "sgd"

    def __init__(self, learning_rate=1e-2):
        super().__init__()
        self.learning_rate = learning_rate

    def _sparse_kwargs(self):
        return dict(lr=self.learning_rate)

    def _dense_step(self, params, grads, step):
        ops.dense_opt_step(params, grads, None, None, optimizer="sgd", lr=self.learning_rate)
