"""Multi-GPU form of the hot path: sharded embedding tables + all-to-all, data-parallel MLPs.

One process per GPU (torchrun), `torch.distributed` for the plumbing (NCCL on B200s over
NVLink/NVSwitch; gloo in the CPU tests).  The reference itself only replicates the full table on
every GPU (tf.distribute.MirroredStrategy, ctr/train.py:71-84); sharding is what the north star
adds:

  row-wise   owner = row mod G, local row = row div G        (one huge / shared table, config 3)
  table-wise table t lives on rank t mod G                    (26 equal tables, config 2)

Per step and rank (B_local samples, n = B_local*F lookups):
  plan      bucket the lookups by owner (rb_bucket_by_owner; static for table-wise), exchange the
            bucket sizes, all-to-all the local row ids;
  forward   the owner gathers the requested rows from its shard (rb_gather_fwd), an all-to-all
            returns them, and the interaction kernel reads them IN PLACE through the inverse
            permutation (the receive buffer is addressed like a table) — no un-permute pass;
  backward  interaction backward -> dE, permuted into bucket order (rb_gather_fwd with the
            permutation), all-to-all back to the owners, which run the sorted scatter + fused
            optimizer row update on their shard.  Every row has exactly one owner, so duplicate
            rows from different ranks meet in ONE deterministic segmented reduction.
  MLP grads all-reduced (SUM over replicas, the reduction MirroredStrategy applies with
            Reduction.NONE losses — SURVEY A.5/A.7).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import nn

from . import ops
from .layers import MLP, Embedding, GradSource, LookupGroup


@dataclass
class ExchangePlan:
    """Routing of one batch's lookups to their owner ranks."""
    send_rows: torch.Tensor      # int64[n]   local row ids, bucket order (owner 0's first)
    perm: torch.Tensor           # int32[n]   bucket slot -> lookup position
    inv_perm: torch.Tensor       # int32[n]   lookup position -> bucket slot
    send_counts: List[int]       # lookups this rank sends to each owner
    recv_counts: List[int]       # lookups this rank receives from each source
    recv_rows: torch.Tensor      # int64[sum(recv_counts)] local row ids to serve, source-major
    shape: tuple                 # (B_local, F)


def _a2a(out, inp, out_splits, in_splits, group):
    dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
    return out


class _ShardedLookupFn(torch.autograd.Function):
    """rows = all_to_all(gather(shard, recv_rows)); forward returns the receive buffer itself
    ([n, D], bucket order); backward routes the gradient rows back to the owners."""

    @staticmethod
    def forward(ctx, anchor, semb, plan):
        served = ops.gather_fwd(semb.shard.embeddings, plan.recv_rows)                    # [n_recv, D]
        ret = torch.empty(plan.send_rows.numel(), semb.output_dim, dtype=torch.float32, device=served.device)
        _a2a(ret, served, plan.send_counts, plan.recv_counts, semb.group)
        ctx.semb, ctx.plan = semb, plan
        return ret

    @staticmethod
    def backward(ctx, d_ret):
        semb, plan = ctx.semb, ctx.plan
        d_ret = d_ret.contiguous()
        recv = torch.empty(plan.recv_rows.numel(), semb.output_dim, dtype=torch.float32, device=d_ret.device)
        _a2a(recv, d_ret, plan.recv_counts, plan.send_counts, semb.group)
        semb.shard._record(LookupGroup(plan.recv_rows, 1, GradSource.per_position(recv, 1)))
        return None, None, None


class _PermutedInteractFn(torch.autograd.Function):
    """DLRM concat + DotInteraction + tail over rows that sit in bucket order in `ret`: the
    kernel reads row (b,f) at ret[inv_perm[b,f]] (fused-gather form of rb_dot_interaction_fwd), and
    its backward hands d(ret) back in bucket order."""

    @staticmethod
    def forward(ctx, ret, inv_perm, perm, dense_vec, flags):
        dense_vec = dense_vec.contiguous()
        si, sg, tail = flags
        out = ops.dot_interaction_fwd(table=ret, idx=inv_perm, dense_vec=dense_vec, self_interaction=si, skip_gather=sg, tail=tail)
        ctx.save_for_backward(ret, inv_perm, perm, dense_vec)
        ctx.flags = flags
        return out

    @staticmethod
    def backward(ctx, dOut):
        ret, inv_perm, perm, dense_vec = ctx.saved_tensors
        si, sg, tail = ctx.flags
        if dOut.stride(-1) != 1:
            dOut = dOut.contiguous()
        dE, d_dense = ops.dot_interaction_bwd(dOut, table=ret, idx=inv_perm, dense_vec=dense_vec, self_interaction=si,
                                              skip_gather=sg, tail=tail)
        d_ret = ops.gather_fwd(dE.view(-1, dE.shape[-1]), perm)          # natural order -> bucket order
        return d_ret, None, None, d_dense, None


class _UnpermuteFn(torch.autograd.Function):
    """E[b,f,:] = ret[inv_perm[b,f],:] (the un-fused lookup result) and its inverse in backward."""

    @staticmethod
    def forward(ctx, ret, inv_perm, perm):
        ctx.save_for_backward(perm)
        return ops.gather_fwd(ret, inv_perm)

    @staticmethod
    def backward(ctx, dE):
        (perm,) = ctx.saved_tensors
        dE = dE.contiguous()
        return ops.gather_fwd(dE.view(-1, dE.shape[-1]), perm), None, None


class ShardedEmbedding(nn.Module):
    """`Embedding(input_dim, output_dim, num_tables=T)` whose rows are spread over the ranks of
    `group`.  Call surface as layers.Embedding (`__call__`, `interact`); the local shard is a
    layers.Embedding, so the optimizers drive it unchanged."""

    def __init__(self, input_dim: int, output_dim: int, *, num_tables: int = 1, sharding: str = "row", hash_mod: int = 0,
                 group=None, device=None, generator: Optional[torch.Generator] = None):
        super().__init__()
        if sharding not in ("row", "table"):
            raise ValueError(sharding)
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.input_dim, self.output_dim, self.num_tables = int(input_dim), int(output_dim), int(num_tables)
        self.sharding, self.hash_mod = sharding, int(hash_mod)
        G, T = self.world, self.num_tables
        if sharding == "table":
            if T < G:
                raise ValueError(f"table-wise sharding needs at least as many tables ({T}) as ranks ({G})")
            self.table_owner = [t % G for t in range(T)]
            self.owned = [t for t in range(T) if t % G == self.rank]
            local_rows = self.input_dim * len(self.owned)
        else:
            total = self.input_dim * T
            local_rows = (total - self.rank + G - 1) // G                  # rows r with r mod G == rank
        self.shard = Embedding(max(local_rows, 1), output_dim, device=device, generator=generator)
        dev = self.shard.embeddings.device
        self._row_offset = (torch.arange(T, dtype=torch.int64, device=dev) * self.input_dim) if T > 1 else None
        self._static = {}

    # ---- shard <-> full table (tests, checkpoints) ---------------------------------------------------------
    def load_full_table(self, full: torch.Tensor) -> None:
        """Adopt this rank's rows of a full [input_dim*num_tables, D] table."""
        full = torch.as_tensor(full, dtype=torch.float32)
        if self.sharding == "row":
            mine = full[self.rank::self.world]
        else:
            V = self.input_dim
            mine = torch.cat([full[t * V:(t + 1) * V] for t in self.owned], dim=0) if self.owned else full[:1]
        self.shard.embeddings.copy_(mine.to(self.shard.embeddings.device))

    def full_row_ids(self) -> torch.Tensor:
        """Global row id of every local shard row (inverse of the sharding map)."""
        n = self.shard.embeddings.shape[0]
        dev = self.shard.embeddings.device
        if self.sharding == "row":
            return torch.arange(n, device=dev, dtype=torch.int64) * self.world + self.rank
        V = self.input_dim
        return torch.cat([torch.arange(V, device=dev, dtype=torch.int64) + t * V for t in self.owned])

    # ---- routing ------------------------------------------------------------------------------------------------
    def _table_wise_static(self, B: int, F: int, device):
        key = (B, F)
        if key not in self._static:
            if F != self.num_tables:
                raise ValueError(f"table-wise sharding takes [B, {self.num_tables}] indices")
            G = self.world
            cols = [[t for t in range(F) if self.table_owner[t] == q] for q in range(G)]
            # bucket order: for owner q, sample-major then the owner's tables in ascending order
            pos = torch.arange(B * F, dtype=torch.int64).reshape(B, F)
            perm = torch.cat([pos[:, c].reshape(-1) for c in cols])
            inv = torch.empty_like(perm)
            inv[perm] = torch.arange(B * F, dtype=torch.int64)
            # local row offset of table t on its owner: index among the owner's tables * input_dim
            local_off = torch.tensor([cols[self.table_owner[t]].index(t) * self.input_dim for t in range(F)], dtype=torch.int64)
            self._static[key] = (perm.to(torch.int32).to(device), inv.to(torch.int32).to(device), local_off.to(device),
                                 [B * len(c) for c in cols])
        return self._static[key]

    def plan(self, idx: torch.Tensor) -> ExchangePlan:
        if idx.dim() != 2:
            raise ValueError("sharded lookups take [B, F] indices")
        B, F = idx.shape
        G = self.world
        dev = idx.device
        if self.sharding == "row":
            off = self._row_offset if self.num_tables > 1 else None
            if off is not None and F != self.num_tables:
                raise ValueError(f"a {self.num_tables}-table embedding takes [B, {self.num_tables}] indices")
            send_rows, perm, inv_perm, counts = ops.bucket_by_owner(idx, G, L=F, field_row_offset=off, hash_mod=self.hash_mod)
            recv_counts_t = torch.empty_like(counts)
            dist.all_to_all_single(recv_counts_t, counts, group=self.group)
            both = torch.stack([counts, recv_counts_t]).cpu()                # the one host sync of the step
            send_counts, recv_counts = both[0].tolist(), both[1].tolist()
        else:
            perm, inv_perm, local_off, send_counts = self._table_wise_static(B, F, dev)
            ids = idx
            if self.hash_mod:
                ids = ops.hash_ids(idx, self.hash_mod)[0]
            send_rows = (ids.to(torch.int64) + local_off[None]).reshape(-1).index_select(0, perm.to(torch.int64))
            # every source sends B_src * |my tables| lookups; equal local batches are required
            recv_counts = [B * len(self.owned)] * G
        recv_rows = torch.empty(sum(recv_counts), dtype=torch.int64, device=dev)
        _a2a(recv_rows, send_rows, recv_counts, send_counts, self.group)
        return ExchangePlan(send_rows, perm, inv_perm.reshape(B, F), send_counts, recv_counts, recv_rows, (B, F))

    # ---- call surface --------------------------------------------------------------------------------------------
    def lookup_rows(self, plan: ExchangePlan) -> torch.Tensor:
        return _ShardedLookupFn.apply(self.shard._anchor, self, plan)

    def forward(self, idx: torch.Tensor, plan: Optional[ExchangePlan] = None) -> torch.Tensor:
        plan = plan or self.plan(idx)
        return _UnpermuteFn.apply(self.lookup_rows(plan), plan.inv_perm, plan.perm)

    def interact(self, idx: torch.Tensor, dense_vec: torch.Tensor, self_interaction=False, skip_gather=True, tail=True,
                 plan: Optional[ExchangePlan] = None) -> torch.Tensor:
        plan = plan or self.plan(idx)
        return _PermutedInteractFn.apply(self.lookup_rows(plan), plan.inv_perm, plan.perm, dense_vec,
                                         (self_interaction, skip_gather, tail))


class ShardedDLRM(nn.Module):
    """ctr/model.py:34-58 with the table sharded over the process group and the MLPs replicated."""

    def __init__(self, bottom_mlp_units: Sequence[int], top_mlp_units: Sequence[int], embedding_size: int, vocab_size: int,
                 num_cat_fea: int, num_int_fea: int, *, num_tables: int = 1, sharding: str = "row", group=None, device=None,
                 compute_dtype: Optional[torch.dtype] = None, generator: Optional[torch.Generator] = None):
        super().__init__()
        if bottom_mlp_units[-1] != embedding_size:
            raise ValueError("bottom_mlp_units[-1] must equal embedding_size")       # ctr/model.py:52,55
        self.group = group
        self.bottom_mlp = MLP(bottom_mlp_units, "relu", compute_dtype=compute_dtype, generator=generator)
        self.top_mlp = MLP(top_mlp_units, "sigmoid", compute_dtype=compute_dtype, generator=generator)
        self.embedding_layer = ShardedEmbedding(vocab_size, embedding_size, num_tables=num_tables, sharding=sharding,
                                                group=group, device=device, generator=generator)
        self.num_cat_fea, self.num_int_fea, self.embedding_size = num_cat_fea, num_int_fea, embedding_size
        self._synced = False

    def sync_dense_parameters(self) -> None:
        """Replicas start from rank 0's MLP weights (MirroredStrategy mirrors variables)."""
        for p in self.parameters():
            dist.broadcast(p.data, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            if hasattr(p, "_rb_shadow"):
                del p._rb_shadow              # .data writes do not bump the version the bf16 shadow is keyed on
        self._synced = True

    def forward(self, inputs, training=None, mask=None):
        int_features = inputs["int_features"].reshape(-1, self.num_int_fea)
        cat_features = inputs["cat_features"].reshape(-1, self.num_cat_fea)
        if not self._synced:         # build both towers, then adopt rank 0's weights BEFORE anything is computed from them
            if len(self.bottom_mlp.kernels) == 0:
                self.bottom_mlp.build(self.num_int_fea, int_features.device)
            if len(self.top_mlp.kernels) == 0:
                self.top_mlp.build((self.num_cat_fea + 1) ** 2 + self.embedding_size, int_features.device)
            self.sync_dense_parameters()
        bmlp_output = self.bottom_mlp(int_features)
        tmlp_input = self.embedding_layer.interact(cat_features, bmlp_output, False, True, True, plan=inputs.get("plan"))
        output = self.top_mlp(tmlp_input)
        return output.squeeze(1)

    def reduce_dense_grads(self) -> None:
        """SUM over replicas of the MLP gradients in one flat all-reduce (called by the optimizers)."""
        params = [p for p in self.parameters() if p.grad is not None]
        if not params or dist.get_world_size(self.group) == 1:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        o = 0
        for p in params:
            n = p.numel()
            p.grad.copy_(flat[o:o + n].view_as(p.grad))
            o += n
