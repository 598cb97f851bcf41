"""Row-wise sharded embedding over NVLink/NVSwitch PEER MEMORY: the gather is the collective.

One process per GPU.  Row r of the (T-table) embedding lives on rank r mod G at local row r div G.
Instead of exchanging rows and gradients with all-to-alls (sharded.py, the NCCL path), every rank
maps the other ranks' buffers into its address space (CUDA IPC) and the kernels read them directly:

  forward   rb_dot_interaction_fwd_sharded: the interaction kernel's cp.async ring pulls each sample's
            26 rows from whichever GPU owns them (~(G-1)/G of the reads cross NVLink) while the MMA of
            the previous sample runs — no all-to-all, no serve-side gather, no staging buffer;
  backward  rb_dot_interaction_bwd_sharded writes dE[B_local, F, D] into this rank's shared buffer; the
            OWNER of a row then pulls the gradient rows of all its lookups — from all ranks — inside
            the segmented reduction (rb_sparse_bwd_apply_p2p), sums duplicates in one deterministic
            order and applies the fused optimizer update to its shard;
  routing   rb_bucket_by_owner publishes (local row, position) per owner in peer memory; the owner
            collects its slices (rb_p2p_collect_keys) into a STATIC-capacity pair list and radix-sorts
            it on a side stream during the forward.

Only two tiny all-reduces per step order the ranks (after the bucket arrays are published; after all
backward kernels finished, before any shard is updated), so every shape is static and the whole
sharded step can be captured as one CUDA graph.  Dense MLPs are data-parallel (sharded.ShardedDLRM).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import nn

from . import _lib, ops
from ._lib import check, lib
from .layers import MLP


class SharedBuffer:
    """Device memory allocated by the library (cudaMalloc) with its IPC handle; `.tensor` is a zero-copy torch view."""

    def __init__(self, shape: Sequence[int], dtype: torch.dtype, device: torch.device):
        self.shape, self.dtype, self.device = tuple(int(s) for s in shape), dtype, device
        nbytes = max(1, int(torch.tensor([], dtype=dtype).element_size()))
        for s in self.shape:
            nbytes *= s
        self.nbytes = max(nbytes, 256)
        ptr = C.c_void_p()
        handle = (C.c_ubyte * _lib.RB_IPC_HANDLE_BYTES)()
        with torch.cuda.device(device):
            check(lib.rb_shared_alloc(self.nbytes, C.byref(ptr), handle), "rb_shared_alloc")
        self.ptr, self.handle = int(ptr.value), bytes(handle)
        typestr = {torch.float32: "<f4", torch.int64: "<i8", torch.int32: "<i4", torch.uint8: "|u1", torch.bfloat16: "<i2"}[dtype]
        self.__cuda_array_interface__ = dict(shape=self.shape, typestr=typestr, data=(self.ptr, False), version=3, strides=None)
        self.tensor = torch.as_tensor(self, device=device)
        if dtype == torch.bfloat16:       # the array interface has no bf16: carried as int16, viewed back
            self.tensor = self.tensor.view(torch.bfloat16)
        assert self.tensor.data_ptr() == self.ptr

    def free(self):
        if self.ptr:
            self.tensor = None
            check(lib.rb_shared_free(self.ptr), "rb_shared_free")
            self.ptr = 0


class PeerLink:
    """How the ranks allocate peer-readable buffers and order themselves.  `alloc(name, shape, dtype)` is a
    collective: every rank allocates a buffer of the SAME shape and gets back (its tensor, the device
    pointers of all ranks' buffers — possibly as a callable resolved later); `barrier()` is a
    stream-ordered rendezvous of all ranks."""

    world: int
    rank: int

    def alloc(self, name: str, shape: Sequence[int], dtype: torch.dtype):
        raise NotImplementedError

    def barrier(self, channel: int = 0) -> None:
        raise NotImplementedError


class DistPeerLink(PeerLink):
    """Real ranks (one process per GPU).  Buffers are CUDA virtual-memory allocations shared through
    torch.distributed's symmetric-memory rendezvous (2 MB pages, file-descriptor exchange): a table shard of
    several GB mapped with legacy cudaIpc handles reads ~20x slower under random access (measured: 27 GB/s
    against 490 GB/s).  `legacy_ipc=True` keeps the cudaIpc route (rb_shared_alloc / rb_ipc_open) for
    small buffers or hosts without symmetric memory.  The barrier is a 1-element NCCL all-reduce on the
    current stream (capturable in a CUDA graph)."""

    def __init__(self, group=None, device=None, legacy_ipc: Optional[bool] = None):
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._token = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.legacy_ipc = bool(int(os.environ.get("RB_P2P_LEGACY_IPC", "0"))) if legacy_ipc is None else legacy_ipc
        self._keep = []
        # Rendezvous of the ranks on the stream: a device-side barrier over symmetric-memory signal pads (a few
        # microseconds, one kernel, capturable) instead of a 1-element NCCL all-reduce (45-75 us at 8 GPUs).
        self._bar = None
        if self.world > 1 and not self.legacy_ipc and os.environ.get("RB_P2P_BARRIER", "symm") == "symm":
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty(64, dtype=torch.float32, device=self.device)
            self._bar = (t, symm_mem.rendezvous(t, self.group))

    def alloc(self, name, shape, dtype):
        if self.world == 1:
            buf = SharedBuffer(shape, dtype, self.device)
            self._keep.append(buf)
            return buf.tensor, [buf.ptr]
        if not self.legacy_ipc:
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty(*[int(x) for x in shape], dtype=dtype, device=self.device)
            hdl = symm_mem.rendezvous(t, self.group)
            self._keep.append((t, hdl))
            return t, [int(p) for p in hdl.buffer_ptrs]
        buf = SharedBuffer(shape, dtype, self.device)
        self._keep.append(buf)
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, buf.handle, group=self.group)
        ptrs = []
        for k, h in enumerate(handles):
            if k == self.rank:
                ptrs.append(buf.ptr)
                continue
            p = C.c_void_p()
            raw = (C.c_ubyte * _lib.RB_IPC_HANDLE_BYTES).from_buffer_copy(h)
            check(lib.rb_ipc_open(raw, C.byref(p)), f"rb_ipc_open({name}, rank {k})")
            ptrs.append(int(p.value))
        return buf.tensor, ptrs

    def barrier(self, channel: int = 0):
        if self._bar is not None:
            self._bar[1].barrier(channel=channel)
            return
        dist.all_reduce(self._token, group=self.group)
        self._token.zero_()


class LocalPeerLink(PeerLink):
    """G emulated ranks inside ONE process on ONE GPU (tests, single-GPU development): pointers are
    shared directly and the caller runs the ranks' phases in lock-step, so the barrier is a no-op."""

    def __init__(self, world: int, rank: int, registry: Dict[str, List[Optional[SharedBuffer]]], device=None):
        self.world, self.rank, self.registry = world, rank, registry
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())

    def alloc(self, name, shape, dtype):
        buf = SharedBuffer(shape, dtype, self.device)
        slots = self.registry.setdefault(name, [None] * self.world)
        slots[self.rank] = buf
        return buf.tensor, (lambda: [b.ptr for b in slots])   # resolved lazily, once every emulated rank has registered

    def barrier(self, channel: int = 0):
        pass


def _ptr_array(ptrs: Sequence[int]):
    return (C.c_void_p * len(ptrs))(*ptrs)


class _P2PInteractFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, emb, idx, dense_vec, flags, out_dtype, pad_to, ones_col=False):
        dense_vec = dense_vec.contiguous()
        out = emb._interaction_fwd(idx, dense_vec, flags, out_dtype, pad_to, ones_col)
        ctx.emb, ctx.idx, ctx.flags = emb, idx, flags
        ctx.save_for_backward(dense_vec)
        return out

    @staticmethod
    def backward(ctx, dOut):
        (dense_vec,) = ctx.saved_tensors
        if dOut.stride(-1) != 1:
            dOut = dOut.contiguous()
        d_dense = ctx.emb._interaction_bwd(ctx.idx, dense_vec, ctx.flags, dOut)
        return None, None, None, d_dense, None, None, None, None


class P2PShardedEmbedding(nn.Module):
    """`Embedding(input_dim, output_dim, num_tables=T)` row-wise sharded over `link.world` GPUs with the
    exchange done by the kernels over peer memory.  Call surface used by the models: `interact(...)`,
    `apply_pending(...)` (driven by optimizers.*), `load_full_table` / `full_row_ids` (tests)."""

    def __init__(self, input_dim: int, output_dim: int, *, num_tables: int = 1, link: PeerLink, device=None,
                 generator: Optional[torch.Generator] = None, capacity_factor: float = 1.25, bf16_shadow: bool = True,
                 table_rows: Optional[Sequence[int]] = None, replicate_rows_upto: int = 0):
        super().__init__()
        self.link = link
        self.world, self.rank = link.world, link.rank
        if self.world > _lib.RB_MAX_RANKS:
            raise ValueError(f"at most {_lib.RB_MAX_RANKS} ranks")
        if self.world > 1 and "RB_PDL" not in os.environ:
            # programmatic dependent launch pays on one GPU (step 1.328 -> 1.306 ms) and costs on the sharded step
            # (r2_69, N = 2: 1.380 -> 1.400 ms): off for every kernel this process launches from here on
            lib.rb_set_pdl(0)
        self.input_dim, self.output_dim, self.num_tables = int(input_dim), int(output_dim), int(num_tables)
        self.capacity_factor = float(capacity_factor)
        self.save_rows = True     # forward keeps the bf16 operand rows so the backward does not cross NVLink again
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        G = self.world
        # Table t starts at global row t * pitch.  pitch is coprime with G so that the SAME id of different tables
        # (the OOV / padding id 0 is the hottest row of every table, ctr/tfrecord_io.py:61-64) lands on different
        # ranks instead of piling up on rank 0; costs at most a few unused rows per table.
        self.table_rows = None if table_rows is None else [int(r) for r in table_rows]
        # Tables of at most `replicate_rows_upto` rows are NOT sharded: every rank keeps a full copy (`small_table`), reads it
        # locally in the forward, and all ranks apply the same update — each rank sums its own lookups' gradient rows per
        # row, the G short (row, sum) lists are all-gathered, one sparse update over their concatenation runs everywhere.
        # Row-wise sharding sends every lookup of a 3-row table to the same three owners: at the Criteo-Terabyte
        # cardinalities 11 of 26 tables (42 % of the lookups) have <= 2208 rows.  Real ranks only (the update is a collective).
        rows_list = self.table_rows if self.table_rows is not None else [self.input_dim] * self.num_tables
        self._small = [t for t, r in enumerate(rows_list) if r <= int(replicate_rows_upto)] if (
            int(replicate_rows_upto) > 0 and G > 1 and isinstance(link, DistPeerLink) and len(rows_list) > 1) else []
        if len(self._small) == len(rows_list):
            self._small = []                  # nothing left to shard: keep the plain layout
        self.small_base = None
        if self._small:
            small_set = set(self._small)
            big = [t for t in range(len(rows_list)) if t not in small_set]
            starts, end = [0] * len(rows_list), 0
            for j, t in enumerate(big):          # the sharded tables, staggered over the ranks like below
                s0 = end + ((j % G) - end % G) % G
                starts[t] = s0
                end = s0 + rows_list[t]
            shard_total = end
            self.small_base = (shard_total + G - 1) // G * G       # virtual rows [small_base, total_rows): the replicated tables
            off = 0
            for t in self._small:
                starts[t] = self.small_base + off
                off += rows_list[t]
            self.small_total = off
            self.num_tables = len(rows_list)
            self.table_rows = list(rows_list)
            self._starts = starts
            self.pitch = None
            self._big = big
            total = shard_total
        elif self.table_rows is not None:
            # per-table row counts (BASELINE config 3): table t starts at the first row >= the previous table's end whose
            # owner is rank t mod G — the same staggering of the tables' hot id 0
            self.num_tables = len(self.table_rows)
            starts, end = [], 0
            for t, r in enumerate(self.table_rows):
                s0 = end + ((t % G) - end % G) % G
                starts.append(s0)
                end = s0 + r
            self._starts = starts
            self.pitch = None
            total = end
        else:
            self.pitch = self.input_dim
            if self.num_tables > 1:
                while math.gcd(self.pitch, G) != 1:
                    self.pitch += 1
            self._starts = [t * self.pitch for t in range(self.num_tables)]
            self.table_rows = [self.input_dim] * self.num_tables
            total = self.pitch * self.num_tables
        if not self._small:
            self._big = list(range(self.num_tables))
        self.total_rows = total if not self._small else self.small_base + self.small_total      # what the kernels range-check ids against
        self.local_rows = max((total - self.rank + G - 1) // G, 1)       # rows r with r mod G == rank
        shard_rows = max((total + G - 1) // G, 1)                        # same shape on every rank (rank 0's count)
        self._shard_full, self._shard_ptrs = link.alloc("shard", (shard_rows, self.output_dim), torch.float32)
        self.embeddings = self._shard_full[: self.local_rows]
        self._shard_full.uniform_(-0.05, 0.05, generator=generator)     # Keras default initialiser
        # bf16 shadow of the shard: what the other ranks' forwards read over NVLink (half the bytes, and exactly the
        # MMA operands); kept in step by the optimizer row update
        self.use_shadow = bool(bf16_shadow) and self.world > 1
        self.small_table = self.small_shadow = None
        self._shadow_full = self._shadow_ptrs = self._shadow_ptr_dev = None
        if self.use_shadow:
            self._shadow_full, self._shadow_ptrs = link.alloc("shadow", (shard_rows, self.output_dim), torch.bfloat16)
            self.refresh_shadow()
        self._row_offset = (torch.tensor(self._starts, dtype=torch.int64, device=self.device) if self.num_tables > 1 else None)
        self._rows_dev = torch.tensor(self.table_rows, dtype=torch.int64, device=self.device)
        self._unsharded_starts = torch.tensor([0] + self.table_rows[:-1], dtype=torch.int64, device=self.device).cumsum(0)
        big_dev = torch.tensor(self._big, dtype=torch.int64, device=self.device)
        self._starts_dev = torch.tensor(self._starts, dtype=torch.int64, device=self.device)[big_dev]      # sharded tables, ascending
        self._big_rows_dev = self._rows_dev[big_dev]
        self._big_unsharded_starts = self._unsharded_starts[big_dev]
        self.small_state: Dict[str, torch.Tensor] = {}
        if self._small:
            D = self.output_dim
            self.small_table = torch.empty(self.small_total + 1, D, dtype=torch.float32, device=self.device)   # + one scratch row
            self.small_table.uniform_(-0.05, 0.05, generator=generator)
            self.small_table[-1].zero_()
            dist.broadcast(self.small_table, src=dist.get_global_rank(link.group, 0), group=link.group)         # one copy, everywhere
            self.small_shadow = self.small_table.to(torch.bfloat16)
            self._small_cols = torch.tensor(self._small, dtype=torch.int64, device=self.device)
            self._small_rel_off = torch.tensor([self._starts[t] - self.small_base for t in self._small], dtype=torch.int64, device=self.device)
        self._anchor = torch.zeros((), dtype=torch.float32, device=self.device, requires_grad=True)
        self.opt_state: Dict[str, torch.Tensor] = {}
        self._shape = None           # (B_local, F) the step buffers were built for
        self._pending = False
        self._side: Optional[torch.cuda.Stream] = None
        self._shard_ptr_dev: Optional[torch.Tensor] = None
        self._routed_by_caller = False
        self._begun = None
        self._grad_ready = None
        self._apply_done = None
        self._small_stream: Optional[torch.cuda.Stream] = None
        self._small_done = None

    # ---- lazily built per-shape step buffers -----------------------------------------------------------------
    def _resolve(self, v):
        return v() if callable(v) else v

    def _build(self, B: int, F: int) -> None:
        if self._shape == (B, F):
            return
        if self._shape is not None:
            raise ValueError(f"P2PShardedEmbedding was built for idx {self._shape}, got {(B, F)} (static shapes)")
        dev, G, D = self.device, self.world, self.output_dim
        n = B * F
        self.n_local = n
        self.capacity = int(n * self.capacity_factor) + 1024
        self._dE, self._dE_ptrs = self.link.alloc("dE", (n, D), torch.float32)
        self._b_rows, self._rows_ptrs = self.link.alloc("b_rows", (n,), torch.int64)
        self._b_perm, self._perm_ptrs = self.link.alloc("b_perm", (n,), torch.int32)
        self._b_counts, self._counts_ptrs = self.link.alloc("b_counts", (max(G, 8),), torch.int64)
        self._inv_perm = torch.empty(n, dtype=torch.int32, device=dev)
        self._bucket_ws = torch.empty(max(lib.rb_bucket_by_owner_workspace_bytes(n, G), 256), dtype=torch.uint8, device=dev)
        self._sort_ws = ops.sparse_workspace(self.capacity, D, self.local_rows + 1, dev)
        self._n_valid = torch.zeros(1, dtype=torch.int32, device=dev)
        self._x_saved = torch.empty(B, F + 1, D, dtype=torch.bfloat16, device=dev)    # forward -> backward operand rows
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        # the flag is mirrored into pinned host memory behind every collect (an async copy on the side stream, capturable),
        # so that the NEXT step's host code can notice an overflow without synchronising (poll_overflow)
        self._overflow_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        if self._small:
            Fs = len(self._small)
            ns = B * Fs
            cap = min(ns, self.small_total)
            self._small_cap = cap
            self._idx_small = None                                      # [B, Fs] ids of the replicated tables' columns (built per step)
            slot = torch.full((F,), -1, dtype=torch.int32)
            slot[torch.tensor(self._small)] = torch.arange(Fs, dtype=torch.int32)
            self._de_slot = slot.to(dev)
            self._dE_small = torch.empty(B, Fs, D, dtype=torch.float32, device=dev)
            self._su_rows = torch.empty(cap + 8, dtype=torch.int64, device=dev)
            self._su_grad = torch.empty(cap + 8, D, dtype=torch.float32, device=dev)
            self._su_num = torch.zeros(1, dtype=torch.int64, device=dev)
            self._cap_arange = torch.arange(cap, dtype=torch.int64, device=dev)
            self._scratch_row = torch.full((1,), self.small_total, dtype=torch.int64, device=dev)
            self._zero = torch.zeros((), dtype=torch.float32, device=dev)
            # one record per (row, summed gradient): D gradient values + the row id as int32 bits (rb_replicated_rows_update)
            self._pack_pad = torch.empty(cap, D + 1, dtype=torch.float32, device=dev)
            self._pack_all = torch.empty(G, cap, D + 1, dtype=torch.float32, device=dev)
            self._small_ws = ops.sparse_workspace(max(ns, 1), D, self.small_total, dev)
        self._shape = (B, F)
        self._side = torch.cuda.Stream(device=dev)

    def refresh_shadow(self) -> None:
        """Re-derive the bf16 shadow from the fp32 shard (after loading weights; the optimizer keeps it in step afterwards)."""
        if self._shadow_full is not None:
            self._shadow_full.copy_(self._shard_full)
        if self.small_shadow is not None:
            self.small_shadow.copy_(self.small_table)

    def _shadow_ptrs_dev(self):
        if not self.use_shadow:
            return None
        if self._shadow_ptr_dev is None:
            self._shadow_ptr_dev = torch.tensor(self._resolve(self._shadow_ptrs), dtype=torch.int64, device=self.device)
        return self._shadow_ptr_dev.data_ptr()

    def _shard_ptrs_dev(self) -> torch.Tensor:
        if self._shard_ptr_dev is None:
            self._shard_ptr_dev = torch.tensor(self._resolve(self._shard_ptrs), dtype=torch.int64, device=self.device)
        return self._shard_ptr_dev

    # ---- shard <-> full table (tests, checkpoints) ---------------------------------------------------------------
    def full_row_ids(self) -> torch.Tensor:
        """For every local shard row: its row in the UNSHARDED [input_dim * num_tables, D] table, or -1 for the
        pad rows the table pitch introduces."""
        g = torch.arange(self.local_rows, device=self.device, dtype=torch.int64) * self.world + self.rank   # global (staggered) row
        t = torch.searchsorted(self._starts_dev, g, right=True) - 1
        i = g - self._starts_dev[t]
        ok = i < self._big_rows_dev[t]
        return torch.where(ok, self._big_unsharded_starts[t] + i, torch.full_like(g, -1))

    def cap_ids(self, ids: torch.Tensor) -> torch.Tensor:
        """row-in-table = id mod rows(table) for non-negative raw ids [B, num_tables]."""
        return ids % self._rows_dev[None].to(ids.dtype)

    def load_full_table(self, full: torch.Tensor) -> None:
        """Adopt this rank's rows of an unsharded [sum(table_rows), D] table."""
        full = torch.as_tensor(full, dtype=torch.float32).to(self.device)
        ids = self.full_row_ids()
        ok = ids >= 0
        self.embeddings[ok] = full[ids[ok]]
        for t in self._small:
            r, o = self.table_rows[t], self._starts[t] - self.small_base
            u = int(self._unsharded_starts[t])
            self.small_table[o:o + r] = full[u:u + r]
        self.refresh_shadow()

    def small_into_full(self, full: torch.Tensor, local: Optional[torch.Tensor] = None) -> None:
        """full[unsharded rows of the replicated tables] = this rank's copy (or `local`, a tensor shaped like it)."""
        src = self.small_table if local is None else local
        for t in self._small:
            r, o = self.table_rows[t], self._starts[t] - self.small_base
            u = int(self._unsharded_starts[t])
            full[u:u + r] = src[o:o + r].to(full.device)

    def scatter_into_full(self, full: torch.Tensor, local: Optional[torch.Tensor] = None) -> None:
        """full[unsharded row] = this shard's rows (or `local`, e.g. an optimizer state of the same shape)."""
        ids = self.full_row_ids()
        ok = ids >= 0
        src = self.embeddings if local is None else local
        full[ids[ok]] = src[ok].to(full.device)

    # ---- the step, phase by phase (DistPeerLink: called in this order by interact/apply_pending) ---------------------
    def route(self, idx: torch.Tensor) -> None:
        """Phase 1: publish this rank's lookups bucketed by owner (peer-readable)."""
        B, F = idx.shape
        self._build(B, F)
        off = self._row_offset if self.num_tables > 1 else None
        skip = self.small_base if self._small else (1 << 63) - 1      # lookups of the replicated tables belong to no owner
        check(lib.rb_bucket_by_owner_skip(idx.data_ptr(), ops._idx(idx), B * F, F, ops._ptr(off), 0, self.world, skip,
                                          self._b_rows.data_ptr(), self._b_perm.data_ptr(), self._inv_perm.data_ptr(),
                                          self._b_counts.data_ptr(),
                                          self._bucket_ws.data_ptr(), self._bucket_ws.numel(), ops._stream()), "rb_bucket_by_owner_skip")
        if self._small:
            if self._idx_small is None or self._idx_small.dtype != idx.dtype:
                self._idx_small = torch.empty(B, len(self._small), dtype=idx.dtype, device=idx.device)
            torch.index_select(idx, 1, self._small_cols, out=self._idx_small)

    def collect_and_sort(self) -> None:
        """Phase 2 (after every rank's route()): gather the pairs addressed to this owner and radix-sort them."""
        check(lib.rb_p2p_collect_keys(self.world, self.rank, self.n_local, _ptr_array(self._resolve(self._rows_ptrs)),
                                      _ptr_array(self._resolve(self._perm_ptrs)), _ptr_array(self._resolve(self._counts_ptrs)),
                                      self.local_rows, self.capacity, self._sort_ws.data_ptr(), self._sort_ws.numel(),
                                      self.output_dim, self._n_valid.data_ptr(), self.overflow.data_ptr(), ops._stream()),
              "rb_p2p_collect_keys")
        sel = C.c_int32(0)
        check(lib.rb_sparse_bwd_prepare_collected(self.local_rows, self.output_dim, self.capacity, self._sort_ws.data_ptr(),
                                                  self._sort_ws.numel(), C.byref(sel), ops._stream()), "rb_sparse_bwd_prepare_collected")
        self._sel = int(sel.value)
        self._overflow_host.copy_(self.overflow, non_blocking=True)

    def _interaction_fwd(self, idx, dense_vec, flags, out_dtype, pad_to, ones_col=False):
        si, sg, tail = flags
        B, F = idx.shape
        D = self.output_dim
        Fp = F + 1
        width = ops.interaction_ncols(Fp, si, sg) + (D if tail else 0)
        stride = (width + pad_to - 1) // pad_to * pad_to if out_dtype == torch.bfloat16 else width
        out = torch.empty(B, stride, dtype=out_dtype, device=idx.device)
        off = self._row_offset if self.num_tables > 1 else None
        check(lib.rb_dot_interaction_fwd_sharded_rep(self._shard_ptrs_dev().data_ptr(), self.world, self.total_rows, idx.data_ptr(),
                                                     ops._idx(idx), ops._ptr(off), dense_vec.data_ptr(), B, F, D, int(si), int(sg),
                                                     int(tail), out.data_ptr(),
                                                     _lib.RB_BF16_ONES if (ones_col and out_dtype == torch.bfloat16 and stride > width)
                                                     else ops._float_type(out_dtype), stride,
                                                     self._x_saved.data_ptr() if self.save_rows else None, self._shadow_ptrs_dev(),
                                                     self.small_base if self._small else self.total_rows, ops._ptr(self.small_table),
                                                     ops._ptr(self.small_shadow), ops._stream()), "rb_dot_interaction_fwd_sharded_rep")
        return out

    def _interaction_bwd(self, idx, dense_vec, flags, dOut):
        si, sg, tail = flags
        B, F = idx.shape
        D = self.output_dim
        d_dense = torch.empty(B, D, dtype=torch.float32, device=idx.device)
        off = self._row_offset if self.num_tables > 1 else None
        if self._small and not self.save_rows:
            raise RuntimeError("replicated tables need save_rows=True (the backward reads the operand rows the forward saved)")
        check(lib.rb_dot_interaction_bwd_sharded_split(self._shard_ptrs_dev().data_ptr(), self.world, self.total_rows, idx.data_ptr(),
                                                       ops._idx(idx), ops._ptr(off), dense_vec.data_ptr(), B, F, D, int(si), int(sg),
                                                       int(tail), dOut.data_ptr(), ops._float_type(dOut.dtype), int(dOut.stride(0)),
                                                       self._dE.data_ptr(), d_dense.data_ptr(),
                                                       self._x_saved.data_ptr() if self.save_rows else None,
                                                       ops._ptr(self._de_slot) if self._small else None,      # replicated tables' rows
                                                       ops._ptr(self._dE_small) if self._small else None,      # go to their own tensor
                                                       len(self._small), ops._stream()),
              "rb_dot_interaction_bwd_sharded_split")
        self._pending = True
        self._grad_ready = torch.cuda.current_stream().record_event()
        return d_dense

    def begin_step(self, idx: torch.Tensor) -> None:
        """Start this step's routing on the side stream as soon as the ids exist (the model calls it before the
        bottom MLP, which needs no embeddings): publish the bucket arrays, meet the other ranks, then collect and
        sort the pairs this rank owns.  interact() waits for the rendezvous, apply_pending() for the sort."""
        idx = idx.contiguous()
        B, F = idx.shape
        if not torch.cuda.is_current_stream_capturing():
            self.poll_overflow()
        self._build(B, F)
        main, side = torch.cuda.current_stream(), self._side
        side.wait_stream(main)                   # ids are ready; last step's apply (on main) precedes this rank's arrival
        with torch.cuda.stream(side):
            self.route(idx)
            self.link.barrier(0)                 # every rank's bucket arrays are published and last step's updates are done
            self._routed_ev = side.record_event()
            if torch.is_grad_enabled():
                self.collect_and_sort()
            self._sorted_ev = side.record_event()
        if not torch.cuda.is_current_stream_capturing():
            idx.record_stream(side)
        self._begun = idx

    def interact(self, idx: torch.Tensor, dense_vec: torch.Tensor, self_interaction=False, skip_gather=True, tail=True,
                 out_dtype=torch.float32, pad_to=1, routed=False, ones_col=False) -> torch.Tensor:
        """ctr/model.py:49-55 on the sharded table.  Unless the caller already ran route()/collect_and_sort()
        (`routed=True`, lock-step emulation) the step's routing runs on the side stream (begin_step)."""
        idx = idx.contiguous()
        self._routed_by_caller = bool(routed)
        if not routed:
            if self._begun is None or self._begun.data_ptr() != idx.data_ptr():
                self.begin_step(idx)
            self._begun = None
            torch.cuda.current_stream().wait_event(self._routed_ev)
        return _P2PInteractFn.apply(self._anchor, self, idx, dense_vec.float(), (self_interaction, skip_gather, tail), out_dtype, pad_to,
                                    ones_col)

    def apply_pending(self, kind: str, step: int, lr: float, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
                      initial_accumulator_value=0.1, alpha_dev=None) -> int:
        """Owner-side segmented reduction + optimizer row update, gradient rows pulled from all ranks' dE."""
        if not self._pending:
            return 0
        self._pending = False
        st = self.opt_state
        fresh_state = False
        if kind == "adam_lazy":
            if "m" not in st:
                st["m"], st["v"] = torch.zeros_like(self.embeddings), torch.zeros_like(self.embeddings)
                fresh_state = True
            s0, s1 = st["m"], st["v"]
        elif kind == "adagrad":
            if "acc" not in st:
                st["acc"] = torch.full_like(self.embeddings, initial_accumulator_value)
                fresh_state = True
            s0, s1 = st["acc"], None
        elif kind == "sgd":
            s0 = s1 = None
        else:
            raise ValueError(f"the peer-memory path supports adam_lazy / adagrad / sgd, not {kind}")
        opt = ops._opt_params(kind, step, lr, beta_1, beta_2, epsilon, alpha_dev)

        def launch():
            self._apply_rows(s0, s1, opt)

        if self._routed_by_caller:
            launch()
            return self.n_local
        # on the side stream (which already holds this step's sort): start as soon as this rank's dE is written, meet the
        # other ranks, update — overlapping the rest of the backward, the dense all-reduce and the dense step
        side = self._side
        if self._grad_ready is not None and not fresh_state:
            side.wait_event(self._grad_ready)
        else:
            side.wait_stream(torch.cuda.current_stream())
        self._grad_ready = None
        if self._small:
            # the replicated tables' update shares nothing with the sharded one but dE: its own stream, beside it
            if self._small_stream is None:
                self._small_stream = torch.cuda.Stream(device=self.device)
            self._small_stream.wait_stream(side)
            with torch.cuda.stream(self._small_stream):
                self._apply_small(kind, opt, initial_accumulator_value)
                self._small_done = self._small_stream.record_event()
        with torch.cuda.stream(side):
            self.link.barrier(1)                 # every rank's backward has written its dE and stopped reading the shards
            launch()
            if self._small:
                side.wait_event(self._small_done)
            self._apply_done = side.record_event()
        return self.n_local

    def _apply_small(self, kind: str, opt, initial_accumulator_value: float) -> None:
        """The replicated tables' step (current stream = side stream, after this rank's dE is written): local duplicate-row sum
        (rb_sparse_bwd_dedup) -> all-gather of the G (row, sum) lists, padded to a static length with a scratch row ->
        ONE sparse update over their concatenation (rb_sparse_bwd_update_groups), identical on every rank."""
        B, F = self._shape
        D, G, Fs, cap = self.output_dim, self.world, len(self._small), self._small_cap
        st = self.small_state
        if kind == "adam_lazy":
            if "m" not in st:
                st["m"], st["v"] = torch.zeros_like(self.small_table), torch.zeros_like(self.small_table)
            s0, s1 = st["m"], st["v"]
        elif kind == "adagrad":
            if "acc" not in st:
                st["acc"] = torch.full_like(self.small_table, initial_accumulator_value)
            s0, s1 = st["acc"], None
        else:
            s0 = s1 = None
        self._small_reduce()
        self._small_exchange()
        self._small_update(s0, s1, opt)

    def _small_reduce(self) -> None:
        B, F = self._shape
        D, G, Fs, cap = self.output_dim, self.world, len(self._small), self._small_cap
        grad = ops.GradSource.per_position(self._dE_small, Fs).to_c()        # written by the backward itself (de_slot)
        check(lib.rb_sparse_bwd_dedup(self.small_total, D, self._idx_small.data_ptr(), ops._idx(self._idx_small), B * Fs, Fs,
                                      self._small_rel_off.data_ptr(), 0, C.byref(grad), self._su_rows.data_ptr(), self._su_grad.data_ptr(),
                                      self._su_num.data_ptr(), self._small_ws.data_ptr(), self._small_ws.numel(),
                                      ops._ptr(ops.oob_flag(self.device)), ops._stream()), "rb_sparse_bwd_dedup (replicated tables)")
        valid = self._cap_arange < self._su_num                    # records behind the last unique row: scratch id, zero gradient
        self._pack_pad[:, :D].copy_(torch.where(valid[:, None], self._su_grad[:cap], self._zero))
        self._pack_pad.view(torch.int32)[:, D].copy_(torch.where(valid, self._su_rows[:cap], self._scratch_row))

    def _small_exchange(self) -> None:
        dist.all_gather_into_tensor(self._pack_all.view(-1, self.output_dim + 1), self._pack_pad, group=self.link.group)

    def _small_update(self, s0, s1, opt) -> None:
        check(lib.rb_replicated_rows_update(self.small_table.data_ptr(), ops._ptr(s0), ops._ptr(s1), self.small_total, self.output_dim,
                                            self._pack_all.data_ptr(), self.world, self._small_cap, C.byref(opt),
                                            ops._ptr(self.small_shadow), ops._stream()), "rb_replicated_rows_update")

    def _apply_rows(self, s0, s1, opt) -> None:
        """The C call of the apply phase, on the current (side) stream: what bench.py brackets with CUDA events."""
        _, F = self._shape
        check(lib.rb_sparse_bwd_apply_p2p(self.embeddings.data_ptr(), ops._ptr(s0), ops._ptr(s1), self.local_rows, self.output_dim,
                                          self.world, self.n_local, F, _ptr_array(self._resolve(self._dE_ptrs)), self.capacity,
                                          self._n_valid.data_ptr(), C.byref(opt), self._sort_ws.data_ptr(), self._sort_ws.numel(),
                                          self._sel, ops._ptr(self._shadow_full), ops._stream()), "rb_sparse_bwd_apply_p2p")

    def join(self) -> None:
        if self._apply_done is not None:
            torch.cuda.current_stream().wait_event(self._apply_done)
            self._apply_done = None

    def poll_overflow(self) -> None:
        """Non-blocking: raises if an EARLIER step's collect saw more lookups addressed to this owner than the static
        capacity (`capacity_factor` x the per-rank lookups + 1024; the mean load is 1.0 x).  The pairs beyond the capacity
        were cut off, i.e. their gradients were NOT applied — training must not continue silently.  Called at the start
        of every step (begin_step) and by graph.GraphedTrainStep before each replay; check_overflow() is the synchronising form.
        Skewed ids concentrate on few owners (a 3-row table sends all its lookups to 3 ranks): BASELINE config 3 runs with
        capacity_factor = 2.0."""
        if self._shape is not None and int(self._overflow_host[0]) != 0:
            self._overflow_host.zero_()
            raise RuntimeError("an owner received more lookups than its static capacity in an earlier step and dropped the excess "
                               "gradient rows; raise capacity_factor (p2p.P2PShardedEmbedding)")

    def check_overflow(self) -> None:
        if int(self.overflow.item()) != 0:
            self.overflow.zero_()
            self._overflow_host.zero_()
            raise RuntimeError("an owner received more lookups than its static capacity; raise capacity_factor")


class P2PShardedDLRM(nn.Module):
    """ctr/model.py:34-58 with the table sharded row-wise over peer memory and the MLPs replicated."""

    def __init__(self, bottom_mlp_units: Sequence[int], top_mlp_units: Sequence[int], embedding_size: int, vocab_size: int,
                 num_cat_fea: int, num_int_fea: int, *, num_tables: int = 1, group=None, link: Optional[PeerLink] = None, device=None,
                 compute_dtype: Optional[torch.dtype] = None, generator: Optional[torch.Generator] = None,
                 capacity_factor: float = 1.25, bf16_shadow: bool = True, table_rows: Optional[Sequence[int]] = None,
                 replicate_rows_upto: int = 0):
        super().__init__()
        if bottom_mlp_units[-1] != embedding_size:
            raise ValueError("bottom_mlp_units[-1] must equal embedding_size")       # ctr/model.py:52,55
        self.group = group
        self.link = link if link is not None else DistPeerLink(group, device)
        self.bottom_mlp = MLP(bottom_mlp_units, "relu", compute_dtype=compute_dtype, generator=generator)
        self.top_mlp = MLP(top_mlp_units, "sigmoid", compute_dtype=compute_dtype, generator=generator)
        self.embedding_layer = P2PShardedEmbedding(vocab_size, embedding_size, num_tables=num_tables, link=self.link, device=device,
                                                   generator=generator, capacity_factor=capacity_factor, bf16_shadow=bf16_shadow,
                                                   table_rows=table_rows, replicate_rows_upto=replicate_rows_upto)
        self.num_cat_fea, self.num_int_fea, self.embedding_size = num_cat_fea, num_int_fea, embedding_size
        self._synced = False
        self._flat = None
        self._reduce_stream = None
        self._reduce_ev = None

    def sync_dense_parameters(self) -> None:
        if isinstance(self.link, DistPeerLink) and self.link.world > 1:
            for p in self.parameters():
                dist.broadcast(p.data, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
                if hasattr(p, "_rb_shadow"):
                    del p._rb_shadow          # .data writes do not bump the version the bf16 shadow is keyed on
        self._synced = True

    def forward(self, inputs, training=None, mask=None, routed=False):
        int_features = inputs["int_features"].reshape(-1, self.num_int_fea)
        cat_features = inputs["cat_features"].reshape(-1, self.num_cat_fea)
        width = (self.num_cat_fea + 1) ** 2 + self.embedding_size
        if not self._synced:         # build both towers, then adopt rank 0's weights BEFORE anything is computed from them
            if len(self.bottom_mlp.kernels) == 0:
                self.bottom_mlp.build(self.num_int_fea, int_features.device)
            if len(self.top_mlp.kernels) == 0:
                self.top_mlp.build(width, int_features.device)
            self.sync_dense_parameters()
            self._attach_flat_grads()
        if not routed:
            self.embedding_layer.begin_step(cat_features)      # routing + rendezvous overlap the bottom MLP
        bmlp_output = self.bottom_mlp(int_features)
        bf16 = self.top_mlp.compute_dtype == torch.bfloat16
        ones = bf16 and width % 8 != 0     # a spare pad column carries 1.0: the first top layer's bias gradient comes with its dW GEMM
        tmlp_input = self.embedding_layer.interact(cat_features, bmlp_output, False, True, True,
                                                   out_dtype=torch.bfloat16 if bf16 else torch.float32, pad_to=8 if bf16 else 1,
                                                   routed=routed, ones_col=ones)
        return (self.top_mlp(tmlp_input, ones_col=True) if ones else self.top_mlp(tmlp_input)).squeeze(1)

    def _attach_flat_grads(self) -> None:
        """All MLP gradients live in ONE flat buffer (the parameters' .grad are views of it): the replicas' sum is
        a single all-reduce with no gather / scatter copies around it."""
        params = [p for p in self.parameters() if p.requires_grad]
        if self._flat is None or self._flat.numel() != sum(p.numel() for p in params):
            self._flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
        o = 0
        for p in params:
            n = p.numel()
            p.grad = self._flat[o:o + n].view_as(p)
            o += n

    def reset_dense_grads(self) -> None:
        """Called by the optimizers instead of dropping the .grad tensors."""
        if self._flat is not None:
            self._flat.zero_()

    def reduce_dense_grads(self) -> None:
        """SUM over replicas of the MLP gradients (MirroredStrategy with Reduction.NONE losses).
        Started on a side stream so that it overlaps the sparse apply; wait_dense_grads() joins it before the
        dense optimizer step (the optimizers call both)."""
        self._reduce_ev = None
        if self._flat is not None and isinstance(self.link, DistPeerLink) and self.link.world > 1:
            if self._reduce_stream is None:
                self._reduce_stream = torch.cuda.Stream(device=self._flat.device)
            main = torch.cuda.current_stream()
            self._reduce_stream.wait_stream(main)
            with torch.cuda.stream(self._reduce_stream):
                dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
                self._reduce_ev = self._reduce_stream.record_event()

    def wait_dense_grads(self) -> None:
        if self._reduce_ev is not None:
            torch.cuda.current_stream().wait_event(self._reduce_ev)
            self._reduce_ev = None
