"""The whole training step as ONE CUDA graph.

At config-2 sizes a DLRM step is ~140 kernel launches in ~2 ms: launched from Python the GPU waits
on the host.  `GraphedTrainStep` captures forward + loss + backward + optimizer (dense and sparse,
including the side-stream radix sort that overlaps the forward) once and replays it; the only
per-step host work is copying the batch into the graph's static input buffers and refreshing the
step-dependent Adam scalar (`optimizers.Adam.prepare_step`, read by the kernels from device memory
through `rb_opt_params.alpha_t_dev`).  The kernels and their order are exactly those of the eager
path — the graph changes who launches them, not what runs.
"""
from __future__ import annotations

import gc
import os
from typing import Callable, Optional, Sequence

import torch


class GraphedTrainStep:
    """step(batch) == { prob = model(inputs); loss = loss_fn(prob, label); loss.backward();
    optimizer.apply_gradients(model) } with the same results as the eager sequence.

    batch = (cat_features int[B,F], int_features f32[B,13], label) CUDA tensors of the captured shapes.
    The returned loss is a static device tensor overwritten by every replay."""

    def __init__(self, model: torch.nn.Module, optimizer, loss_fn: Callable, sample_batch: Sequence[torch.Tensor], warmup: int = 3,
                 fuse_sparse_updates: bool = False):
        if not all(t.is_cuda for t in sample_batch):
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (there is no CPU path)")
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        from .model import bce_clipped as _bce_clipped
        # Keras' clipped binary cross-entropy on a model that offers the fused form (DLRM.forward_bce: Dense(1, sigmoid) + loss + the
        # head's backward in one kernel).  Opt-in (RB_FUSED_HEAD=1): bit-identical gradients, but measured neutral on the step
        # (r2_45 / r2_47: 1.340 / 1.345 ms against 1.345 / 1.344 ms) — the three small kernels it replaces were not what the chain waits on
        self._fused_loss = loss_fn is _bce_clipped and hasattr(model, "forward_bce") and os.environ.get("RB_FUSED_HEAD", "0") == "1"
        self._pollers = [m for m in model.modules() if hasattr(m, "poll_overflow")]
        self.static = tuple(torch.empty_like(t) for t in sample_batch)
        if hasattr(optimizer, "enable_device_scalars"):
            optimizer.enable_device_scalars(sample_batch[0].device)
        # a captured step always runs backward + apply_gradients together, so the optimizer's row update of the rows a step
        # touches once MAY run inside the backward (optimizers.fuse_sparse_updates).  Opt-in (argument or RB_FUSED_UPDATE=1):
        # measured slower than the two-kernel chain on a B200 (r2_22 / r2_23) — the exactly-rounded Adam arithmetic
        # needs the occupancy of the reduction kernel to hide, which the interaction kernel's shared-memory ring does not leave
        if (fuse_sparse_updates or os.environ.get("RB_FUSED_UPDATE", "0") == "1") and hasattr(optimizer, "fuse_sparse_updates"):
            optimizer.fuse_sparse_updates(model)
        for d, s in zip(self.static, sample_batch):
            d.copy_(s)
        # Autograd graphs of earlier eager steps pin their AccumulateGrad nodes to the stream they ran on (the
        # legacy default stream breaks capture): the caller must not hold such a graph (e.g. a non-detached loss).
        gc.collect()
        # warm-up off the default stream (lazy builds, workspaces, cuBLAS handles, autograd threads)
        main = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=sample_batch[0].device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.optimizer.prepare_step()
                self._body()
        main.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.optimizer.prepare_step()                       # host half of the captured step
        # Stream priorities are recorded into the captured kernel nodes.  The step's critical chain runs at the highest one,
        # the weight-gradient stream below it (layers.MLP), the sort + row-update stream at the default: when the HBM-bound
        # row update and the towers' backward are runnable together, the Dense CTAs are placed first and the update fills the
        # SMs around them (timeline r2_19: the bottom tower's backward moved from behind the update to beside it)
        capture_stream = torch.cuda.Stream(device=sample_batch[0].device, priority=int(os.environ.get("RB_PRIO_MAIN", "-2")))
        with torch.cuda.graph(self.graph, stream=capture_stream):
            self.loss = self._body()
        self.steps_run = max(1, warmup) + 1                 # capture itself does not execute: counted when replayed below
        self.graph.replay()
        self._capture_stream = capture_stream
        self._bound = {}                                    # data_ptr of a registered batch's first tensor -> (graph, loss)
        self._read_stream: Optional[torch.cuda.Stream] = None   # step(loss_out=...): the loss goes to the host beside the next replay
        self._reads = {}                                    # id(graph) -> [step-done event, read-done event, read pending]

    def bind_inputs(self, batches: Sequence[Sequence[torch.Tensor]]) -> None:
        """Capture one more graph per batch in `batches` that reads THOSE tensors in place: step(batch) with a registered batch
        replays its own graph and skips the copy into the static input buffers (an input pipeline that fills K device staging
        buffers in turn registers the K buffers once).  Same kernels, same order, same results; capture executes nothing and
        leaves the optimizer's step count alone."""
        opt = self.optimizer
        for batch in batches:
            batch = tuple(batch)
            if len(batch) != len(self.static) or any(b.shape != s.shape or b.dtype != s.dtype or not b.is_cuda or not b.is_contiguous()
                                                     for b, s in zip(batch, self.static)):
                raise ValueError("bind_inputs takes contiguous CUDA batches of the captured shapes and dtypes")
            key = batch[0].data_ptr()
            if key in self._bound:
                continue
            saved = (opt.iterations, opt._prepared)
            opt._prepared = True                            # the body must not advance the host-side step count
            static, self.static = self.static, batch
            g = torch.cuda.CUDAGraph()
            try:
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=self._capture_stream):
                    loss = self._body()
            finally:
                self.static = static
                opt.iterations, opt._prepared = saved
            self._bound[key] = (g, loss, batch)

    def _body(self) -> torch.Tensor:
        cat, dense, label = self.static
        inputs = {"cat_features": cat, "int_features": dense}
        if self._fused_loss:
            loss = self.model.forward_bce(inputs, label)       # the same loss, head + loss + head backward in one kernel
        else:
            loss = self.loss_fn(self.model(inputs), label)
        loss.backward()
        self.optimizer.apply_gradients(self.model)
        return loss.detach()

    def step(self, batch: Sequence[torch.Tensor], loss_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Replay the step on `batch`.  `loss_out` (a pinned host float32 scalar / 1-element tensor) receives the step's loss through
        an asynchronous copy on a private stream: queued on the launching stream, the 4-byte copy would sit between two graph
        launches and cost ~40 us a step (profiles r2_60 -> r2_62); the replay that next rewrites the same loss tensor waits for it."""
        bound = self._bound.get(batch[0].data_ptr()) if self._bound else None
        if bound is not None and all(b.data_ptr() == r.data_ptr() for b, r in zip(batch, bound[2])):
            graph, loss = bound[0], bound[1]                # this batch's own graph reads it in place
        else:
            graph, loss = self.graph, self.loss
            for d, s in zip(self.static, batch):
                d.copy_(s, non_blocking=True)
        self.optimizer.prepare_step()
        for m in self._pollers:
            m.poll_overflow()            # a sharded table that dropped gradient rows in an earlier replay (p2p.poll_overflow)
        main = torch.cuda.current_stream()
        events = self._reads.get(id(graph))
        if events is not None and events[2]:
            main.wait_event(events[1])                      # the last read of the loss tensor this replay rewrites
            events[2] = False
        graph.replay()
        self.steps_run += 1
        if loss_out is not None:
            if self._read_stream is None:
                self._read_stream = torch.cuda.Stream(device=loss.device)
            if events is None:
                events = self._reads[id(graph)] = [torch.cuda.Event(), torch.cuda.Event(), False]     # step done, read done, read pending
            events[0].record(main)
            self._read_stream.wait_event(events[0])
            with torch.cuda.stream(self._read_stream):
                loss_out.copy_(loss.reshape(loss_out.shape), non_blocking=True)
                events[1].record(self._read_stream)
            events[2] = True
        return loss
