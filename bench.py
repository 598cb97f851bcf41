#!/usr/bin/env python
"""Headline benchmark: DLRM Criteo-shape training throughput (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one full training step of DLRM (ctr/model.py:34-58 + Adam, ctr/train.py:80,97) on one
synthetic Criteo-shaped batch: fused lookup + dot interaction forward, MLPs forward/backward,
interaction backward, and the sorted backward scatter with the fused sparse Adam row update.
N = 1 workload = BASELINE config 2: emb dim 64, batch 65536, 26 tables of 1M rows, dot interaction.
N > 1 (launched by torchrun, one rank per GPU): the same model with the tables row-wise sharded
over the ranks and B_local = 65536 per GPU (weak scaling), NCCL all-to-all each way.

One JSON line is printed by rank 0 (contract in the task statement): `value` = samples/s with the
inputs resident in HBM; `e2e` = samples/s through the public API from pinned HOST buffers with the
host->device copies and the device->host loss read inside the timed region; `roofline` for the
dominant C-ABI call; `cpu_baseline` = the torch-CPU restatement of the reference timed on this
box's host cores on a bounded sample.  `--impl reference` times only that restatement.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep stdout to the one JSON line

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "train samples/s, DLRM Criteo-shape"
UNIT = "samples/s"
BOTTOM, TOP = [512, 256, 64], [512, 256, 1]
F_CAT, F_INT = 26, 13
# MLPerf-DLRM capping (row = id mod 40M) of the Criteo-Terabyte cardinalities: 187.8M rows in 26 tables
CRITEO_TB_ROWS = [39884406, 39043, 17289, 7420, 20263, 3, 7120, 1543, 63, 38532951, 2953546, 403346, 10, 2208, 11938, 155, 4, 976, 14,
                  39979771, 25641295, 39664984, 585935, 12972, 108, 36]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="per-GPU batch")
    ap.add_argument("--emb-dim", type=int, default=64)
    ap.add_argument("--rows-per-table", type=int, default=1_000_000)
    ap.add_argument("--tables", type=int, default=26, choices=[1, 26])
    ap.add_argument("--dist", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--mlp-dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--mlp-backend", default="tcgen05", choices=["tcgen05", "cublas"],
                    help="bf16 towers on the hand-written tcgen05 Dense kernels (csrc/mlp.cu) or on torch / cuBLASLt (the round-1 path, for A/B)")
    ap.add_argument("--sharding", default="row", choices=["row", "table"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: rows/gradients read by the kernels over NVLink peer memory (p2p) or exchanged with NCCL all-to-alls")
    ap.add_argument("--ring", type=int, default=8, help="distinct synthetic batches cycled through")
    ap.add_argument("--collapse-mlp", action="store_true",
                    help="single GPU, opt-in: evaluate each tower (linear hidden layers, ctr/layers.py:8) as one affine map "
                         "(layers._CollapsedAffineFn); the default runs the layers one GEMM at a time like the reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--criteo-tb", action="store_true",
                    help="BASELINE config 3 (N > 1 only): 26 tables with the capped Criteo-Terabyte cardinalities (187.8M rows), row-wise "
                         "sharded.  This is the default workload of every N > 1 run on the peer-memory path")
    ap.add_argument("--config2-sharded", action="store_true",
                    help="N > 1: keep BASELINE config 2 (26 x 1M-row tables) row-wise sharded instead of config 3")
    ap.add_argument("--capacity-factor", type=float, default=0.0,
                    help="static per-owner pair capacity of the sharded table in units of the per-rank lookups (0 = 2.0 for config 3, 1.25 otherwise)")
    ap.add_argument("--replicate-small", type=int, default=4096,
                    help="config 3: tables of at most this many rows are replicated on every rank instead of row-sharded (0 = shard all; "
                         "r2_39 at N = 8: 1.911 -> 1.624 ms per step)")
    ap.add_argument("--no-bind-inputs", action="store_true",
                    help="replay ONE graph and copy every batch into its static input buffers first (default: one graph per resident / staging batch)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying one CUDA graph per step")
    ap.add_argument("--cpu-sample-batch", type=int, default=0,
                    help="batch of the CPU reference arm; 0 = the GPU arm's batch (its Keras dense-Adam passes cost the same whatever the batch)")
    ap.add_argument("--sustain-seconds", type=float, default=2.0, help="length of the second, sustained timed region (value_sustained)")
    ap.add_argument("--no-extra", action="store_true", help="skip the short lines of the other BASELINE configs / id distributions (extra)")
    ap.add_argument("--ref-adam", default="tf_dense", choices=["tf_dense", "lazy"],
                    help="Adam semantics of the CPU reference arm: Keras' dense passes (what the reference runs) or lazy")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# synthetic Criteo-shaped data: seed 4 (+rank), uniform or Zipf(1.05)+2% id 0
# ---------------------------------------------------------------------------------------------------

def synth_batches(n, B, V, dist, seed, pin, table_rows=None):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        if table_rows is not None and dist == "uniform":       # uniform inside every table's own cardinality
            cards = torch.tensor(table_rows, dtype=torch.float64)
            cat = (torch.rand(B, F_CAT, generator=g, dtype=torch.float64) * cards[None]).to(torch.int64)
            cat = torch.minimum(cat, (cards[None] - 1).to(torch.int64))
        elif table_rows is not None:     # Zipf(1.05) folded into every table's cardinality + 2 % forced id 0
            u = torch.rand(B, F_CAT, generator=g, dtype=torch.float64).clamp_(min=1e-12)
            cat = u.pow(-1.0 / 0.05).clamp_(max=2.0 ** 62).to(torch.int64) % torch.tensor(table_rows, dtype=torch.int64)[None]
            cat[torch.rand(B, F_CAT, generator=g) < 0.02] = 0
        elif dist == "uniform":
            cat = torch.randint(0, V, (B, F_CAT), generator=g, dtype=torch.int64)
        else:
            u = torch.rand(B, F_CAT, generator=g, dtype=torch.float64).clamp_(min=1e-12)
            cat = (u.pow(-1.0 / 0.05).clamp_(max=2.0 ** 62).to(torch.int64) % V)      # Pareto tail ~ Zipf(1.05), folded mod V
            cat[torch.rand(B, F_CAT, generator=g) < 0.02] = 0                          # OOV -> 0 (ctr/tfrecord_io.py:61-64)
        dense = torch.log1p(torch.randint(0, 1000, (B, F_INT), generator=g).float())  # ctr/tfrecord_io.py:48-53
        label = (torch.rand(B, generator=g) < 0.25).to(torch.int64)
        batch = (cat, dense, label)
        if pin:
            batch = tuple(t.pin_memory() for t in batch)
        out.append(batch)
    return out


# ---------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc, self.thread = [], None, None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # under load = the upper half of the samples (the sampler also sees idle gaps)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU reference arm (torch-CPU restatement of the reference; oracle/torch_cpu_ref.py)
# ---------------------------------------------------------------------------------------------------

def time_cpu_reference(args, steps, warmup, budget_s=None):
    from oracle.torch_cpu_ref import DLRMRef
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.cpu_sample_batch or args.batch
    ref = DLRMRef(BOTTOM[:-1] + [args.emb_dim], TOP, args.emb_dim, args.rows_per_table, num_tables=args.tables, seed=4,
                  adam=args.ref_adam)
    batches = synth_batches(2, B, args.rows_per_table, args.dist, seed=4, pin=False)
    t_start = time.perf_counter()
    for i in range(warmup):
        ref.train_step(*batches[i % 2])
        if budget_s and time.perf_counter() - t_start > budget_s / 2:
            break
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        ref.train_step(*batches[i % 2])
        times.append(time.perf_counter() - t0)
        if budget_s and time.perf_counter() - t_start > budget_s:
            break
    sec = sum(times) / len(times)
    sample = (f"{len(times)} steps of B={B} ({'the GPU batch' if B == args.batch else f'1/{max(1, args.batch // B)} of the GPU batch'}) on the same DLRM "
              f"({args.tables}x{args.rows_per_table}-row tables, D={args.emb_dim}); torch-CPU restatement of ctr/model.py + "
              f"Keras Adam ({args.ref_adam}), TensorFlow absent")
    return dict(value=B / sec, unit=UNIT, cores=cores, kind="port", sample=sample), sec * 1e3, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms, done = time_cpu_reference(args, max(1, args.steps), max(0, args.warmup), budget_s=240)
    cfg = workload_config(args, args.gpus)
    if args.gpus > 1:
        cfg["note"] = ("one replica's step on this host's cores, BASELINE config 2 tables: the b200 arm at N > 1 runs config 3, whose 187.8 M rows "
                       "(48 GB of fp32 table + 96 GB of Keras Adam state, every row moved every step) do not fit a bounded CPU sample; same model, "
                       "same batch per replica")
        cfg["parallelism"] = "one replica on the host CPU"
    line = dict(metric=METRIC, value=base["value"], unit=UNIT, n_gpus=args.gpus, steps=done, warmup=args.warmup, ms_per_step=ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=cfg, cpu_baseline=base,
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    if getattr(args, "criteo_tb", False):
        what = (f"DLRM Criteo-Terabyte-shape (BASELINE config 3): emb dim {args.emb_dim}, batch {args.batch} per GPU, 26 tables with the capped "
                f"Criteo-Terabyte cardinalities ({sum(CRITEO_TB_ROWS)} rows, 3 .. 39979771 per table), dot interaction, Adam")
    else:
        what = (f"DLRM Criteo-shape (BASELINE config 2): emb dim {args.emb_dim}, batch {args.batch} per GPU, "
                f"{args.tables} table(s) x {args.rows_per_table} rows, dot interaction, Adam")
    return dict(workload=what,
                global_batch=args.batch * world, tables=args.tables, rows_per_table=args.rows_per_table, emb_dim=args.emb_dim,
                bottom_mlp=BOTTOM[:-1] + [args.emb_dim], top_mlp=TOP, ids=args.dist, mlp_dtype=args.mlp_dtype,
                mlp_backend=("tcgen05 Dense kernels (csrc/mlp.cu)" if getattr(args, "mlp_backend", "tcgen05") == "tcgen05" else "torch / cuBLASLt")
                if args.mlp_dtype == "bf16" else "torch fp32",
                sparse_optimizer="adam_lazy (Keras Adam formula on the touched rows; the CPU arm runs Keras' dense passes, tf_dense: "
                                 "the two agree on step 1 from zero state and differ afterwards, tests/test_gpu_kernels.py)",
                mlp_evaluation="collapsed affine map per tower (opt-in)" if getattr(args, "collapse_mlp", False) and world == 1
                else "layer by layer",
                parallelism="single GPU" if world == 1 else (
                    f"row-wise sharded tables x{world}, rows and gradient rows read by the kernels over NVLink peer memory + data-parallel MLPs"
                    + (f"; the {sum(1 for r in CRITEO_TB_ROWS if r <= args.replicate_small)} tables of <= {args.replicate_small} rows replicated "
                       "on every rank (local reads, one all-gathered update)" if getattr(args, "criteo_tb", False) and args.replicate_small > 0 else "")
                    if args.exchange == "p2p" else f"{args.sharding}-wise sharded tables x{world}, NCCL all-to-all + data-parallel MLPs"),
                l2="inputs larger than L2: tables %.1f GB, ring of %d batches, 436 MB gradient tensor per step" % (
                    args.tables * args.rows_per_table * args.emb_dim * 4 / 1e9, args.ring))


# ---------------------------------------------------------------------------------------------------
# short lines for the other BASELINE configs and id distributions (N = 1; the `extra` object of the JSON line)
# ---------------------------------------------------------------------------------------------------

ESMM_VOCAB = [238635, 98, 14, 3, 8, 4, 4, 3, 5, 467298, 6929, 263942, 106399, 5888, 104830, 51878, 37148, 4]      # esmm/train.py:197-215


def _time_loop(fn, iters, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def extra_lines(args, dev, model, graphed_step, peak):
    """SURVEY §8d asks for both id distributions and for T = 1 as well as T = 26; BASELINE configs 1, 4 and 5 get one short
    line each.  Everything here is a few dozen launches; inputs larger than L2 wherever the config is."""
    from recommender_b200 import ops
    from recommender_b200.graph import GraphedTrainStep
    from recommender_b200.model import DLRM, DeepFM, bce_clipped, bce_logits
    from recommender_b200.ops import GradSource, LookupGroup
    from recommender_b200.optimizers import Adam
    D, V, B = args.emb_dim, args.rows_per_table, args.batch
    cd = torch.bfloat16 if args.mlp_dtype == "bf16" else None
    out = {}

    def dlrm_line(tables, dist, step_fn=None):
        if step_fn is None:
            g = torch.Generator(device=dev).manual_seed(4)
            m = DLRM(BOTTOM[:-1] + [D], TOP, D, V, F_CAT, F_INT, num_tables=tables, device=dev, compute_dtype=cd, generator=g)
            for mod in m.modules():
                if hasattr(mod, "backend"):
                    mod.backend = args.mlp_backend
            first = [tuple(t.to(dev) for t in b) for b in synth_batches(1, B, V, dist, seed=4, pin=False)]
            step_fn = GraphedTrainStep(m, Adam(), bce_clipped, first[0], warmup=3).step
        batches = [tuple(t.to(dev) for t in b) for b in synth_batches(4, B, V, dist, seed=5, pin=False)]
        ms = _time_loop(lambda i: step_fn(batches[i % 4]), 20)
        rows0 = batches[0][0] if tables == 1 else batches[0][0] + (torch.arange(tables, device=dev) * V)[None]
        return dict(ms_per_step=ms, samples_per_s=B / ms * 1e3, unique_rows=int(torch.unique(rows0).numel()), tables=tables, ids=dist)

    if graphed_step is not None and args.dist != "zipf":
        out[f"dlrm_cfg2_T{args.tables}_zipf"] = dlrm_line(args.tables, "zipf", graphed_step)      # the headline model, skewed ids
    other_t = 1 if args.tables == 26 else 26
    if other_t == 1:      # one shared 1M-row table: the reference's own layout (ctr/model.py:42)
        out["dlrm_cfg2_T1_uniform"] = dlrm_line(1, "uniform")
        out["dlrm_cfg2_T1_zipf"] = dlrm_line(1, "zipf")

    # BASELINE config 1: DeepFM through ctr/train.py's constants, B = 1024, D = 16, one shared 1M-row table, MLP 512-256-1
    class _Logits(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, inputs):
            return self.m.logits(inputs)

    g = torch.Generator(device=dev).manual_seed(4)
    B1 = 1024
    fm = _Logits(DeepFM(16, 1_000_000, 13, 26, [512, 256, 1], device=dev, generator=g, compute_dtype=cd))
    ring = [tuple(t.to(dev) for t in b) for b in synth_batches(4, B1, 1_000_000, "uniform", seed=4, pin=False)]
    ring = [(c, d, l.float()) for c, d, l in ring]
    gs = GraphedTrainStep(fm, Adam(), bce_logits, ring[0], warmup=3)
    ms = _time_loop(lambda i: gs.step(ring[i % 4]), 50)
    out["deepfm_cfg1_B1024"] = dict(ms_per_step=ms, samples_per_s=B1 / ms * 1e3, launch_mode="cuda_graph",
                                    note="launch-latency-bound on a B200 (SURVEY §8d): 3.6 MB of gather per step")
    del fm, gs

    # BASELINE config 4: masked mean over a behaviour history (dien/layers.py:5-17), L = 100, D = 32, item table of 400k rows
    Lh, Dh, Vh = 100, 32, 400_000
    g = torch.Generator(device=dev).manual_seed(4)
    wt = torch.empty(Vh, Dh, device=dev).uniform_(-0.05, 0.05, generator=g)
    lens = torch.randint(1, Lh + 1, (B, 1), device=dev, generator=g)
    hist = torch.randint(1, Vh, (B, Lh), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
    hist = torch.where(torch.arange(Lh, device=dev)[None] < lens, hist, torch.zeros_like(hist))     # trailing pads (dien/data_loader.py:44)
    pout = torch.empty(B, Dh, device=dev)
    ms_f = _time_loop(lambda i: ops.bag_pool_fwd(wt, hist, "masked_mean", out=pout, out_stride=Dh), 20)
    valid = int((hist != 0).sum())
    fbytes = valid * Dh * 4 + B * Lh * 4 + B * Dh * 4
    dpool = torch.randn(B, Dh, device=dev, generator=g) * 1e-3
    cnt = (hist != 0).sum(1).float()
    mh, vh = torch.zeros_like(wt), torch.zeros_like(wt)
    stp = [0]

    def hupd(i):
        stp[0] += 1
        grp = LookupGroup(hist, Lh, GradSource.per_bag([dpool], scale="masked_mean", mask_idx=hist, count=cnt))
        ops.sparse_bwd_update(wt, mh, vh, [grp], optimizer="adam_lazy", step=stp[0])
    ms_b = _time_loop(hupd, 10)
    out["dien_cfg4_masked_mean"] = dict(batch=B, history=Lh, emb_dim=Dh, valid_positions=valid, fwd_ms=ms_f, fwd_algorithmic_bytes=fbytes,
                                        fwd_gbs=fbytes / ms_f / 1e6, fwd_frac=fbytes / ms_f / 1e6 / peak, bwd_update_ms=ms_b,
                                        samples_per_s_fwd_bwd=B / (ms_f + ms_b) * 1e3)
    del wt, mh, vh, hist

    # SURVEY §8f rank 4 at BASELINE config 4's shape: DIN attention pooling (dien/layers.py:34-59) over a behaviour history of
    # 100 positions, item + category tables of D = 32 each (E = 64), ~50 % trailing pads.  One step = target + history lookups,
    # the 4E -> 80 -> 40 -> 1 attention MLP on the valid positions (tcgen05 Dense kernels), weighted pooling, the whole
    # backward and one sparse Adam update per table.
    from recommender_b200.din import DIN
    Vi, Vc, Dd = 400_000, 2_000, 32
    g = torch.Generator(device=dev).manual_seed(4)
    din = DIN(Vi, Dd, Vc, Dd, device=dev, generator=g)
    lens = torch.randint(1, Lh + 1, (B, 1), device=dev, generator=g)
    hi = torch.randint(1, Vi, (B, Lh), device=dev, generator=g)
    hi = torch.where(torch.arange(Lh, device=dev)[None] < lens, hi, torch.zeros_like(hi))
    hc = torch.where(hi != 0, hi % (Vc - 1) + 1, torch.zeros_like(hi))
    ti = torch.randint(1, Vi, (B, 1), device=dev, generator=g)
    din_in = dict(target_item=ti, target_cat=ti % (Vc - 1) + 1, pos_his_item=hi, pos_his_cat=hc)
    d_out = torch.randn(B, 4 * Dd, device=dev, generator=g) * 1e-3
    din_opt = Adam()
    t_fwd = [0.0]

    def din_step(i):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = din(din_in)
        e1.record()
        o.backward(d_out)
        din_opt.apply_gradients(din)
        t_fwd[0] = (e0, e1)
    ms_d = _time_loop(din_step, 8)
    torch.cuda.synchronize()
    valid_d = int((hi != 0).sum())
    flop = 2 * valid_d * (8 * Dd * 80 + 80 * 40 + 40) * 3      # forward + input-gradient + weight-gradient products
    # the gather kernels of the unit alone, against the HBM roofline (algorithmic bytes per valid position: DESIGN §4)
    E_d = 2 * Dd
    hist = ops.DinHistory(din.item_embedding.embeddings, hi, din.cat_embedding.embeddings, hc)
    off_d = ops.din_offsets(hist)
    tgt_d = torch.randn(B, E_d, device=dev, generator=g) * 0.05
    w_d = torch.randn(valid_d, device=dev, generator=g)
    drep_d = torch.randn(B, E_d, device=dev, generator=g) * 1e-3
    dX_d = (torch.randn(valid_d, 4 * E_d, device=dev, generator=g) * 1e-3).to(torch.bfloat16)
    din_kernels = {}
    for name, fn, nbytes in (
            ("build_features", lambda i: ops.din_build_features(hist, tgt_d, off_d, valid_d, 4 * E_d), valid_d * (E_d * 4 + 4 * E_d * 2) + B * Lh * 16),
            ("pool_fwd", lambda i: ops.din_pool_fwd(hist, off_d, w_d), valid_d * (E_d * 4 + 4) + B * Lh * 16),
            ("pool_bwd_weights", lambda i: ops.din_pool_bwd_weights(hist, off_d, drep_d, valid_d), valid_d * (E_d * 4 + 4) + B * Lh * 16),
            ("feature_bwd", lambda i: ops.din_feature_bwd(hist, tgt_d, off_d, dX_d, w_d, drep_d), valid_d * (E_d * 4 * 2 + 4 * E_d * 2) + B * Lh * 16)):
        ms_k = _time_loop(fn, 5, warm=2)
        din_kernels[name] = dict(ms=ms_k, algorithmic_bytes=int(nbytes), gbs=nbytes / ms_k / 1e6, frac=nbytes / ms_k / 1e6 / peak)
    # the reference's own arithmetic on the host cores (oracle.local_activation_unit forward + backward, numpy): a bounded sample
    import numpy as _np
    from oracle import ctr_oracle as _O
    rng_c = _np.random.default_rng(4)
    Bc = 256
    his_c = rng_c.normal(0, 0.05, size=(Bc, Lh, E_d)).astype(_np.float32)
    tgt_c = rng_c.normal(0, 0.05, size=(Bc, E_d)).astype(_np.float32)
    mask_c = _np.arange(Lh)[None] < rng_c.integers(1, Lh + 1, size=(Bc, 1))
    lay_c = [(rng_c.normal(0, 0.1, size=(a, b)).astype(_np.float32), _np.zeros(b, _np.float32)) for a, b in ((4 * E_d, 80), (80, 40), (40, 1))]
    t_c = time.perf_counter()
    rep_c, cache_c = _O.local_activation_unit(tgt_c, his_c, mask_c, lay_c)
    _O.local_activation_unit_backward(cache_c, _np.ones_like(rep_c))
    cpu_s = time.perf_counter() - t_c
    out["din_cfg4_attention"] = dict(batch=B, history=Lh, emb_dim=2 * Dd, valid_positions=valid_d, ms_per_step=ms_d, kernels=din_kernels,
                                     cpu_baseline=dict(value=int(mask_c.sum()) / cpu_s, unit="valid positions/s", cores=os.cpu_count(), kind="port",
                                                       sample=f"oracle.local_activation_unit forward + backward (numpy, fp32) on {Bc} samples x {Lh} positions"),
                                     fwd_ms=t_fwd[0][0].elapsed_time(t_fwd[0][1]), samples_per_s=B / ms_d * 1e3,
                                     positions_per_s=valid_d / ms_d * 1e3, attention_mlp_tflops=flop / ms_d / 1e9,
                                     note="eager (one host read-back of the valid-position count per step sizes the GEMMs)")
    del din, din_opt, hi, hc

    # BASELINE config 5: 26 tables (vocab sizes cycled from esmm/train.py:197-215), D = 32, bag size 1, gradients of 2 (ESMM) and
    # 10 (MMOE) consumers added inside the scatter (esmm/esmm.py:15-24).  The tables live back to back in ONE tensor
    # (layers.Embedding(table_rows=...)): the 26 lookups + concat are one rb_gather_fwd over [B, 26] ids with per-field row
    # offsets, and the 26 sparse Adam applies one rb_sparse_bwd_update whose gradient source lists the consumers.
    D5, T5 = 32, 26
    vocab = [ESMM_VOCAB[k % len(ESMM_VOCAB)] for k in range(T5)]
    g = torch.Generator(device=dev).manual_seed(4)
    rows5 = sum(vocab)
    tab = torch.empty(rows5, D5, device=dev).uniform_(-0.05, 0.05, generator=g)
    m5, v5 = torch.zeros_like(tab), torch.zeros_like(tab)
    off5 = torch.tensor([0] + vocab[:-1], dtype=torch.int64, device=dev).cumsum(0)
    idx5 = torch.stack([torch.randint(0, v, (B,), device=dev, generator=g, dtype=torch.int64) for v in vocab], dim=1).to(torch.int32).contiguous()
    width = D5 * T5
    emb = torch.empty(B, T5, D5, device=dev)
    ms_f5 = _time_loop(lambda i: ops.gather_fwd(tab, idx5, L=T5, field_row_offset=off5, out=emb, out_stride=D5), 20)
    gbytes = B * T5 * (D5 * 4 * 2 + 4)
    res5 = dict(batch=B, tables=T5, emb_dim=D5, vocab_min=min(vocab), vocab_max=max(vocab), rows=rows5, gather_concat_ms=ms_f5,
                gather_algorithmic_bytes=gbytes, gather_gbs=gbytes / ms_f5 / 1e6, gather_frac=gbytes / ms_f5 / 1e6 / peak)
    uniq = int(torch.unique(idx5.to(torch.int64) + off5[None]).numel())
    for nc in (2, 10):
        cons = [torch.randn(B, width, device=dev, generator=g) * 1e-3 for _ in range(nc)]
        stp5 = [0]

        def upd5(i):
            stp5[0] += 1
            src = GradSource(cons, [width] * nc, [D5] * nc)
            ops.sparse_bwd_update(tab, m5, v5, [LookupGroup(idx5, T5, src, field_row_offset=off5)], optimizer="adam_lazy", step=stp5[0])
        ms_u = _time_loop(upd5, 10, warm=2)
        ubytes = B * T5 * (nc * D5 * 4 + 4) + uniq * D5 * 4 * 6
        res5[f"scatter_adam_{nc}_consumers_ms"] = ms_u
        res5[f"scatter_adam_{nc}_consumers_algorithmic_bytes"] = ubytes
        res5[f"scatter_adam_{nc}_consumers_gbs"] = ubytes / ms_u / 1e6
        del cons
    res5["unique_rows"] = uniq
    out["esmm_cfg5_multi_table"] = res5
    return out


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------

class CallTimer:
    """CUDA-event timing of every C-ABI call on the launching (current) stream."""

    def __init__(self, ops, names):
        self.ops, self.names, self.records, self.enabled = ops, names, {n: [] for n in names}, False
        self._orig = {}

    def install(self):
        for n in self.names:
            self._orig[n] = getattr(self.ops, n)
            setattr(self.ops, n, self._wrap(n, self._orig[n]))

    def install_on(self, obj, methods, prefix):
        """Also time bound methods of `obj` (the peer-memory embedding calls the library directly)."""
        for m in methods:
            name = prefix + m
            self.records[name] = []
            setattr(obj, m, self._wrap(name, getattr(obj, m)))

    def _wrap(self, name, fn):
        def timed(*a, **k):
            if not self.enabled:
                return fn(*a, **k)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = fn(*a, **k)
            e.record()
            self.records[name].append((s, e))
            return out
        return timed

    def summary(self):
        return {n: (sum(s.elapsed_time(e) for s, e in r) / len(r), len(r)) for n, r in self.records.items() if r}


def peer_memory_usable(args, dev) -> bool:
    """Pre-flight for the peer-memory path: every rank tries a tiny symmetric-memory allocation + rendezvous; unless ALL
    succeed, all ranks fall back to the NCCL all-to-all path together (a one-sided failure would deadlock later)."""
    import torch.distributed as dist
    if args.exchange != "p2p":
        return False
    ok = 1
    try:
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(64, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD)
        hdl.barrier(channel=0)
        torch.cuda.synchronize()
    except Exception as e:                                        # noqa: BLE001 - any failure means "not usable here"
        print(f"# peer-memory pre-flight failed on rank {dist.get_rank()}: {type(e).__name__}: {e}", file=sys.stderr)
        ok = 0
    flag = torch.tensor([ok], device=dev, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    usable = bool(int(flag.item()))
    if not usable:
        args.exchange = "nccl"
    return usable


def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    from recommender_b200 import build, ops
    if local_rank == 0:
        build.build()              # a no-op when the in-tree .so matches the sources
    if world > 1:
        dist.barrier()
    from recommender_b200.model import DLRM, bce_clipped
    from recommender_b200.optimizers import Adam

    D, V, T, B = args.emb_dim, args.rows_per_table, args.tables, args.batch
    if world > 1 and args.exchange == "p2p" and not args.config2_sharded and T == 26:
        args.criteo_tb = True          # BASELINE's multi-GPU workload (config 3) is what an N > 1 run measures
    cap_factor = args.capacity_factor or (2.0 if args.criteo_tb else 1.25)
    cd = torch.bfloat16 if args.mlp_dtype == "bf16" else None
    gen = torch.Generator(device=dev).manual_seed(4)
    if world == 1:
        model = DLRM(BOTTOM[:-1] + [D], TOP, D, V, F_CAT, F_INT, num_tables=T, device=dev, compute_dtype=cd, generator=gen,
                     collapse_linear=args.collapse_mlp)
    elif peer_memory_usable(args, dev):
        from recommender_b200.p2p import P2PShardedDLRM
        model = P2PShardedDLRM(BOTTOM[:-1] + [D], TOP, D, V, F_CAT, F_INT, num_tables=T, device=dev, compute_dtype=cd, generator=gen,
                               table_rows=CRITEO_TB_ROWS if args.criteo_tb else None, capacity_factor=cap_factor,
                               replicate_rows_upto=args.replicate_small if args.criteo_tb else 0)
    else:
        from recommender_b200.sharded import ShardedDLRM
        args.criteo_tb = False           # the NCCL all-to-all fallback runs config 2 sharded
        model = ShardedDLRM(BOTTOM[:-1] + [D], TOP, D, V, F_CAT, F_INT, num_tables=T, device=dev, compute_dtype=cd, generator=gen,
                            sharding=args.sharding)
    for m in model.modules():
        if hasattr(m, "backend"):
            m.backend = args.mlp_backend
    opt = Adam()

    if args.criteo_tb and not (world > 1 and args.exchange == "p2p"):
        raise SystemExit("--criteo-tb needs the peer-memory sharded path (N > 1)")
    host = synth_batches(args.ring, B, V, args.dist, seed=4 + rank, pin=True, table_rows=CRITEO_TB_ROWS if args.criteo_tb else None)
    resident = [tuple(t.to(dev) for t in b) for b in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def train_step(batch):
        cat, dense, label = batch
        prob = model({"cat_features": cat, "int_features": dense})
        loss = bce_clipped(prob, label)
        loss.backward()
        opt.apply_gradients(model)
        return loss.detach()       # no reference to the autograd graph survives the step (CUDA-graph capture needs that)

    timer = CallTimer(ops, ["dot_interaction_fwd", "dot_interaction_bwd", "sparse_bwd_update", "sparse_bwd_prepare", "sparse_bwd_apply",
                            "gather_fwd", "bucket_by_owner", "dense_opt_step", "colsum"])
    timer.install()
    if world > 1 and args.exchange == "p2p":
        # every one of these runs with the stream it launches on as torch's current stream (route / collect_and_sort / _apply_rows
        # inside `with torch.cuda.stream(side)`), so the bracketing events sit on the launching stream; the rendezvous barriers
        # are outside the brackets
        timer.install_on(model.embedding_layer, ["route", "collect_and_sort", "_interaction_fwd", "_interaction_bwd", "_apply_rows"], "p2p.")
        if getattr(model.embedding_layer, "_small", None):
            timer.install_on(model.embedding_layer, ["_small_reduce", "_small_exchange", "_small_update"], "p2p.")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(args.warmup, 3)):
        loss = train_step(resident[i % args.ring])
    float(loss.item())
    ops.check_oob(dev)
    if hasattr(model.embedding_layer, "check_overflow"):
        model.embedding_layer.check_overflow()       # static per-owner capacity of the peer-memory path

    # ---- eager pass: every C-ABI call bracketed by CUDA events on its launching stream (kernel table, roofline) ----
    k_eager = min(args.steps, 10)
    barrier()
    launches0 = ops.kernel_launches()
    timer.enabled = True
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(k_eager):
        loss = train_step(resident[i % args.ring])
    g1.record()
    barrier()
    timer.enabled = False
    launches_per_step = (ops.kernel_launches() - launches0) / k_eager
    eager_ms_step = max_over_ranks(g0.elapsed_time(g1)) / k_eager

    # ---- the step as one CUDA graph (single GPU; the sharded step has host-side exchange counts) --------------
    use_graph = (world == 1 or args.exchange == "p2p") and not args.no_graph     # the NCCL all-to-all path has host-side counts
    if use_graph:
        from recommender_b200.graph import GraphedTrainStep
        del loss
        graphed = GraphedTrainStep(model, opt, bce_clipped, resident[0], warmup=max(args.warmup, 3))
        if not args.no_bind_inputs:
            graphed.bind_inputs(resident)        # one graph per resident batch: replays read the batch in place (no copy into static inputs)
        train_step = graphed.step
        for i in range(3):
            loss = train_step(resident[i % args.ring])
        float(loss.item())

    # ---- value: inputs resident in HBM ---------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = train_step(resident[i % args.ring])
    e1.record()
    barrier()
    launches = int(round(launches_per_step * args.steps))      # the graph replays exactly the eager step's kernels
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)
    final_loss = float(loss.item())
    if hasattr(model.embedding_layer, "check_overflow"):
        model.embedding_layer.check_overflow()

    # ---- the same loop for >= --sustain-seconds: the number a long job sees (clocks settle under the power cap) ----------
    sustained = None
    if args.sustain_seconds > 0:
        n_sus = max(args.steps, int(args.sustain_seconds * 1e3 / ms_step) + 1)
        sampler2 = ClockSampler(local_rank)
        if rank == 0:
            sampler2.start()
        barrier()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        for i in range(n_sus):
            loss = train_step(resident[i % args.ring])
        u1.record()
        barrier()
        sus_ms = max_over_ranks(u0.elapsed_time(u1))
        sustained = dict(value=B * world * n_sus / (sus_ms / 1e3), unit=UNIT, steps=n_sus, seconds=sus_ms / 1e3, ms_per_step=sus_ms / n_sus,
                         clocks=sampler2.stop() if rank == 0 else None)

    # ---- N > 1: real-rank parity of the sharded path against the unsharded model (small problem, untimed) ---------------
    parity = None
    if world > 1 and args.exchange == "p2p":
        from recommender_b200 import p2p_selfcheck
        parity = {}
        for tag, dt in (("fp32_towers", None), ("bf16_towers", torch.bfloat16)):
            try:
                parity[tag] = p2p_selfcheck.run(dev, compute_dtype=dt)
                if args.criteo_tb and args.replicate_small > 0 and tag == "bf16_towers":
                    # the same check on unequal tables with the small ones replicated (the path config 3 takes)
                    parity["bf16_towers_replicated_small_tables"] = p2p_selfcheck.run(
                        dev, compute_dtype=dt, table_rows=[5000] * 20 + [3, 14, 63, 155, 976, 2208], replicate_rows_upto=4096)
            except Exception as e:                               # noqa: BLE001 - reported in the line, never hidden
                parity[tag] = dict(error=f"{type(e).__name__}: {e}")
                break

    # ---- e2e: pinned host buffers -> H2D copies, loss read back, all inside the timed region ---------------
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream()
        stage = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
        if use_graph and not args.no_bind_inputs:
            graphed.bind_inputs(stage)           # the two staging buffers the H2D copies land in are graph inputs themselves
        h2d_done = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        loss_host = torch.zeros(args.steps + 3, dtype=torch.float32).pin_memory()

        def prefetch(i):
            buf = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[buf])
                for dst, src in zip(stage[buf], host[i % args.ring]):
                    dst.copy_(src, non_blocking=True)
                h2d_done[buf].record(copy_stream)

        host_loop_ms = [0.0]

        def e2e_loop(n, base, h2d=True, d2h=True):
            for ev in free:
                ev.record(main)
            prefetch(0)
            t_host = time.perf_counter()
            for i in range(n):
                buf = i % 2
                if i + 1 < n and (h2d or i == 0):
                    prefetch(i + 1)
                if h2d or i < 2:
                    main.wait_event(h2d_done[buf])
                # the step's result, read every step: GraphedTrainStep.step(loss_out=) copies it to the host on its own stream (on the
                # main stream the 4-byte copy sits between two graph launches: ~40 us per step, r2_60 -> r2_62)
                if use_graph and d2h:
                    loss = train_step(stage[buf], loss_out=loss_host[base + i])
                else:
                    loss = train_step(stage[buf])
                    if d2h:
                        loss_host[base + i].copy_(loss.detach(), non_blocking=True)
                free[buf].record(main)
            host_loop_ms[0] = (time.perf_counter() - t_host) * 1e3 / max(n, 1)   # host time to ENQUEUE one step (before the sync)
            torch.cuda.synchronize()

        e2e_loop(3, 0)
        barrier()
        t0 = time.perf_counter()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        e2e_loop(args.steps, 3)
        s1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_ms = max_over_ranks(max(s0.elapsed_time(s1), 0.0)) / args.steps
        probe = None
        if os.environ.get("RB_E2E_PROBE", "0") == "1":       # diagnostic: where the distance to `value` comes from (untimed extras)
            probe = {}
            for tag, kw in (("full", {}), ("no_d2h", dict(d2h=False)), ("no_h2d", dict(h2d=False)), ("neither", dict(h2d=False, d2h=False))):
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                p0.record()
                e2e_loop(args.steps, 3, **kw)
                p1.record()
                torch.cuda.synchronize()
                probe[tag] = p0.elapsed_time(p1) / args.steps
        e2e = dict(value=B * world / (e2e_ms / 1e3), unit=UNIT, h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=4,
                   ms_per_step=e2e_ms, wall_ms_per_step=wall_ms / args.steps, host_enqueue_ms_per_step=host_loop_ms[0],
                   **({"probe_ms_per_step": probe} if probe else {}),
                   api=("GraphedTrainStep.step (one CUDA graph: DLRM.__call__ + bce_clipped + backward + Adam.apply_gradients)"
                        if use_graph else "DLRM.__call__ + bce_clipped + backward + Adam.apply_gradients")
                       + " on batches staged from pinned host memory")

    # ---- roofline of the dominant C-ABI call -------------------------------------------------------------
    calls = timer.summary()
    N = B * F_CAT
    Fp = F_CAT + 1
    cat0 = resident[0][0]
    rows0 = cat0 if T == 1 else cat0 + (torch.arange(T, device=dev) * V)[None]
    U = int(torch.unique(rows0).numel())
    # bytes of one interaction row as it crosses HBM: fp32 [F'^2+D], or bf16 padded to 8 columns for the bf16 top MLP
    width = Fp * Fp + D
    row_bytes = ((width + 7) // 8 * 8) * 2 if cd == torch.bfloat16 else width * 4
    algo = {
        # N*idxB + N*D*4 (rows) + B*D*4 (dense vec) + B*row (out)
        "dot_interaction_fwd": N * 8 + N * D * 4 + B * D * 4 + B * row_bytes,
        # dOut + re-gathered rows + ids + dense vec in; dE + d_dense out
        "dot_interaction_bwd": B * row_bytes + N * D * 4 + N * 8 + B * D * 4 + N * D * 4 + B * D * 4,
        # N*D*4 (dE) + N*idxB + U*D*4*6 (read+write of var, m, v)
        "sparse_bwd_update": N * D * 4 + N * 8 + U * D * 4 * 6,
        # the same bytes when keys + radix sort ran ahead on the side stream (rb_sparse_bwd_prepare, overlapped with the forward)
        "sparse_bwd_apply": N * D * 4 + N * 8 + U * D * 4 * 6,
    }
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = {}
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    kernels = {}
    for name, (ms, cnt) in calls.items():
        if name in algo and world == 1:
            gbs = algo[name] / (ms * 1e-3) / 1e9
            kernels[name] = dict(ms=ms, calls_per_step=cnt / k_eager, algorithmic_bytes=algo[name], achieved_gbs=gbs,
                                 frac=gbs / peak, share_of_step=ms * cnt / k_eager / ms_step)
        else:
            kernels[name] = dict(ms=ms, calls_per_step=cnt / k_eager, share_of_step=ms * cnt / k_eager / ms_step)
    for name in ("dense_opt_step", "colsum"):
        if name in kernels:
            # these small launches are bracketed on the main stream while the row update fills the SMs from its side stream: the
            # event time is mostly time QUEUED for an SM, not kernel time (dense_opt_kernel alone: ~13 us; profiles/r2_20_timeline.txt)
            kernels[name]["note"] = "event time on the main stream beside the side-stream row update: includes queueing for SMs; ms is not kernel time"
            kernels[name].pop("share_of_step", None)
    roofline = None
    timed = {k: v for k, v in kernels.items() if "achieved_gbs" in v}
    if timed:
        top = max(timed, key=lambda k: timed[k]["ms"] * timed[k]["calls_per_step"])
        roofline = dict(bound="hbm", kernel=top, achieved=timed[top]["achieved_gbs"], peak=peak, unit="GB/s", frac=timed[top]["frac"],
                        traffic=traffic.get(top), peak_source=peak_src, unique_rows=U,
                        frac_of_nominal_8tbs=timed[top]["achieved_gbs"] / 8000.0,
                        note="achieved = algorithmic bytes of the whole C-ABI call / its CUDA-event time on the launching stream. "
                             "sparse_bwd_apply = segmented reduction + fused Adam row update; its keys + radix sort "
                             "(sparse_bwd_prepare, overhead, not algorithmic bytes) run on a side stream during the forward; "
                             "sparse_bwd_update = both phases in one call")
    if world > 1 and args.exchange == "p2p":
        # Bytes each rank pulls over NVLink per step: (G-1)/G of the rows (bf16 shadow, 2 B per element) inside the forward kernel,
        # (G-1)/G of the fp32 gradient rows inside the owner-side apply.  Both calls are bracketed by events on the stream they run on.
        remote = (world - 1) / world
        nv_peak, nv_measured = 900.0, 770.0      # nominal per direction per GPU; peer-copy rate measured on this pool (B200_PROFILING.md)
        nv = {}
        # lookups of tables replicated on every rank never leave the GPU: only the sharded tables' rows cross NVLink
        n_small = len(getattr(model.embedding_layer, "_small", []) or [])
        N_sharded = B * (F_CAT - n_small)
        for name, nbytes in (("p2p._interaction_fwd", N_sharded * D * 2 * remote), ("p2p._apply_rows", N_sharded * D * 4 * remote)):
            if name in kernels:
                gbs = nbytes / (kernels[name]["ms"] * 1e-3) / 1e9
                kernels[name].update(nvlink_bytes=int(nbytes), nvlink_gbs=gbs, nvlink_frac=gbs / nv_peak)
                nv[name] = kernels[name]
        if nv:
            top = max(nv, key=lambda k: nv[k]["ms"])
            roofline = dict(bound="nvlink", kernel=top, achieved=nv[top]["nvlink_gbs"], peak=nv_peak, unit="GB/s", frac=nv[top]["nvlink_frac"],
                            traffic=None, peak_source="nominal NVLink 5, 900 GB/s per direction per GPU",
                            frac_of_measured_peer_copy=nv[top]["nvlink_gbs"] / nv_measured,
                            replicated_tables=n_small,
                            note="achieved = bytes this rank's kernel reads from PEER memory over NVLink ((G-1)/G of the SHARDED tables' rows: bf16 shadow rows "
                                 "in the forward, fp32 gradient rows in the owner-side apply) / the CUDA-event time of that C call on its "
                                 "launching stream; the same kernels also move their local HBM share in that time")

    # ---- the second half of BASELINE.json's metric: embedding-gather HBM GB/s (un-pooled lookup, rb_gather_fwd) ----------
    gather = None
    if world == 1:
        table = model.embedding_layer.embeddings
        off = model.embedding_layer.row_offset_for(F_CAT)
        gout = torch.empty(B, F_CAT, D, dtype=torch.float32, device=dev)
        for i in range(3):
            ops.gather_fwd(table, resident[i % args.ring][0], L=F_CAT, field_row_offset=off, out=gout, out_stride=D)
        gs_, ge_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gs_.record()
        for i in range(20):
            ops.gather_fwd(table, resident[i % args.ring][0], L=F_CAT, field_row_offset=off, out=gout, out_stride=D)
        ge_.record()
        torch.cuda.synchronize()
        g_ms = gs_.elapsed_time(ge_) / 20
        g_bytes = N * D * 4 * 2 + N * 8        # rows + ids + output
        gather = dict(metric="embedding-gather HBM GB/s", ms=g_ms, algorithmic_bytes=g_bytes, achieved_gbs=g_bytes / (g_ms * 1e-3) / 1e9,
                      frac=g_bytes / (g_ms * 1e-3) / 1e9 / peak, peak=peak, lookups=N, emb_dim=D)
        del gout
        # the dominant call once more with nothing else on the GPU (in the step its events also count the time its CTAs queue behind
        # kernels of the other streams)
        if roofline is not None and roofline.get("kernel") == "sparse_bwd_apply" and "m" in model.embedding_layer.opt_state:
            from recommender_b200.ops import GradSource, LookupGroup
            st = model.embedding_layer.opt_state
            dE = torch.randn(B, F_CAT, D, device=dev) * 1e-3
            ws = ops.sparse_workspace(N, D, table.shape[0], dev)
            ts = []
            for i in range(8):
                grp = [LookupGroup(resident[i % args.ring][0], F_CAT, GradSource.per_position(dE, F_CAT), field_row_offset=off)]
                sel = ops.sparse_bwd_prepare(table.shape[0], D, grp, ws)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                ops.sparse_bwd_apply(table, st["m"], st["v"], grp, ws, sel, optimizer="adam_lazy", step=1000 + i)
                a1.record()
                torch.cuda.synchronize()
                ts.append(a0.elapsed_time(a1))
            ts.sort()
            alone_ms = ts[len(ts) // 2]
            alone_gbs = algo["sparse_bwd_apply"] / (alone_ms * 1e-3) / 1e9
            roofline.update(alone_ms=alone_ms, achieved_alone=alone_gbs, frac_alone=alone_gbs / peak, frac_alone_of_nominal_8tbs=alone_gbs / 8000.0)
            del dE, ws

    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            extra = extra_lines(args, dev, model, train_step if use_graph else None, peak)
        except Exception as e:                                   # noqa: BLE001 - the headline line must still be printed
            extra = dict(error=f"{type(e).__name__}: {e}")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del resident
        torch.cuda.empty_cache()
        cpu, _, _ = time_cpu_reference(args, steps=2, warmup=1, budget_s=90)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_step,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32 tables/optimizer, bf16 MMA operands (f32 accumulate)",
                    data="synthetic", config=workload_config(args, world), e2e=e2e, gpu_launches=int(launches), roofline=roofline,
                    kernels=kernels, cpu_baseline=cpu, clocks=clocks, final_loss=final_loss, host_cores=os.cpu_count(),
                    launch_mode="cuda_graph" if use_graph else "eager", eager_ms_per_step=eager_ms_step, embedding_gather=gather,
                    value_sustained=None if sustained is None else sustained["value"], sustained=sustained, parity=parity, extra=extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        # a captured graph holding NCCL kernels plus peer-mapped buffers makes interpreter teardown unreliable:
        # meet once more, flush, and leave without running destructors
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
