"""`python -m recommender_b200.train` — the entry point of /root/reference/ctr/train.py on the B200 path.

Same flags (ctr/train.py:13-19), same constants (:62-66, :73-82: 13 + 26 features, 1M-row shared table, embedding 16,
DLRM 512-256-64-16 / 512-256-1, DeepFM 512-256-1, Adam defaults, 3 epochs), same flow: read_tfrecord -> batch -> fit
with AUC / BinaryAccuracy, validation after every epoch, early stopping and best-weights checkpoint on val_auc
(:85-97).  What Keras `fit` does around the step is restated in ~100 lines of host code; the step itself is the CUDA
hot path, replayed as one CUDA graph (graph.GraphedTrainStep), and the batches come from tfrecord_io (record file or
raw Criteo text).  Not reproduced: TensorBoard logging, MirroredStrategy (one process drives one GPU here; the
multi-GPU step is P2PShardedDLRM, bench.py), and tf.data's 100-batch shuffle buffer (batches arrive in file order).
"""
from __future__ import annotations

import argparse
import os
import time
from typing import Dict, List, Optional

import torch

from . import tfrecord_io
from .graph import GraphedTrainStep
from .model import DLRM, DeepFM, bce_clipped
from .optimizers import Adam


class AUC:
    """tf.keras.metrics.AUC() with its defaults (ctr/train.py:86): 200 thresholds — -1e-7, i/199 for i = 1..198,
    1 + 1e-7 — a confusion matrix per threshold accumulated over batches, ROC curve, 'interpolation' summation
    (trapezoids over the threshold points)."""

    def __init__(self, num_thresholds: int = 200, device=None):
        eps = 1e-7
        inner = [(i + 1) / (num_thresholds - 1) for i in range(num_thresholds - 2)]
        self.thresholds = torch.tensor([0.0 - eps] + inner + [1.0 + eps], dtype=torch.float32, device=device)
        self.reset_states()

    def reset_states(self) -> None:
        n = self.thresholds.numel()
        self.tp = torch.zeros(n, dtype=torch.float64, device=self.thresholds.device)
        self.fp = torch.zeros_like(self.tp)
        self.pos = torch.zeros((), dtype=torch.float64, device=self.thresholds.device)
        self.neg = torch.zeros_like(self.pos)

    def update_state(self, label: torch.Tensor, prob: torch.Tensor) -> None:
        # predicted positive at threshold t  <=>  prob > t; bucket b = number of thresholds below prob
        b = torch.bucketize(prob.float().reshape(-1), self.thresholds, right=False)      # thresholds[b-1] < prob <= thresholds[b]
        y = label.reshape(-1) > 0
        n = self.thresholds.numel()
        hist_pos = torch.bincount(b[y], minlength=n + 1).double()
        hist_neg = torch.bincount(b[~y], minlength=n + 1).double()
        # threshold i counts every sample whose bucket is > i
        self.tp += hist_pos.flip(0).cumsum(0).flip(0)[1:]
        self.fp += hist_neg.flip(0).cumsum(0).flip(0)[1:]
        self.pos += y.sum()
        self.neg += (~y).sum()

    def result(self) -> float:
        eps = 1e-7                                             # Keras divides with div_no_nan; same effect for empty classes
        tpr = self.tp / torch.clamp(self.pos, min=eps)
        fpr = self.fp / torch.clamp(self.neg, min=eps)
        return float(((fpr[:-1] - fpr[1:]) * (tpr[:-1] + tpr[1:]) / 2.0).sum())


class BinaryAccuracy:
    """tf.keras.metrics.BinaryAccuracy() (threshold 0.5, ctr/train.py:86)."""

    def __init__(self, device=None):
        self.device = device
        self.reset_states()

    def reset_states(self) -> None:
        self.correct = torch.zeros((), dtype=torch.float64, device=self.device)
        self.count = 0

    def update_state(self, label: torch.Tensor, prob: torch.Tensor) -> None:
        self.correct += ((prob.reshape(-1) > 0.5) == (label.reshape(-1) > 0)).sum()
        self.count += label.numel()

    def result(self) -> float:
        return float(self.correct) / max(self.count, 1)


def build_model(model_type: str, embedding_size: int, vocab_size: int, device, seed: int):
    num_int_fea, num_cat_fea = 13, 26                                                   # ctr/train.py:62-63
    gen = torch.Generator(device=device).manual_seed(seed)
    if model_type == "DLRM":
        return DLRM([512, 256, 64, embedding_size], [512, 256, 1], embedding_size, vocab_size, num_cat_fea, num_int_fea,
                    device=device, generator=gen)                                       # :73-76
    if model_type == "DeepFM":
        return DeepFM(embedding_size, vocab_size, num_int_fea, num_cat_fea, [512, 256, 1], device=device, generator=gen)   # :81-83
    raise ValueError(f"model_type must be DLRM or DeepFM, got {model_type!r}")


@torch.no_grad()
def evaluate(model, batches, device) -> Dict[str, float]:
    auc, acc = AUC(device=device), BinaryAccuracy(device=device)
    loss_sum, n = torch.zeros((), dtype=torch.float64, device=device), 0
    for features, label in batches:
        prob = model(features)
        loss_sum += bce_clipped(prob, label).double() * label.numel()
        n += label.numel()
        auc.update_state(label, prob)
        acc.update_state(label, prob)
    return {"loss": float(loss_sum) / max(n, 1), "auc": auc.result(), "binary_accuracy": acc.result(), "samples": n}


def _eager_step(model, optimizer, batch) -> torch.Tensor:
    cat, dense, label = batch
    optimizer.prepare_step()
    loss = bce_clipped(model({"cat_features": cat, "int_features": dense}), label)
    loss.backward()
    optimizer.apply_gradients(model)
    return loss.detach()


def fit(model, optimizer, train_batches, val_batches, epochs: int, device, checkpoint_path: Optional[str] = None,
        patience: int = 3, max_steps: Optional[int] = None, use_graph: bool = True, log=print) -> List[Dict[str, float]]:
    """Keras `fit` (ctr/train.py:97) reduced to what the reference uses: per-epoch training loss, validation metrics,
    EarlyStopping(patience=3, monitor='val_auc', mode='max') and ModelCheckpoint(save_best_only, weights only).

    use_graph: the step is captured once as a CUDA graph and replayed (graph.GraphedTrainStep).  Capturing runs the
    very first batch through the optimizer twice (one warm-up pass, one replay); use_graph=False launches every step
    from Python and applies each batch exactly once."""
    history: List[Dict[str, float]] = []
    step_fn: Optional[GraphedTrainStep] = None
    best, since_best = -1.0, 0
    for epoch in range(epochs):
        loss_sum, seen, steps = torch.zeros((), dtype=torch.float64, device=device), 0, 0
        t0 = time.perf_counter()
        for features, label in train_batches():
            batch = (features["cat_features"], features["int_features"], label)
            if not use_graph:
                loss = _eager_step(model, optimizer, batch)
            elif step_fn is None:
                step_fn = GraphedTrainStep(model, optimizer, bce_clipped, batch, warmup=1)
                loss = step_fn.loss
            else:
                loss = step_fn.step(batch)
            loss_sum += loss.double() * label.numel()
            seen += label.numel()
            steps += 1
            if max_steps is not None and steps >= max_steps:
                break
        if steps == 0:
            raise ValueError("the training file holds no full batch")
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        row = {"epoch": epoch + 1, "loss": float(loss_sum) / seen, "steps": steps, "samples": seen, "seconds": dt,
               "samples_per_s": seen / max(dt, 1e-9)}
        val = evaluate(model, val_batches(), device)
        row.update({f"val_{k}": v for k, v in val.items()})
        history.append(row)
        log(f"Epoch {epoch + 1}/{epochs} - {dt:.2f}s - {row['samples_per_s']:.0f} samples/s - loss: {row['loss']:.4f} - "
            f"val_loss: {val['loss']:.4f} - val_auc: {val['auc']:.4f} - val_binary_accuracy: {val['binary_accuracy']:.4f}")
        if val["auc"] > best:
            best, since_best = val["auc"], 0
            if checkpoint_path is not None:
                os.makedirs(os.path.dirname(os.path.abspath(checkpoint_path)), exist_ok=True)
                torch.save(model.state_dict(), checkpoint_path)                          # save_weights_only=True (:89-93)
        else:
            since_best += 1
            if since_best >= patience:
                log(f"Epoch {epoch + 1}: early stopping")
                break
    return history


def train(argv=None) -> List[Dict[str, float]]:
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("--gpus", type=str, default="0")                                # ctr/train.py:13
    parser.add_argument("--gpu_memory_limit", type=int, default=20480)                  # :14 (its use is commented out there too)
    parser.add_argument("--model_type", type=str, default="DLRM")                       # :15
    parser.add_argument("--train_batch_size", type=int, default=1024)                   # :16
    parser.add_argument("--test_batch_size", type=int, default=4096)                    # :17
    parser.add_argument("--seed", type=int, default=4)                                  # :18
    # what the reference hard-codes (:56-57, :64-66, :68-69)
    parser.add_argument("--train_file", default="./data/train_split.tfrecord", help="record file (write_tfrecord) or raw Criteo text")
    parser.add_argument("--test_file", default="./data/test_split.tfrecord")
    parser.add_argument("--vocab", default=None, help=".npy written by build_vocab(save_to=...); built from --train_file when the "
                                                      "files are raw text and this is not given")
    parser.add_argument("--vocab_size", type=int, default=1_000_000)
    parser.add_argument("--embedding_size", type=int, default=16)
    parser.add_argument("--epochs", type=int, default=3)
    parser.add_argument("--ckpt_path", default="./ckpts")
    parser.add_argument("--max_steps", type=int, default=None, help="stop every epoch after this many steps (smoke runs)")
    parser.add_argument("--no_graph", action="store_true", help="launch every step from Python instead of replaying a CUDA graph")
    args = parser.parse_args(argv)
    first = args.gpus.split(",")[0]
    if "," in args.gpus:
        print(f"one process drives one GPU: using GPU {first} of --gpus {args.gpus} (multi-GPU: P2PShardedDLRM, bench.py --gpus N)")
    device = torch.device("cuda", int(first))
    torch.cuda.set_device(device)
    torch.manual_seed(args.seed)

    vocab = None
    raw_text = not tfrecord_io._is_record_file(args.train_file)
    if raw_text:
        vocab = tfrecord_io.Vocab.load(args.vocab, device) if args.vocab else tfrecord_io.build_vocab(args.train_file, device=device)
        if len(vocab) > args.vocab_size:
            raise ValueError(f"{len(vocab)} dictionary entries do not fit a table of --vocab_size {args.vocab_size} rows")
    model = build_model(args.model_type, args.embedding_size, args.vocab_size, device, args.seed)
    optimizer = Adam()                                                                  # :80 / :84

    def train_batches():          # static shapes for the captured step: the ragged last batch of the file is dropped
        return tfrecord_io.read_tfrecord(args.train_file, vocab, args.train_batch_size, device=device, drop_remainder=True)

    def val_batches():
        return tfrecord_io.read_tfrecord(args.test_file, vocab, args.test_batch_size, device=device)

    checkpoint = os.path.join(args.ckpt_path, args.model_type, "checkpoint.pt")         # :69-70
    return fit(model, optimizer, train_batches, val_batches, args.epochs, device, checkpoint, max_steps=args.max_steps,
               use_graph=not args.no_graph)


if __name__ == "__main__":
    train()
