#!/usr/bin/env python
"""End to end from a Criteo file: `recommender_b200.train` (the reference's ctr/train.py entry point) on a synthetic
1M-line file — dictionary, record file, then one epoch per configuration with the input pipeline inside the timed
region (disk cache -> pinned -> HBM -> CUDA-graph step).  Prints one JSON line.

    python scripts/train_bench.py [--lines 1000000]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from input_bench import synth_block  # noqa: E402
from recommender_b200 import tfrecord_io as io  # noqa: E402
from recommender_b200.train import train  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=1_000_000)
    ap.add_argument("--only", default="", help="comma-separated configuration tags to run (default: all)")
    a = ap.parse_args()
    block = synth_block(20_000)
    res = {"lines": a.lines}
    with tempfile.TemporaryDirectory() as tmp:
        raw, test_raw = os.path.join(tmp, "train.txt"), os.path.join(tmp, "test.txt")
        with open(raw, "wb") as fh:
            fh.write(block * (a.lines // 20_000))
        with open(test_raw, "wb") as fh:
            fh.write(synth_block(20_000, seed=5))
        t0 = time.perf_counter()
        vocab = io.build_vocab(raw, save_to=os.path.join(tmp, "vocab.npy"))
        res["build_vocab_s"] = time.perf_counter() - t0
        res["vocab_size"] = len(vocab)
        t0 = time.perf_counter()
        io.write_tfrecord(raw, os.path.join(tmp, "train.tfrecord"), vocab)
        io.write_tfrecord(test_raw, os.path.join(tmp, "test.tfrecord"), vocab)
        res["write_tfrecord_s"] = time.perf_counter() - t0
        del vocab
        torch.cuda.empty_cache()
        configs = [("dlrm_b1024_reference_defaults", ["--train_batch_size", "1024"]),
                   ("dlrm_b65536", ["--train_batch_size", "65536"]),
                   ("dlrm_b65536_emb64", ["--train_batch_size", "65536", "--embedding_size", "64"]),
                   ("deepfm_b1024_reference_defaults", ["--model_type", "DeepFM", "--train_batch_size", "1024"]),
                   ("dlrm_b65536_raw_text", ["--train_batch_size", "65536", "--train_file", raw, "--test_file", test_raw,
                                             "--vocab", os.path.join(tmp, "vocab.npy")])]
        for tag, extra in configs:
            if a.only and tag not in a.only.split(","):
                continue
            args = ["--train_file", os.path.join(tmp, "train.tfrecord"), "--test_file", os.path.join(tmp, "test.tfrecord"),
                    "--epochs", "2", "--ckpt_path", os.path.join(tmp, "ckpts")] + extra
            hist = train(args)
            last = hist[-1]                                  # epoch 2: graph captured, allocator and page cache warm
            res[tag] = {k: last[k] for k in ("samples", "steps", "seconds", "samples_per_s", "loss", "val_loss", "val_auc")}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
