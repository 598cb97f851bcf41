#!/usr/bin/env python
"""Input-side micro-benchmark: Criteo TSV text resident in HBM -> batch, through the C ABI.

    python scripts/input_bench.py [--lines 1000000] [--iters 10] [--cpu-lines 20000]

Times rb_criteo_index_lines, rb_criteo_parse (tokens only / with the dictionary lookup fused), rb_vocab_build and the
whole `read_tfrecord`-equivalent chain with CUDA events on the launching stream; the text (~250 MB per million lines)
is larger than L2.  Reports lines/s and GB/s of algorithmic bytes (text in + batch out) against the measured HBM peak.
The CPU leg (--cpu-lines > 0) times the restated reference loop (oracle/criteo_oracle.py = ctr/tfrecord_io.py:43-66,
pure Python like the reference) on a bounded sample, one core — a reported baseline, as in bench.py.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommender_b200 import tfrecord_io as io  # noqa: E402


def synth_block(n_lines, seed=4, n_hot=2000):
    """Criteo-format lines: ~8 % empty columns, a hot set per categorical column plus a long tail."""
    rng = np.random.default_rng(seed)
    hot = rng.integers(0, 2 ** 32, size=(26, n_hot))
    label = (rng.random(n_lines) < 0.25).astype(np.int64)
    ints = rng.integers(0, 5000, size=(n_lines, 13))
    int_empty = rng.random((n_lines, 13)) < 0.1
    cat_tail = rng.integers(0, 2 ** 32, size=(n_lines, 26))
    cat_hot = hot[np.arange(26)[None, :], rng.integers(0, n_hot, size=(n_lines, 26))]
    u = rng.random((n_lines, 26))
    cat = np.where(u < 0.8, cat_hot, cat_tail)
    cat_empty = u > 0.92
    rows = []
    for i in range(n_lines):
        cols = [str(label[i])]
        cols += ["" if e else str(v) for v, e in zip(ints[i], int_empty[i])]
        cols += ["" if e else f"{v:08x}" for v, e in zip(cat[i], cat_empty[i])]
        rows.append("\t".join(cols))
    return ("\n".join(rows) + "\n").encode("ascii")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=1_000_000)
    ap.add_argument("--block-lines", type=int, default=20_000)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--cpu-lines", type=int, default=20_000)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    block = synth_block(a.block_lines)
    reps = max(1, a.lines // a.block_lines)
    text_host = block * reps
    n_lines = a.block_lines * reps
    text = io.to_device(text_host, dev)
    nbytes = text.numel()
    peaks_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6552.0
    res = {"lines": n_lines, "text_bytes": nbytes, "bytes_per_line": nbytes / n_lines, "hbm_peak_gbs": peak}

    def timeit(name, fn, alg_bytes):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(a.iters):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        ts = sorted(s.elapsed_time(e) for s, e in evs)
        ms = ts[len(ts) // 2]
        res[name] = dict(ms=ms, lines_per_s=n_lines / ms * 1e3, algorithmic_bytes=alg_bytes, gbs=alg_bytes / ms / 1e6,
                         frac_of_hbm_peak=alg_bytes / ms / 1e6 / peak)

    # building blocks, with the buffers allocated once (the library never allocates)
    starts = io.index_lines(text)
    assert starts.numel() == n_lines
    ws = torch.empty(io.lib.rb_criteo_index_workspace_bytes(nbytes), dtype=torch.uint8, device=dev)
    room = torch.empty(n_lines, dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    st = lambda: torch.cuda.current_stream().cuda_stream     # noqa: E731
    timeit("index_lines", lambda: io.check(io.lib.rb_criteo_index_lines(text.data_ptr(), nbytes, n_lines, room.data_ptr(), cnt.data_ptr(),
                                                                         ws.data_ptr(), ws.numel(), st())),
           2 * nbytes + 8 * n_lines)                      # two passes over the text by design + the offsets
    label = torch.empty(n_lines, dtype=torch.int64, device=dev)
    ints = torch.empty(n_lines, 13, dtype=torch.float32, device=dev)
    tokens = torch.empty(n_lines, 26, dtype=torch.int64, device=dev)
    cats = torch.empty(n_lines, 26, dtype=torch.int64, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    out_bytes = n_lines * (8 + 13 * 4 + 26 * 8)
    timeit("parse_tokens", lambda: io.check(io.lib.rb_criteo_parse(text.data_ptr(), nbytes, starts.data_ptr(), n_lines, label.data_ptr(),
                                                                   ints.data_ptr(), tokens.data_ptr(), None, None, None, 0,
                                                                   flag.data_ptr(), st())), nbytes + 8 * n_lines + out_bytes)
    assert int(flag.item()) == 0
    vws = torch.empty(io.lib.rb_vocab_build_workspace_bytes(tokens.numel()), dtype=torch.uint8, device=dev)
    vout = torch.empty(tokens.numel() // 11 + 1, dtype=torch.int64, device=dev)
    vcnt = torch.zeros(1, dtype=torch.int64, device=dev)
    timeit("vocab_build", lambda: io.check(io.lib.rb_vocab_build(tokens.data_ptr(), tokens.numel(), 10, vout.data_ptr(), vout.numel(),
                                                                 vcnt.data_ptr(), vws.data_ptr(), vws.numel(), st())),
           tokens.numel() * 8)
    vocab = io.Vocab(vout[: int(vcnt.item())].clone())
    res["vocab_size"] = len(vocab)
    timeit("parse_with_lookup", lambda: io.check(io.lib.rb_criteo_parse(text.data_ptr(), nbytes, starts.data_ptr(), n_lines,
                                                                        label.data_ptr(), ints.data_ptr(), None, cats.data_ptr(),
                                                                        vocab.table_keys.data_ptr(), vocab.table_vals.data_ptr(),
                                                                        vocab.capacity, flag.data_ptr(), st())),
           nbytes + 8 * n_lines + out_bytes)
    assert int(flag.item()) == 0
    res["oov_or_id0_fraction"] = float((cats == 0).float().mean())
    # the user-facing chain: index + parse + lookup with allocation and the count read-back (one host sync) inside
    for _ in range(2):
        io.parse(text, vocab)                             # the caching allocator now holds the output buffers
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        io.parse(text, vocab)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / a.iters * 1e3
    res["parse_api_wall"] = dict(ms=ms, lines_per_s=n_lines / ms * 1e3, text_gbs=nbytes / ms / 1e6)
    # the file on disk (page cache) -> pinned staging -> HBM -> batches of 65536: what `read_tfrecord` delivers end to end
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "day.txt")
        with open(path, "wb") as fh:
            fh.write(text_host)
        for _ in io.read_tfrecord(path, vocab, 65536):    # warm-up: page cache, pinned buffer, allocator
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        seen = 0
        for feats, lab in io.read_tfrecord(path, vocab, 65536):
            seen += lab.numel()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        assert seen == n_lines
        res["read_tfrecord_file_wall"] = dict(ms=ms, lines_per_s=n_lines / ms * 1e3, h2d_bytes=nbytes, batch=65536)
        # the same content as a preprocessed record file (write_tfrecord once, read every epoch)
        rec = os.path.join(tmp, "day.tfrecord")
        t0 = time.perf_counter()
        assert io.write_tfrecord(path, rec, vocab) == n_lines
        res["write_tfrecord_wall"] = dict(ms=(time.perf_counter() - t0) * 1e3, record_bytes=io.RECORD_BYTES)
        for _ in io.read_tfrecord(rec, batch_size=65536):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        seen = 0
        for feats, lab in io.read_tfrecord(rec, batch_size=65536):
            seen += lab.numel()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        assert seen == n_lines
        res["read_tfrecord_records_wall"] = dict(ms=ms, lines_per_s=n_lines / ms * 1e3, h2d_bytes=n_lines * io.RECORD_BYTES,
                                                 host_gbs=n_lines * io.RECORD_BYTES / ms / 1e6, batch=65536)
        t0 = time.perf_counter()
        v2 = io.build_vocab(path)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        assert len(v2) == len(vocab)
        res["build_vocab_file_wall"] = dict(ms=ms, lines_per_s=n_lines / ms * 1e3)
    if a.cpu_lines > 0:
        from oracle import criteo_oracle as CO           # the CPU leg only: restated ctr/tfrecord_io.py:43-66
        sample = CO.split_lines(block)[: a.cpu_lines]
        imp = [f"MISSING_{f:02d}" for f in range(26)]
        t0 = time.perf_counter()
        ref_vocab = CO.build_vocab(sample, imp)
        t1 = time.perf_counter()
        CO.transform(sample, ref_vocab, imp)
        t2 = time.perf_counter()
        res["cpu_baseline"] = dict(kind="port", cores=1, sample=f"{len(sample)} lines, pure-Python per-line loop as in the reference",
                                   build_vocab_lines_per_s=len(sample) / (t1 - t0), transform_lines_per_s=len(sample) / (t2 - t1))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
