"""Input side without a GPU.

1. oracle/criteo_oracle.py against tests/golden/criteo_tsv.npz — the outputs of the reference's OWN `build_vocab` and
   `write_tfrecord` (ctr/tfrecord_io.py imported byte-for-byte under the tf shim, tests/golden/make_golden_criteo.py).
   This pins the oracle.
2. recommender_b200/csrc/criteo_fields.h — the field-level code the parse kernel runs — compiled for the HOST with g++
   (tests/criteo_fields_host.cpp) against the oracle: integers bit-exact, token keys bit-exact, log within 4 ULP.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import criteo_oracle as CO

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def g(golden):
    d = golden("criteo_tsv")
    d["imp"] = [str(s) for s in d["cat_imputation"]]
    d["train_lines"] = CO.split_lines(d["train_tsv"].tobytes())
    d["test_lines"] = CO.split_lines(d["test_tsv"].tobytes())
    return d


def test_oracle_vocab_matches_reference(g):
    vocab = CO.build_vocab(g["train_lines"], g["imp"])
    assert list(vocab.keys()) == [str(t) for t in g["vocab_tokens"]]           # same entries, same ids (first-seen order)
    assert list(vocab.values()) == list(range(len(vocab)))
    assert sum(t.endswith("\n") for t in vocab) > 0                            # C26's tokens keep their newline
    assert sum(t in g["imp"] for t in vocab) == 26                             # every column's imputation token is frequent


@pytest.mark.parametrize("split", ["train", "test"])
def test_oracle_records_match_reference(g, split):
    vocab = CO.build_vocab(g["train_lines"], g["imp"])
    ints, cats, label = CO.transform(g[f"{split}_lines"], vocab, g["imp"])
    np.testing.assert_array_equal(ints, g[f"{split}_int_features"])            # same numpy float32 log: bit-exact
    np.testing.assert_array_equal(cats, g[f"{split}_cat_features"])
    np.testing.assert_array_equal(label, g[f"{split}_label"])
    assert not g["test_tsv"].tobytes().endswith(b"\n")                         # the fixture covers a last line without newline


def test_short_line_raises_like_the_reference(g):
    with pytest.raises(IndexError):
        CO.transform_line("1\t2\t3\n", {}, g["imp"])


def test_packed_keys_are_an_identity(g):
    """Distinct dictionary strings <-> distinct 64-bit keys on the fixture (tokens of at most 8 ASCII bytes)."""
    toks = set()
    for ln in g["train_lines"] + g["test_lines"]:
        toks.update(CO.cat_tokens_of_line(ln, g["imp"]))
    keys = {CO.key_of(t, g["imp"]) for t in toks}
    assert len(keys) == len(toks)
    body = "0a1b2c3d"
    assert CO.pack_token(body) != CO.pack_token(body + "\n")
    assert CO.pack_token(body + "\n") == CO.pack_token(body) | CO.NEWLINE_BIT
    assert len({CO.missing_key(f) for f in range(26)}) == 26 and all(CO.missing_key(f) & 0xFF == 0 for f in range(26))
    with pytest.raises(ValueError):
        CO.pack_token("123456789")


# ---- the kernel's field-level code, built for the host ----------------------------------------------------------------

@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = tmp_path_factory.mktemp("cf") / "libcriteo_fields_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out), os.path.join(HERE, "criteo_fields_host.cpp")],
                   check=True)
    dll = C.CDLL(str(out))
    dll.t_parse_line.restype = C.c_int
    dll.t_parse_line.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    dll.t_mix64.restype = C.c_uint64
    dll.t_mix64.argtypes = [C.c_uint64]
    dll.t_vocab_find.restype = C.c_int64
    dll.t_vocab_find.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]
    dll.t_table_build.restype = None
    dll.t_table_build.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_uint64]
    return dll


def _host_parse(dll, line: str):
    has_nl = line.endswith("\n")
    raw = (line[:-1] if has_nl else line).encode("ascii")
    label = C.c_int64(0)
    ints = (C.c_float * 13)()
    keys = (C.c_uint64 * 26)()
    err = dll.t_parse_line(raw, len(raw), int(has_nl), C.byref(label), ints, keys)
    return err, label.value, np.array(ints[:], dtype=np.float32), np.array(keys[:], dtype=np.uint64)


def ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


@pytest.mark.parametrize("split", ["train", "test"])
def test_host_fields_match_oracle(host, g, split):
    vocab = CO.build_vocab(g["train_lines"], g["imp"])
    lines = g[f"{split}_lines"]
    ref_int, ref_cat, ref_label = CO.transform(lines, vocab, g["imp"])
    ref_keys = CO.token_keys(lines, g["imp"])
    vkeys = CO.vocab_keys(vocab, g["imp"])
    cap = 2048
    tkeys, tvals = np.empty(cap, np.uint64), np.empty(cap, np.int32)
    host.t_table_build(vkeys.ctypes.data, len(vkeys), tkeys.ctypes.data, tvals.ctypes.data, cap - 1)
    worst = 0
    for k, ln in enumerate(lines):
        err, label, ints, keys = _host_parse(host, ln)
        assert err == 0
        assert label == ref_label[k]
        np.testing.assert_array_equal(keys, ref_keys[k])
        worst = max(worst, int(ulp_diff(ints, ref_int[k]).max()))
        ids = [host.t_vocab_find(tkeys.ctypes.data, tvals.ctypes.data, cap - 1, int(x)) for x in keys]
        assert ids == ref_cat[k].tolist()
    assert worst <= 4, f"log(x + 1) differs from numpy's float32 log by {worst} ULP"


def test_host_mix64_matches_oracle(host):
    rng = np.random.default_rng(0)
    ks = rng.integers(0, 2 ** 63, size=64, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    ref = CO.mix64(ks)
    assert [host.t_mix64(int(k)) for k in ks] == [int(r) for r in ref]


def test_host_error_bits_and_quirks(host, g):
    cols = ["1"] + ["5"] * 13 + ["0a1b2c3d"] * 26
    ok = "\t".join(cols) + "\n"
    err, label, ints, keys = _host_parse(host, ok)
    assert err == 0 and label == 1
    np.testing.assert_array_equal(ints, np.log(np.full(13, 5, np.float32) + 1))
    assert keys[0] == CO.pack_token("0a1b2c3d") and keys[25] == CO.pack_token("0a1b2c3d\n")
    # a 41st column: C26 is no longer the last one and loses the newline (str.split semantics)
    err, _, _, keys = _host_parse(host, "\t".join(cols + ["extra"]) + "\n")
    assert err == 0 and keys[25] == CO.pack_token("0a1b2c3d")
    # no trailing newline on the file's last line
    err, _, _, keys = _host_parse(host, "\t".join(cols))
    assert err == 0 and keys[25] == CO.pack_token("0a1b2c3d")
    # empty columns: integer '' -> 0 -> log(1) = 0; categorical '' and the bare '\n' -> the column's imputation key
    e = list(cols)
    e[3], e[20], e[39] = "", "", ""
    err, _, ints, keys = _host_parse(host, "\t".join(e) + "\n")
    assert err == 0 and ints[2] == 0.0 and keys[6] == CO.missing_key(6) and keys[25] == CO.missing_key(25)
    # negative -> 0 (:48-49); beyond float32's integers: int64 -> float32 rounds to nearest even like numpy
    e = list(cols)
    e[1], e[2] = "-7", "123456789012"
    err, _, ints, _ = _host_parse(host, "\t".join(e) + "\n")
    assert err == 0 and ints[0] == 0.0
    assert ulp_diff(ints[1], np.log(np.array([123456789012]).astype(np.float32) + 1)).max() <= 4
    # errors
    assert _host_parse(host, "1\t2\t3\n")[0] == 1                                   # short line (IndexError in the reference)
    e = list(cols)
    e[5] = "12x"
    assert _host_parse(host, "\t".join(e) + "\n")[0] == 2                           # int() raises ValueError
    e = list(cols)
    e[30] = "123456789"
    assert _host_parse(host, "\t".join(e) + "\n")[0] == 4                           # no 64-bit key
    assert _host_parse(host, "\t" + "\t".join(cols[1:]) + "\n")[0] == 2             # int('') for the label


# ---- host half of the chunked file reader ------------------------------------------------------------------------------

@pytest.mark.parametrize("final_newline", [True, False])
@pytest.mark.parametrize("chunk", [64, 100, 257, 4096, 1 << 20])
def test_fill_chunks_yields_whole_lines(tmp_path, chunk, final_newline):
    """recommender_b200.tfrecord_io._fill_chunks: every chunk ends on a line boundary, nothing is lost or repeated,
    whatever the chunk size (here down to barely more than one line)."""
    from recommender_b200.tfrecord_io import CriteoFormatError, _fill_chunks
    rng = np.random.default_rng(chunk)
    lines = [b"x" * int(rng.integers(0, 60)) for _ in range(200)]
    text = b"\n".join(lines) + (b"\n" if final_newline else b"")
    path = tmp_path / "f.txt"
    path.write_bytes(text)
    view = np.empty(min(chunk, len(text)), dtype=np.uint8)
    got = []
    with open(path, "rb") as fh:
        for cut in _fill_chunks(fh, view):
            piece = view[:cut].tobytes()
            got.append(piece)
    assert b"".join(got) == text
    assert all(p.endswith(b"\n") for p in got[:-1])
    assert all(len(p) <= chunk for p in got)
    if chunk >= len(text):
        assert len(got) == 1
    # a line that does not fit the buffer is an error, not a silent split
    path.write_bytes(b"y" * 300 + b"\n")
    with open(path, "rb") as fh, pytest.raises(CriteoFormatError):
        list(_fill_chunks(fh, np.empty(128, dtype=np.uint8)))
    path.write_bytes(b"")
    with open(path, "rb") as fh:
        assert list(_fill_chunks(fh, np.empty(16, dtype=np.uint8))) == []


def test_host_parse_int_matches_python_int(host):
    """parse_int (SWAR path for up to 8 digits, byte loop beyond) against Python's int() on what Criteo files hold."""
    host.t_parse_int.restype = C.c_int
    host.t_parse_int.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
    rng = np.random.default_rng(3)
    cases = ["0", "7", "-1", "+5", "00012", "99999999", "-9999999", "123456789", "-123456789012345678", "999999999999999999"]
    for digits in range(1, 19):
        for _ in range(40):
            v = int(rng.integers(0, 10 ** digits))
            cases += [str(v), str(-v), str(v).zfill(digits)]
    for text in cases:
        out = C.c_int64(0)
        assert host.t_parse_int(text.encode(), len(text), C.byref(out)) == 1, text
        assert out.value == int(text), text
    for text in ["", "-", "+", "12x", "x12", "1 2", "1.5", "--1", "1-", "12345678x", "1" * 19, "\t1", "1e3", "0x10", ":", "/"]:
        out = C.c_int64(0)
        assert host.t_parse_int(text.encode(), len(text), C.byref(out)) == 0, text


@pytest.mark.parametrize("world", [1, 2, 8])
@pytest.mark.parametrize("n,B", [(0, 4), (3, 4), (64, 8), (70, 8), (1000, 64)])
def test_batch_ranges_shard_the_file_across_ranks(n, B, world):
    """recommender_b200.tfrecord_io._batch_ranges: the ranks' batches are disjoint, cover the file in order, and with
    drop_remainder every rank sees the same number of full batches (matched collectives in a data-parallel step)."""
    from recommender_b200.tfrecord_io import _batch_ranges
    per_rank = [list(_batch_ranges(n, B, r, world, False)) for r in range(world)]
    merged = sorted(x for pr in per_rank for x in pr)
    assert merged == [(s, min(s + B, n)) for s in range(0, n, B)]
    for r, pr in enumerate(per_rank):
        assert all((s // B) % world == r for s, _ in pr)
    full = [list(_batch_ranges(n, B, r, world, True)) for r in range(world)]
    assert len({len(f) for f in full}) == 1
    assert all(e - s == B for f in full for s, e in f)
    assert sum(len(f) for f in full) == (n // B) - (n // B) % world
    with pytest.raises(ValueError):
        list(_batch_ranges(n, B, world, world, False))


def test_record_reader_thread_on_cpu(tmp_path, monkeypatch):
    """recommender_b200.tfrecord_io._read_records with its two CUDA touch points (pinned allocation, copy-done event)
    replaced: the reader thread, the staging-buffer hand-back and early close of the generator, without a GPU."""
    import threading
    import torch
    from recommender_b200 import tfrecord_io as io

    class Done:
        def wait(self):
            pass
    monkeypatch.setattr(io, "_pinned", lambda *shape: torch.empty(*shape, dtype=torch.uint8))
    monkeypatch.setattr(io, "_CopyDone", Done)
    n = 1003
    g = torch.Generator().manual_seed(0)
    label = torch.randint(0, 2, (n,), generator=g)
    ints = torch.randn(n, 13, generator=g)
    cats = torch.randint(0, 1 << 31, (n, 26), generator=g)
    raw = io._pack_records(label, ints, cats)
    assert raw.shape == (n, io.RECORD_BYTES) and io.RECORD_BYTES == 160
    with pytest.raises(io.CriteoFormatError):
        io._pack_records(label + 2 ** 31, ints, cats)                       # does not fit the 32-bit field
    path = tmp_path / "x.tfrecord"
    path.write_bytes(io._record_header() + raw.numpy().tobytes())
    assert io._is_record_file(str(path))
    cpu = torch.device("cpu")
    for B in (1, 7, 64, 1003, 5000):
        batches = list(io._read_records(str(path), B, cpu, False))
        assert [len(b[1]) for b in batches] == [B] * (n // B) + ([n % B] if n % B else [])
        assert torch.equal(torch.cat([b[1] for b in batches]), label)
        assert torch.equal(torch.cat([b[0]["cat_features"] for b in batches]), cats)
        assert torch.equal(torch.cat([b[0]["int_features"] for b in batches]), ints)
    parts = [list(io._read_records(str(path), 16, cpu, False, r, 4)) for r in range(4)]
    order = [parts[k % 4][k // 4] for k in range(sum(len(p) for p in parts))]
    assert torch.equal(torch.cat([b[1] for b in order]), label)
    # closing the generator early stops the reader thread
    before = threading.active_count()
    it = io._read_records(str(path), 8, cpu, False)
    next(it)
    it.close()
    assert threading.active_count() <= before
    (tmp_path / "bad").write_bytes(io._record_header() + b"123")
    with pytest.raises(io.CriteoFormatError):
        list(io._read_records(str(tmp_path / "bad"), 4, cpu, False))
    (tmp_path / "old").write_bytes(io.RECORD_MAGIC + np.array([1, 13, 26, 268], dtype="<i4").tobytes() + b"\0" * 40)
    with pytest.raises(io.CriteoFormatError):
        io._is_record_file(str(tmp_path / "old"))                            # another layout is refused, not misread
    (tmp_path / "text").write_bytes(b"1\t2\t3\n")
    assert not io._is_record_file(str(tmp_path / "text"))
