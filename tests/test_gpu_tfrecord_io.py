"""Input side on the GPU: recommender_b200/tfrecord_io.py -> C ABI -> csrc/criteo_input.cu against

* tests/golden/criteo_tsv.npz — records and dictionary produced by the reference's OWN ctr/tfrecord_io.py, and
* oracle/criteo_oracle.py on larger seeded files.

Bars: line offsets, labels, token keys, dictionary (entries AND ids) and cat_features bit-exact; int_features within
4 ULP of numpy's float32 log (the device's logf and numpy's SIMD log are both within a few ULP of the true value).
"""
import numpy as np
import pytest
import torch

from oracle import criteo_oracle as CO

pytestmark = pytest.mark.gpu

ULP = 4


@pytest.fixture(scope="module")
def io(cuda_lib):
    from recommender_b200 import tfrecord_io
    return tfrecord_io


@pytest.fixture(scope="module")
def g(golden):
    d = golden("criteo_tsv")
    d["imp"] = [str(s) for s in d["cat_imputation"]]
    d["train_lines"] = CO.split_lines(d["train_tsv"].tobytes())
    d["test_lines"] = CO.split_lines(d["test_tsv"].tobytes())
    return d


def ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def u64(t):
    return t.cpu().numpy().view(np.uint64)


def line_starts(text: bytes):
    out, pos = [], 0
    for ln in CO.split_lines(text):
        out.append(pos)
        pos += len(ln)
    return np.array(out, dtype=np.int64)


@pytest.mark.parametrize("text", [b"", b"\n", b"a", b"a\n", b"a\nb", b"a\n\nb\n", b"\n\n\n", b"x" * 5000 + b"\n" + b"y" * 17])
def test_index_lines_edge_cases(io, text):
    starts = io.index_lines(io.to_device(text)).cpu().numpy()
    np.testing.assert_array_equal(starts, line_starts(text))


@pytest.mark.parametrize("split", ["train", "test"])
def test_index_lines_golden(io, g, split):
    text = g[f"{split}_tsv"].tobytes()
    starts = io.index_lines(io.to_device(text)).cpu().numpy()
    np.testing.assert_array_equal(starts, line_starts(text))
    assert len(starts) == len(g[f"{split}_label"])


def test_vocab_matches_reference_dictionary(io, g):
    """build_vocab (ctr/tfrecord_io.py:15-35): same entries, same ids as the reference's own run."""
    feats, _ = io.parse(io.to_device(g["train_tsv"].tobytes()), None)
    np.testing.assert_array_equal(u64(feats["cat_tokens"]), CO.token_keys(g["train_lines"], g["imp"]))
    vocab = io.vocab_from_tokens(feats["cat_tokens"])
    ref = {str(t): k for k, t in enumerate(g["vocab_tokens"])}
    np.testing.assert_array_equal(u64(vocab.keys), CO.vocab_keys(ref, g["imp"]))


@pytest.mark.parametrize("split", ["train", "test"])
def test_records_match_reference_writer(io, g, split):
    """write_tfrecord's per-line body (:43-66) as frozen from the reference, dictionary built on the GPU."""
    train = io.to_device(g["train_tsv"].tobytes())
    vocab = io.vocab_from_tokens(io.parse(train, None)[0]["cat_tokens"])
    feats, label = io.parse(io.to_device(g[f"{split}_tsv"].tobytes()), vocab)
    np.testing.assert_array_equal(label.cpu().numpy(), g[f"{split}_label"])
    np.testing.assert_array_equal(feats["cat_features"].cpu().numpy(), g[f"{split}_cat_features"])
    assert feats["cat_features"].dtype == torch.int64 and feats["int_features"].dtype == torch.float32     # :81-82
    assert ulp_diff(feats["int_features"].cpu().numpy(), g[f"{split}_int_features"]).max() <= ULP


def test_files_chunks_and_batches(io, g, tmp_path):
    """build_vocab(file) and read_tfrecord(file, vocab, B) with chunks far smaller than the file: chunk seams and the
    batch carry-over must not show."""
    train, test = tmp_path / "train.txt", tmp_path / "test.txt"
    train.write_bytes(g["train_tsv"].tobytes())
    test.write_bytes(g["test_tsv"].tobytes())
    vocab = io.build_vocab(str(train), chunk_bytes=10_000, save_to=str(tmp_path / "data" / "cat_fea_vocab.npy"))
    ref = {str(t): k for k, t in enumerate(g["vocab_tokens"])}
    np.testing.assert_array_equal(u64(vocab.keys), CO.vocab_keys(ref, g["imp"]))
    again = io.Vocab.load(str(tmp_path / "data" / "cat_fea_vocab.npy"))
    np.testing.assert_array_equal(u64(again.keys), u64(vocab.keys))
    for path, split, B in ((train, "train", 64), (test, "test", 10)):
        batches = list(io.read_tfrecord(str(path), again, B, chunk_bytes=7_001))
        n = len(g[f"{split}_label"])
        assert [len(b[1]) for b in batches] == [B] * (n // B) + ([n % B] if n % B else [])
        cats = torch.cat([b[0]["cat_features"] for b in batches]).cpu().numpy()
        label = torch.cat([b[1] for b in batches]).cpu().numpy()
        ints = torch.cat([b[0]["int_features"] for b in batches]).cpu().numpy()
        np.testing.assert_array_equal(cats, g[f"{split}_cat_features"])
        np.testing.assert_array_equal(label, g[f"{split}_label"])
        assert ulp_diff(ints, g[f"{split}_int_features"]).max() <= ULP
    assert len(list(io.read_tfrecord(str(test), again, 10, drop_remainder=True))) == len(g["test_label"]) // 10


@pytest.mark.parametrize("min_count", [10, 0, 3])
def test_larger_file_against_oracle(io, min_count):
    """20k lines (5 MB, > 1000 index tiles), dictionary threshold varied; everything integer bit-exact."""
    imp = [f"MISSING_{f:02d}" for f in range(26)]
    text = CO.synth_tsv(20_000, seed=11, n_hot=300)
    lines = CO.split_lines(text)
    count = {}
    for ln in lines:
        for t in CO.cat_tokens_of_line(ln, imp):
            count[t] = count.get(t, 0) + 1
    vocab_ref = {t: k for k, t in enumerate(t for t, c in count.items() if c > min_count)}
    dev = io.to_device(text)
    feats, label = io.parse(dev, None)
    np.testing.assert_array_equal(u64(feats["cat_tokens"]), CO.token_keys(lines, imp))
    vocab = io.vocab_from_tokens(feats["cat_tokens"], min_count)
    np.testing.assert_array_equal(u64(vocab.keys), CO.vocab_keys(vocab_ref, imp))
    feats, label = io.parse(dev, vocab)
    ints, cats, lab = CO.transform(lines, vocab_ref, imp)
    np.testing.assert_array_equal(feats["cat_features"].cpu().numpy(), cats)
    np.testing.assert_array_equal(label.cpu().numpy(), lab)
    assert ulp_diff(feats["int_features"].cpu().numpy(), ints).max() <= ULP
    tokens = io.parse(dev, None)[0]["cat_tokens"]
    assert tokens.shape == (20_000, 26)
    np.testing.assert_array_equal(vocab.lookup(tokens).cpu().numpy(), cats)          # the stand-alone lookup entry point


def test_parsed_batch_feeds_the_model(io, g):
    """The batch goes straight into the DLRM of ctr/model.py's call surface (ids index one shared table)."""
    from recommender_b200.model import DLRM
    train = io.to_device(g["train_tsv"].tobytes())
    vocab = io.vocab_from_tokens(io.parse(train, None)[0]["cat_tokens"])
    feats, label = io.parse(train, vocab)
    model = DLRM([32, 16], [32, 1], 16, len(vocab), 26, 13)
    prob = model({"cat_features": feats["cat_features"], "int_features": feats["int_features"]})
    assert prob.shape == (len(label),) and torch.isfinite(prob).all()
    from recommender_b200 import ops
    ops.check_oob("cuda")


def test_error_bits(io):
    cols = ["1"] + ["5"] * 13 + ["0a1b2c3d"] * 26
    good = "\t".join(cols) + "\n"

    def flags(text):
        feats, label = io.parse(io.to_device(text.encode("ascii")), None, raise_on_error=False)
        return int(feats["error_flag"].item()), feats, label

    assert flags(good * 3)[0] == 0
    bits, feats, label = flags(good + "1\t2\t3\n" + good)                        # short line: IndexError in the reference
    assert bits == 1
    np.testing.assert_array_equal(label.cpu().numpy(), [1, 0, 1])                # the good lines are unaffected
    assert (u64(feats["cat_tokens"])[[0, 2], 0] == CO.pack_token("0a1b2c3d")).all()
    bad_int = list(cols)
    bad_int[5] = "12x"
    assert flags(good + "\t".join(bad_int) + "\n")[0] == 2
    long_tok = list(cols)
    long_tok[30] = "123456789"
    assert flags("\t".join(long_tok) + "\n" + good)[0] == 4
    long_line = list(cols)
    long_line[1] = "7" * 1100
    bits, feats, label = flags(good + "\t".join(long_line) + "\n" + good)
    assert bits == 16
    np.testing.assert_array_equal(label.cpu().numpy(), [1, 0, 1])
    with pytest.raises(io.CriteoFormatError):
        io.parse(io.to_device(b"1\t2\n"), None)
    # quirks: a 41st column takes the newline away from C26; the last line of a file may lack its newline
    feats = flags("\t".join(cols + ["extra"]) + "\n" + "\t".join(cols))[1]
    assert (u64(feats["cat_tokens"])[:, 25] == CO.pack_token("0a1b2c3d")).all()
    feats = flags(good)[1]
    assert u64(feats["cat_tokens"])[0, 25] == CO.pack_token("0a1b2c3d\n")


def test_arguments_are_validated(io):
    from recommender_b200._lib import RecsysError
    text = io.to_device(b"1\t2\n")
    with pytest.raises(RecsysError):
        io.parse(text[1:], None)                                                 # not 16-byte aligned
    with pytest.raises(RecsysError):
        io.to_device(b"abc", device="cpu")


def test_record_file_round_trip(io, g, tmp_path):
    """write_tfrecord (ctr/tfrecord_io.py:38-75) -> read_tfrecord (:78-96): the preprocessed record file delivers the
    reference's records bit for bit (the int_features are stored, not recomputed), in any batch size."""
    train, test = tmp_path / "train.txt", tmp_path / "test.txt"
    train.write_bytes(g["train_tsv"].tobytes())
    test.write_bytes(g["test_tsv"].tobytes())
    vocab = io.build_vocab(str(train))
    for path, split in ((train, "train"), (test, "test")):
        out = tmp_path / f"{split}.tfrecord"
        n = io.write_tfrecord(str(path), str(out), vocab, chunk_bytes=20_000)
        assert n == len(g[f"{split}_label"]) and out.stat().st_size == 64 + io.RECORD_BYTES * n
        direct = io.parse(io.to_device(path.read_bytes()), vocab)
        for B in (1, 7, 64, 10_000):
            batches = list(io.read_tfrecord(str(out), batch_size=B))
            assert [len(b[1]) for b in batches] == [B] * (n // B) + ([n % B] if n % B else [])
            cats = torch.cat([b[0]["cat_features"] for b in batches])
            ints = torch.cat([b[0]["int_features"] for b in batches])
            label = torch.cat([b[1] for b in batches])
            np.testing.assert_array_equal(cats.cpu().numpy(), g[f"{split}_cat_features"])
            np.testing.assert_array_equal(label.cpu().numpy(), g[f"{split}_label"])
            assert torch.equal(ints, direct[0]["int_features"])                      # stored bits = parsed bits
            assert cats.dtype == torch.int64 and ints.dtype == torch.float32 and label.dtype == torch.int64
        assert len(list(io.read_tfrecord(str(out), batch_size=10, drop_remainder=True))) == n // 10
        # one process per GPU: rank r of 2 reads batches r, r + 2, ... — disjoint, together the whole file
        halves = [list(io.read_tfrecord(str(out), batch_size=16, rank=r, world=2)) for r in (0, 1)]
        order = [halves[k % 2][k // 2] for k in range(len(halves[0]) + len(halves[1]))]
        np.testing.assert_array_equal(torch.cat([b[1] for b in order]).cpu().numpy(), g[f"{split}_label"])
        np.testing.assert_array_equal(torch.cat([b[0]["cat_features"] for b in order]).cpu().numpy(), g[f"{split}_cat_features"])
        even = [list(io.read_tfrecord(str(out), batch_size=16, rank=r, world=2, drop_remainder=True)) for r in (0, 1)]
        assert len(even[0]) == len(even[1]) == (n // 16) // 2
    (tmp_path / "bad.tfrecord").write_bytes((tmp_path / "test.tfrecord").read_bytes()[:-5])
    with pytest.raises(io.CriteoFormatError):
        list(io.read_tfrecord(str(tmp_path / "bad.tfrecord"), batch_size=4))
    with pytest.raises(ValueError):
        list(io.read_tfrecord(str(test), batch_size=4))                              # raw text needs the vocabulary
