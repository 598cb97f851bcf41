"""CPU restatement of the reference's Criteo input side — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/ctr/tfrecord_io.py line by line:

* ``build_vocab``   :15-35   one string -> id dictionary across all 26 categorical fields, kept if count > 10,
                              ids dense from 0 in first-seen order;
* ``transform_line`` :43-66   the per-line body of ``write_tfrecord`` (what ends up inside the TFRecord and comes back
                              out of ``read_tfrecord`` :78-96): 13 dense values log(max(x, 0) + 1) in float32, 26 ids with
                              OOV -> 0, the label.

PINNED: tests/golden/criteo_tsv.npz holds the outputs of the reference's OWN ``build_vocab`` / ``write_tfrecord``
(imported byte-for-byte under oracle/tf_shim, tests/golden/make_golden_criteo.py) on a synthetic TSV;
tests/test_oracle_golden.py checks this restatement against it.

The second half (``pack_token`` ...) is the build-defined 64-bit token key of the CUDA path.  The reference keys its
dictionary by Python strings; the build keys by the token's bytes packed into a uint64, which is the same identity for
every token of at most 8 ASCII bytes (Criteo's categorical values are 8 hex digits).
"""
import numpy as np

NUM_INT = 13        # ctr/tfrecord_io.py:8
NUM_CAT = 26        # :9
TOTAL_COLS = 40     # :10
MIN_COUNT = 10      # :31  kept if count > 10


def split_lines(text: bytes):
    """`for line in f` of a text file: every line keeps its trailing newline; a last line without one is kept too."""
    lines = text.split(b"\n")
    out = [ln + b"\n" for ln in lines[:-1]]
    if lines[-1] != b"":
        out.append(lines[-1])
    return [ln.decode("ascii") for ln in out]


def cat_tokens_of_line(line: str, cat_imputation):
    """:21-23 / :54-57 — the 26 categorical tokens of a line after null imputation.  str.split leaves the line's
    trailing newline attached to the last column, so C26's tokens are 'xxxxxxxx\\n' (a different dictionary key from
    the same value in another column); an empty last column is the one-character string '\\n'."""
    cols = line.split("\t")
    toks = []
    for i in range(NUM_INT + 1, TOTAL_COLS):
        tok = cols[i]                                     # IndexError for a short line, as in the reference
        if tok == "" or tok == "\n":
            tok = cat_imputation[i - NUM_INT - 1]
        toks.append(tok)
    return toks


def build_vocab(lines, cat_imputation):
    """:15-35.  Returns the dict token -> id (insertion order = id order)."""
    count = {}
    for line in lines:
        for tok in cat_tokens_of_line(line, cat_imputation):
            count[tok] = count.get(tok, 0) + 1
    vocab = {}
    idx = 0
    for key, c in count.items():                          # dict order = first-seen order
        if c > MIN_COUNT:
            vocab[key] = idx
            idx += 1
    return vocab


def transform_line(line: str, vocab, cat_imputation):
    """:43-66.  Returns (int_features f32[13], cat_features int64[26], label int)."""
    cols = line.split("\t")
    ints = []
    for i in range(1, NUM_INT + 1):
        v = cols[i]
        if v == "":                                       # :46-47
            v = "0"
        if int(v) < 0:                                    # :48-49
            v = "0"
        ints.append(int(v))
    int_array = np.array(ints).astype(np.float32)         # :51-52
    int_array = np.log(int_array + 1)                     # :53   float32 in, float32 out
    cat = [vocab.get(tok, 0) for tok in cat_tokens_of_line(line, cat_imputation)]     # :58-65  OOV -> 0
    return int_array, np.array(cat, dtype=np.int64), int(cols[0])


def transform(lines, vocab, cat_imputation):
    """The batch `read_tfrecord(...).batch(len(lines))` would deliver: (int_features f32[n,13], cat_features i64[n,26],
    label i64[n])."""
    n = len(lines)
    ints = np.zeros((n, NUM_INT), np.float32)
    cats = np.zeros((n, NUM_CAT), np.int64)
    labels = np.zeros(n, np.int64)
    for k, line in enumerate(lines):
        ints[k], cats[k], labels[k] = transform_line(line, vocab, cat_imputation)
    return ints, cats, labels


# --------------------------------------------------------------------------------------------------------------------
# build-defined token keys (the CUDA path's identity for a dictionary key; parity for it is by construction)
# --------------------------------------------------------------------------------------------------------------------

NEWLINE_BIT = 1 << 63


def missing_key(field: int) -> int:
    """Key of the null-imputation token of categorical field `field` (0-based).  Its low byte is zero, which no packed
    non-empty token has, and it differs per field exactly like the reference's 26 random strings (:11-12)."""
    return (field + 1) << 8


def pack_token(tok: str) -> int:
    """uint64 key of a non-imputed token: its bytes little-endian (first character in the low byte); a trailing newline
    (last column only) is carried as bit 63 instead of a ninth byte.  Tokens are ASCII, 1..8 bytes without the newline."""
    nl = tok.endswith("\n")
    body = tok[:-1] if nl else tok
    raw = body.encode("ascii")
    if not (1 <= len(raw) <= 8):
        raise ValueError(f"token {tok!r} does not fit 8 bytes")
    key = int.from_bytes(raw, "little")
    return key | NEWLINE_BIT if nl else key


def key_of(tok: str, cat_imputation) -> int:
    if tok in cat_imputation:
        return missing_key(cat_imputation.index(tok))
    return pack_token(tok)


def vocab_keys(vocab, cat_imputation) -> np.ndarray:
    """uint64[V]: the packed key of every vocabulary entry, in id order."""
    return np.array([key_of(t, cat_imputation) for t in vocab], dtype=np.uint64)


def token_keys(lines, cat_imputation) -> np.ndarray:
    """uint64[n,26]: the packed keys of every categorical token, in the reference's scan order."""
    return np.array([[key_of(t, cat_imputation) for t in cat_tokens_of_line(ln, cat_imputation)] for ln in lines],
                    dtype=np.uint64).reshape(len(lines), NUM_CAT)


def mix64(k: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser: the slot hash of the device-side vocabulary table."""
    k = np.asarray(k, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        k ^= k >> np.uint64(30)
        k *= np.uint64(0xBF58476D1CE4E5B9)
        k ^= k >> np.uint64(27)
        k *= np.uint64(0x94D049BB133111EB)
        k ^= k >> np.uint64(31)
    return k


def synth_tsv(n_lines: int, seed: int = 4, n_hot: int = 40, final_newline: bool = True, hot_seed: int = 4) -> bytes:
    """A synthetic Criteo-format TSV: label, 13 integer columns (some empty, some negative), 26 categorical columns of
    8 hex digits (some empty).  A small hot set per column makes counts cross the > 10 threshold; the same hot strings
    are shared between columns 3 and 7 and between 5 and 26 (the last), so the one-dictionary-for-all-fields quirk and
    the trailing-newline quirk both show."""
    rng = np.random.default_rng(seed)
    hrng = np.random.default_rng([hot_seed, 1])            # files that share hot_seed share their frequent tokens
    hot = [[f"{int(x):08x}" for x in hrng.integers(0, 2 ** 32, size=n_hot)] for _ in range(NUM_CAT)]
    hot[6] = hot[2]
    hot[25] = hot[4]
    rows = []
    for _ in range(n_lines):
        cols = [str(int(rng.random() < 0.25))]
        for _i in range(NUM_INT):
            u = rng.random()
            if u < 0.1:
                cols.append("")
            elif u < 0.15:
                cols.append(str(-int(rng.integers(1, 5))))
            elif u < 0.17:
                cols.append(str(int(rng.integers(2 ** 24, 2 ** 40))))       # beyond float32's exact integers
            else:
                cols.append(str(int(rng.integers(0, 2000))))
        for f in range(NUM_CAT):
            u = rng.random()
            if u < 0.08:
                cols.append("")
            elif u < 0.8:
                cols.append(hot[f][int(rng.integers(0, n_hot))])
            else:
                cols.append(f"{int(rng.integers(0, 2 ** 32)):08x}")
        rows.append("\t".join(cols))
    text = "\n".join(rows)
    if final_newline:
        text += "\n"
    return text.encode("ascii")
