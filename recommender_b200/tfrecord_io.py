"""Input side of the CTR models: Criteo TSV -> batches, on the GPU (mirror of /root/reference/ctr/tfrecord_io.py).

The reference goes text -> Python dict vocabulary -> TFRecord of serialized tensors -> tf.data (`build_vocab` :15-35,
`write_tfrecord` :38-75, `read_tfrecord` :78-96) and is bound by that Python loop.  Here the text is
copied to HBM once and every step of that chain is a kernel of librecsys_b200.so (csrc/criteo_input.cu):

    vocab = build_vocab(train_file)                        # same name, same rule: count > 10, ids in first-seen order
    for features, label in read_tfrecord(raw_file, vocab, batch_size):
        model(features)                                    # {'int_features' f32[B,13], 'cat_features' i64[B,26]}

    write_tfrecord(raw_file, output_file, vocab)           # optional, once: 160-byte binary records instead of text
    for features, label in read_tfrecord(output_file, batch_size=65536): ...

`read_tfrecord` takes the raw text or the record file.  The TFRecord/protobuf container is not reproduced, its content is.
torch only carries the device memory; there is no CPU path (a missing library raises).
"""
from __future__ import annotations

import os
import queue
import threading
from typing import Iterator, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .ops import _need_cuda, _ptr, _stream

num_int = 13        # ctr/tfrecord_io.py:8
num_cat = 26        # :9
total_cols = 40     # :10
MIN_COUNT = 10      # :31

ERROR_BITS = {1: "a line with fewer than 40 columns", 2: "a label / integer column that is not an integer",
              4: "a categorical token longer than 8 bytes", 8: "a non-ASCII byte in a categorical token",
              16: "a line longer than 1024 bytes"}


class CriteoFormatError(ValueError):
    pass


def _device(device) -> torch.device:
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise _lib.RecsysError("recommender_b200.tfrecord_io runs on a CUDA device only (there is no CPU path)")
    return dev


def to_device(text, device=None) -> torch.Tensor:
    """bytes / uint8 array / uint8 tensor -> uint8 tensor in HBM, padded so that the parser's 128-bit loads stay inside
    the allocation.  Returns a view of exactly len(text) bytes."""
    dev = _device(device)
    if isinstance(text, (bytes, bytearray, memoryview)):
        host = torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()) if len(text) else torch.empty(0, dtype=torch.uint8)
    elif isinstance(text, np.ndarray):
        host = torch.from_numpy(np.array(text, dtype=np.uint8))      # a copy: fixtures come in read-only
    else:
        host = text
    n = host.numel()
    buf = torch.zeros((n + 15) // 16 * 16 + 16, dtype=torch.uint8, device=dev)
    buf[:n].copy_(host, non_blocking=True)
    return buf[:n]


def index_lines(text: torch.Tensor, max_lines: Optional[int] = None) -> torch.Tensor:
    """int64[num_lines] byte offsets of the lines of `text` (rb_criteo_index_lines).  Synchronises once to learn the
    count; room is sized for well-formed lines (39 tabs and a label: more than 40 bytes) and the call is repeated with
    the exact count for a file of shorter ones."""
    _need_cuda(text)
    nbytes = text.numel()
    nbytes_ws = lib.rb_criteo_index_workspace_bytes(nbytes)
    if nbytes_ws == 0:
        raise _lib.RecsysError("rb_criteo_index_workspace_bytes rejected the size: feed the file in chunks below 2 GiB")
    ws = torch.empty(nbytes_ws, dtype=torch.uint8, device=text.device)
    count = torch.zeros(1, dtype=torch.int64, device=text.device)
    room = nbytes // total_cols + 1 if max_lines is None else int(max_lines)
    while True:
        starts = torch.empty(room, dtype=torch.int64, device=text.device)
        check(lib.rb_criteo_index_lines(_ptr(text), nbytes, room, _ptr(starts), _ptr(count), _ptr(ws), ws.numel(), _stream()),
              "rb_criteo_index_lines")
        n = int(count.item())
        if n <= room:
            return starts[:n]
        if max_lines is not None:
            raise _lib.RecsysError(f"{n} lines but room for {max_lines}")
        room = n


def check_errors(flag: torch.Tensor) -> None:
    bits = int(flag.item())
    if bits:
        flag.zero_()
        raise CriteoFormatError("malformed Criteo text: " + "; ".join(msg for b, msg in ERROR_BITS.items() if bits & b))


class Vocab:
    """The reference's `cat_fea_vocab` dict (:28-33) as device arrays: `keys` uint64-as-int64[V] in id order and the
    open-addressing table the parser probes."""

    def __init__(self, keys: torch.Tensor):
        _need_cuda(keys)
        self.keys = keys.contiguous()
        n = self.keys.numel()
        cap = 16
        while cap < 2 * n + 1:
            cap *= 2
        self.capacity = cap
        self.table_keys = torch.empty(cap, dtype=torch.int64, device=keys.device)
        self.table_vals = torch.empty(cap, dtype=torch.int32, device=keys.device)
        check(lib.rb_vocab_table_build(_ptr(self.keys), n, _ptr(self.table_keys), _ptr(self.table_vals), cap, _stream()),
              "rb_vocab_table_build")

    def __len__(self) -> int:
        return self.keys.numel()

    def lookup(self, tokens: torch.Tensor) -> torch.Tensor:
        """ids of packed token keys, 0 when absent (:61-64)."""
        _need_cuda(tokens)
        tokens = tokens.contiguous()
        out = torch.empty(tokens.shape, dtype=torch.int64, device=tokens.device)
        check(lib.rb_vocab_lookup(_ptr(tokens), tokens.numel(), _ptr(self.table_keys), _ptr(self.table_vals), self.capacity,
                                  _ptr(out), _stream()), "rb_vocab_lookup")
        return out

    def save(self, path: str) -> None:
        """Stands in for the pickle of :34-35 (keys in id order)."""
        np.save(path, self.keys.cpu().numpy().view(np.uint64))

    @classmethod
    def load(cls, path: str, device=None) -> "Vocab":
        return cls(torch.from_numpy(np.load(path).view(np.int64)).to(_device(device)))


def parse(text: torch.Tensor, vocab: Optional[Vocab] = None, *, line_start: Optional[torch.Tensor] = None, want_tokens=False,
          raise_on_error=True):
    """The body of `write_tfrecord` (:43-66) for every line of `text`, as one kernel (rb_criteo_parse).

    Returns ({'int_features': f32[n,13], 'cat_features': i64[n,26]}, label i64[n]); with vocab=None or want_tokens the
    dict also carries 'cat_tokens' (the packed dictionary keys, i64 bit patterns of uint64)."""
    _need_cuda(text)
    if line_start is None:
        line_start = index_lines(text)
    n = line_start.numel()
    dev = text.device
    label = torch.empty(n, dtype=torch.int64, device=dev)
    ints = torch.empty(n, num_int, dtype=torch.float32, device=dev)
    tokens = torch.empty(n, num_cat, dtype=torch.int64, device=dev) if (want_tokens or vocab is None) else None
    cats = torch.empty(n, num_cat, dtype=torch.int64, device=dev) if vocab is not None else None
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib.rb_criteo_parse(_ptr(text), text.numel(), _ptr(line_start), n, _ptr(label), _ptr(ints), _ptr(tokens), _ptr(cats),
                              _ptr(vocab.table_keys) if vocab is not None else None,
                              _ptr(vocab.table_vals) if vocab is not None else None,
                              vocab.capacity if vocab is not None else 0, _ptr(flag), _stream()), "rb_criteo_parse")
    features = {"int_features": ints}
    if raise_on_error:
        check_errors(flag)               # synchronises
    else:
        features["error_flag"] = flag    # the caller reads the bits (ERROR_BITS) when it chooses to synchronise
    if cats is not None:
        features["cat_features"] = cats
    if tokens is not None:
        features["cat_tokens"] = tokens
    return features, label


def vocab_from_tokens(tokens: torch.Tensor, min_count: int = MIN_COUNT) -> Vocab:
    """`build_vocab`'s counting and numbering (:24-33) over packed tokens in scan order (rb_vocab_build)."""
    _need_cuda(tokens)
    tokens = tokens.contiguous()
    n = tokens.numel()
    nbytes = lib.rb_vocab_build_workspace_bytes(n)
    if nbytes == 0:
        raise _lib.RecsysError("rb_vocab_build_workspace_bytes rejected the size (2^31 tokens per call)")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=tokens.device)
    out = torch.empty(max(n // (min_count + 1), 1), dtype=torch.int64, device=tokens.device)   # a kept token occurs > min_count times
    count = torch.zeros(1, dtype=torch.int64, device=tokens.device)
    check(lib.rb_vocab_build(_ptr(tokens), n, int(min_count), _ptr(out), out.numel(), _ptr(count), _ptr(ws), ws.numel(), _stream()),
          "rb_vocab_build")
    return Vocab(out[: int(count.item())].clone())


def _fill_chunks(fh, view: np.ndarray) -> Iterator[int]:
    """Host half of the chunked reader: fills `view` (uint8) from the binary file `fh` and yields `cut`, the number of
    leading bytes that form whole lines.  When the consumer comes back the unfinished tail is moved to the front and
    reading continues behind it.  The file's last line may lack its newline."""
    rest = 0
    while True:
        got = fh.readinto(memoryview(view)[rest:]) or 0
        total = rest + got
        if total == 0:
            return
        if got == 0:
            cut = total                                       # end of file: a last line without its newline
        else:
            cut, lo, step = 0, total, 1 << 16
            while cut == 0 and lo > 0:                        # last newline, searched backwards a block at a time
                lo = max(0, lo - step)
                hits = np.flatnonzero(view[lo:min(lo + step, total)] == 10)
                if hits.size:
                    cut = lo + int(hits[-1]) + 1
            if cut == 0:
                if total == view.size:
                    raise CriteoFormatError(f"a line longer than chunk_bytes = {view.size}")
                rest = total                                  # no complete line yet: keep reading
                continue
        yield cut
        rest = total - cut
        if rest:
            view[:rest] = view[cut:total].copy()
        if got == 0:
            return


def _device_chunks(path: str, chunk_bytes: int, dev: torch.device) -> Iterator[torch.Tensor]:
    """Whole-line chunks of a text file as uint8 tensors in HBM: the file is read straight into ONE pinned staging
    buffer (no intermediate bytes objects) and copied from there."""
    size = max(1, min(int(chunk_bytes), os.path.getsize(path)))
    stage = torch.empty(size, dtype=torch.uint8, pin_memory=True)
    with open(path, "rb") as fh:
        for cut in _fill_chunks(fh, stage.numpy()):
            out = to_device(stage[:cut], dev)
            torch.cuda.current_stream(dev).synchronize()      # the staging buffer is about to be overwritten
            yield out


def build_vocab(train_file: str, *, device=None, chunk_bytes: int = 1 << 30, min_count: int = MIN_COUNT,
                save_to: Optional[str] = None) -> Vocab:
    """ctr/tfrecord_io.py:15-35.  The file is parsed chunk by chunk; the packed tokens of all chunks stay in HBM
    (8 bytes x 26 per line) and are counted in one pass."""
    dev = _device(device)
    parts = []
    for block in _device_chunks(train_file, chunk_bytes, dev):
        features, _ = parse(block, None)
        parts.append(features["cat_tokens"])
    tokens = torch.cat(parts) if parts else torch.empty(0, num_cat, dtype=torch.int64, device=dev)
    vocab = vocab_from_tokens(tokens, min_count)
    if save_to is not None:
        os.makedirs(os.path.dirname(os.path.abspath(save_to)), exist_ok=True)
        vocab.save(save_to)
    return vocab


# ---- the preprocessed record file (stands in for the TFRecord of ctr/tfrecord_io.py:38-75) -----------------------------
# One fixed-size record per line, in file order: label i32 | int_features f32[13] | cat_features i32[26] = 160 bytes,
# behind a 64-byte header.  Same content as the reference's tf.train.Example (two serialized tensors and the label,
# :66-73) without the protobuf framing; ids and labels are stored in 32 bits (a dictionary has fewer than 2^31
# entries) and widened to the reference's int64 (:82, :88) on the GPU.  The host never touches a record: records move
# disk -> pinned -> HBM as bytes and are split into the three tensors on the GPU.

RECORD_MAGIC = b"RBCRITEO"
RECORD_VERSION = 2
RECORD_HEADER = 64
RECORD_BYTES = 4 + 4 * num_int + 4 * num_cat


def _record_header() -> bytes:
    head = RECORD_MAGIC + np.array([RECORD_VERSION, num_int, num_cat, RECORD_BYTES], dtype="<i4").tobytes()
    return head + b"\0" * (RECORD_HEADER - len(head))


def _pack_records(label: torch.Tensor, ints: torch.Tensor, cats: torch.Tensor) -> torch.Tensor:
    """(label i64[n], int_features f32[n,13], cat_features i64[n,26]) -> uint8[n, 160] on the same device."""
    n = label.numel()
    if n and (int(label.abs().max()) >= 2 ** 31 or int(cats.max()) >= 2 ** 31 or int(cats.min()) < 0):
        raise CriteoFormatError("a label or id does not fit the record file's 32-bit fields")
    return torch.cat([label.to(torch.int32).reshape(n, 1).view(torch.uint8),
                      ints.contiguous().view(torch.uint8).reshape(n, 4 * num_int),
                      cats.to(torch.int32).view(torch.uint8).reshape(n, 4 * num_cat)], dim=1)


def _split_records(raw: torch.Tensor):
    """uint8[n, 160] in HBM -> ({'int_features' f32, 'cat_features' i64}, label i64)."""
    def cols(a: int, b: int, dtype) -> torch.Tensor:
        out = torch.empty(raw.shape[0], b - a, dtype=torch.uint8, device=raw.device)     # fresh, so the dtype view is aligned
        out.copy_(raw[:, a:b])
        return out.view(dtype)

    label = cols(0, 4, torch.int32).reshape(-1).to(torch.int64)
    ints = cols(4, 4 + 4 * num_int, torch.float32)
    cats = cols(4 + 4 * num_int, RECORD_BYTES, torch.int32).to(torch.int64)
    return {"int_features": ints, "cat_features": cats}, label


def write_tfrecord(raw_file: str, output_file: str, vocab: Vocab, *, device=None, chunk_bytes: int = 1 << 28) -> int:
    """ctr/tfrecord_io.py:38-75: raw Criteo text -> the preprocessed record file, once, so that every epoch reads
    160-byte records instead of parsing text.  Returns the number of records."""
    dev = _device(device)
    n = 0
    with open(output_file, "wb") as out:
        out.write(_record_header())
        for block in _device_chunks(raw_file, chunk_bytes, dev):
            features, label = parse(block, vocab)
            raw = _pack_records(label, features["int_features"], features["cat_features"])
            out.write(raw.cpu().numpy())              # the array's buffer goes to the file as is
            n += label.numel()
    return n


def _is_record_file(path: str) -> bool:
    with open(path, "rb") as fh:
        head = fh.read(len(RECORD_MAGIC) + 16)
    if head[: len(RECORD_MAGIC)] != RECORD_MAGIC:
        return False
    fields = np.frombuffer(head[len(RECORD_MAGIC):], dtype="<i4")
    if fields.size != 4 or fields.tolist() != [RECORD_VERSION, num_int, num_cat, RECORD_BYTES]:
        raise CriteoFormatError(f"{path}: record file of another version / layout")
    return True


def _batch_ranges(n: int, batch_size: int, rank: int, world: int, drop_remainder: bool):
    """Record ranges [s, e) of the batches rank `rank` of `world` reads: batch k of the file goes to rank k mod world
    (data-parallel replicas see disjoint batches, no exchange).  With drop_remainder the ranks also get EQUAL batch
    counts — a trailing group of fewer than `world` batches is dropped — so collectives in the step stay matched."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    total = n // batch_size if drop_remainder else -(-n // batch_size)
    if drop_remainder:
        total -= total % world
    for k in range(rank, total, world):
        yield k * batch_size, min((k + 1) * batch_size, n)


def _pinned(*shape) -> torch.Tensor:
    return torch.empty(*shape, dtype=torch.uint8, pin_memory=True)


_stage_pool: dict = {}        # (rows, bytes per row) -> idle pinned staging buffers; page-locking costs milliseconds apiece


def _take_stages(rows: int, count: int):
    idle = _stage_pool.setdefault((rows, RECORD_BYTES), [])
    return [idle.pop() if idle else _pinned(rows, RECORD_BYTES) for _ in range(count)]


def _return_stages(rows: int, stages) -> None:
    _stage_pool.setdefault((rows, RECORD_BYTES), []).extend(stages)


class _CopyDone:
    """Marks the point on the current stream after which a staging buffer may be overwritten."""

    def __init__(self):
        self._ev = torch.cuda.Event()
        self._ev.record()

    def wait(self):
        self._ev.synchronize()


_STAGES = 3       # pinned staging buffers of the record reader: one being filled, one in flight to the GPU, one spare


def _read_records(path: str, batch_size: int, dev: torch.device, drop_remainder: bool, rank: int = 0, world: int = 1):
    """Record file -> batches.  A reader thread reads each batch's records from the file (page cache) straight into a
    pinned staging buffer while the caller's thread issues the H2D copy of the previous one and splits it on the GPU;
    a buffer goes back to the reader once the copy that read it has completed, and to a pool when the epoch ends."""
    size = os.path.getsize(path) - RECORD_HEADER
    if size < 0 or size % RECORD_BYTES:
        raise CriteoFormatError(f"{path}: truncated record file")
    n = size // RECORD_BYTES
    if n == 0:
        return
    stage = _take_stages(batch_size, _STAGES)
    free_q: "queue.Queue" = queue.Queue()
    full_q: "queue.Queue" = queue.Queue()
    for i in range(_STAGES):
        free_q.put(i)

    def reader():
        try:
            with open(path, "rb", buffering=0) as fh:
                at = -1
                for s, e in _batch_ranges(n, batch_size, rank, world, drop_remainder):
                    i = free_q.get()
                    if i is None:                             # the consumer went away
                        return
                    if at != s:
                        fh.seek(RECORD_HEADER + s * RECORD_BYTES)
                    view = memoryview(stage[i].numpy()).cast("B")[: (e - s) * RECORD_BYTES]
                    got = 0
                    while got < len(view):                    # read() releases the GIL; short reads are legal
                        k = fh.readinto(view[got:])
                        if not k:
                            raise CriteoFormatError(f"{path}: file shrank while it was being read")
                        got += k
                    at = e
                    full_q.put((i, e - s))
            full_q.put(None)
        except BaseException as exc:                          # surfaces in the consumer's thread
            full_q.put(exc)

    thread = threading.Thread(target=reader, name="rb-record-reader", daemon=True)
    thread.start()
    in_flight = []                                            # (buffer, copy-done marker), oldest first
    try:
        while True:
            item = full_q.get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            i, rows = item
            raw = stage[i][:rows].to(dev, non_blocking=True)
            in_flight.append((i, _CopyDone()))
            if len(in_flight) >= _STAGES - 1:
                j, done = in_flight.pop(0)
                done.wait()
                free_q.put(j)
            yield _split_records(raw)
    finally:
        free_q.put(None)
        thread.join(timeout=5)
        for _, done in in_flight:
            done.wait()
        if not thread.is_alive():
            _return_stages(batch_size, stage)


def read_tfrecord(tfrecord_file: str, vocab: Optional[Vocab] = None, batch_size: int = 1024, *, device=None,
                  chunk_bytes: int = 1 << 28, drop_remainder: bool = False, rank: int = 0,
                  world: int = 1) -> Iterator[Tuple[dict, torch.Tensor]]:
    """ctr/tfrecord_io.py:78-96 followed by `.batch(batch_size)` (ctr/train.py:59-61): yields
    ({'int_features': f32[B,13], 'cat_features': i64[B,26]}, label i64[B]) in file order.

    `tfrecord_file` is either a record file written by `write_tfrecord`, or the RAW Criteo text itself (then `vocab`
    is required and the text is parsed on the fly: `read_tfrecord(write_tfrecord(raw))` without the file in between).
    One process per GPU: rank r of `world` gets batches r, r + world, ... of a record file (no exchange between the
    replicas); raw text is a single-process format."""
    dev = _device(device)
    batch_size = int(batch_size)
    if _is_record_file(tfrecord_file):
        yield from _read_records(tfrecord_file, batch_size, dev, drop_remainder, rank, world)
        return
    if vocab is None:
        raise ValueError("reading raw Criteo text needs the vocabulary (build_vocab)")
    if world != 1:
        raise NotImplementedError("raw text is read by one process; convert it once with write_tfrecord for several ranks")
    carry = None
    for block in _device_chunks(tfrecord_file, chunk_bytes, dev):
        features, label = parse(block, vocab)
        ints, cats = features["int_features"], features["cat_features"]
        if carry is not None:
            ints, cats, label = torch.cat([carry[0], ints]), torch.cat([carry[1], cats]), torch.cat([carry[2], label])
        n = label.numel()
        full = n // batch_size * batch_size
        for s in range(0, full, batch_size):
            yield {"int_features": ints[s:s + batch_size], "cat_features": cats[s:s + batch_size]}, label[s:s + batch_size]
        carry = (ints[full:], cats[full:], label[full:]) if full < n else None
    if carry is not None and not drop_remainder:
        yield {"int_features": carry[0], "cat_features": carry[1]}, carry[2]
