"""`tf.io` stand-ins for ctr/tfrecord_io.py:38-75 — NOT TensorFlow, test infrastructure only.

The TFRecord container (protobuf framing of tf.train.Example) is not restated: `serialize_tensor` hands the numpy array
through untouched and `TFRecordWriter.write` keeps the records in memory (WRITTEN[path]), so the values the
reference's own `write_tfrecord` computes per line can be frozen as a fixture (tests/golden/make_golden_criteo.py).
"""
import numpy as _np

WRITTEN = {}


class _Serialized:
    def __init__(self, array):
        self._a = _np.array(array)

    def numpy(self):
        return self._a


def serialize_tensor(tensor):
    return _Serialized(tensor)


class TFRecordWriter:
    def __init__(self, path):
        self.path = path
        WRITTEN[path] = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def write(self, record):
        WRITTEN[self.path].append(record)
