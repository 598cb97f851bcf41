"""A ~20-function stand-in for `tensorflow` over torch-CPU tensors.  NOT TensorFlow.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Its only job: let
/root/reference/ctr/model.py and ctr/layers.py be imported BYTE-FOR-BYTE in the
build container (TensorFlow is absent), so their forward graph — and torch
autograd's gradients of that exact graph — can be frozen into tests/golden/.
Each function restates the documented TF op it is named after; only the ops the
reference's hot-path files call are present.
"""
import torch as _t

from . import keras  # noqa: F401  (`from tensorflow import keras`)
from . import linalg, nn  # noqa: F401
from . import io, train  # noqa: F401  (ctr/tfrecord_io.py's writer side)

float32 = _t.float32
int32 = _t.int32
int64 = _t.int64
bool = _t.bool  # noqa: A001  (tf.bool)


def _as(x):
    return x if isinstance(x, _t.Tensor) else _t.as_tensor(x)


def reshape(x, shape):
    return _as(x).reshape(tuple(int(s) for s in shape))


def square(x):
    return x * x


def reduce_sum(x, axis=None, keepdims=False):
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def concat(values, axis):
    return _t.cat(list(values), dim=axis)


def squeeze(x, axis=None):
    return x.squeeze() if axis is None else x.squeeze(axis)


def expand_dims(x, axis):
    return x.unsqueeze(axis)


def repeat(x, repeats, axis):
    """tf.repeat with one repeat count for the whole axis (dien/layers.py:47: repeats=[his_len] on a length-1 axis)."""
    r = repeats[0] if isinstance(repeats, (list, tuple)) else repeats
    return _t.repeat_interleave(x, int(r), dim=axis)


def shape(x):
    return tuple(x.shape)


def matmul(a, b, transpose_a=False, transpose_b=False):
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


def ones_like(x):
    return _t.ones_like(x)


def zeros_like(x):
    return _t.zeros_like(x)


def cast(x, dtype):
    return x.to(dtype)


def where(condition, x, y):
    return _t.where(condition, x, y)


def boolean_mask(tensor, mask):
    """mask has the same shape as tensor here (ctr/layers.py:41): returns the kept elements,
    flattened in row-major order."""
    return tensor[mask]
