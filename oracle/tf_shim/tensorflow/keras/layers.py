"""tf.keras.layers subset for the shim (TEST INFRASTRUCTURE, see tensorflow/__init__.py)."""
import math as _math

import torch as _t


class Layer:
    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


def _activation(name):
    if name is None:
        return lambda x: x
    if callable(name):
        return name
    return {"relu": _t.relu, "sigmoid": _t.sigmoid}[name]


class Dense(Layer):
    """y = act(x @ kernel + bias); kernel [in, units] Glorot-uniform, bias zeros; built lazily."""

    def __init__(self, units, activation=None):
        super().__init__()
        self.units, self.activation = units, _activation(activation)
        self.kernel = self.bias = None

    def build(self, in_dim):
        lim = _math.sqrt(6.0 / (in_dim + self.units))
        self.kernel = ((_t.rand(in_dim, self.units) * 2 - 1) * lim).requires_grad_()
        self.bias = _t.zeros(self.units, requires_grad=True)

    def call(self, x, training=None):
        if self.kernel is None:
            self.build(x.shape[-1])
        return self.activation(x @ self.kernel + self.bias)


class Embedding(Layer):
    """embeddings [input_dim, output_dim] ~ U(-0.05, 0.05); call = gather; mask_zero only
    changes compute_mask (x != 0) — row 0 stays a real trainable row."""

    def __init__(self, input_dim, output_dim, mask_zero=False):
        super().__init__()
        self.mask_zero = mask_zero
        self.embeddings = ((_t.rand(input_dim, output_dim) - 0.5) * 0.1).requires_grad_()

    def call(self, x):
        x = _t.as_tensor(x)
        if x.numel() and (int(x.min()) < 0 or int(x.max()) >= self.embeddings.shape[0]):
            raise IndexError("InvalidArgument: indices out of range")   # TF CPU kernel behaviour
        return self.embeddings[x.long()]

    def compute_mask(self, x, mask=None):
        return (_t.as_tensor(x) != 0) if self.mask_zero else None


class AbstractRNNCell(Layer):
    """Base-class stub so dien/layers.py imports (the recurrent heads are out of scope)."""


class BatchNormalization(Layer):
    """Inference-mode identity at initial statistics (mean 0, var 1, eps 1e-3) — only so that
    dien's BaseModel constructs; the BN-MLP head is out of scope and never goldened."""

    def call(self, x, training=False):
        return x / _math.sqrt(1.0 + 1e-3)
