"""tf.keras subset for the shim (TEST INFRASTRUCTURE, see tensorflow/__init__.py).

Layer / Model: plain callables that forward __call__ to call().  Dense and Embedding own
torch leaf tensors (requires_grad) initialised with the Keras defaults (Glorot uniform /
zeros; U(-0.05, 0.05)); the golden script overwrites them with the oracle's arrays so both
sides load identical weights (TF's RNG streams cannot be matched).
"""
import math as _math

import torch as _t

from . import layers  # noqa: F401


class Model(layers.Layer):
    pass
