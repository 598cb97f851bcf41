"""tf.nn subset for the shim (TEST INFRASTRUCTURE, see tensorflow/__init__.py)."""
import torch as _t


def sigmoid(x):
    return _t.sigmoid(x)


def relu(x):
    return _t.relu(x)
