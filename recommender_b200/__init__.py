"""recommender_b200 — the CTR embedding hot path of neoyinyao/Recommender on B200 (sm_100a).

Public surface (mirrors ctr/layers.py, ctr/model.py and the Keras optimizer the reference builds):
    layers.Embedding / MLP / DotInteraction, model.DeepFM / DLRM, optimizers.Adam / Adagrad / SGD,
    ops.* (one function per C-ABI entry point of include/recsys_b200.h).
Importing the package does not load the CUDA library; the first op call does, and raises if
librecsys_b200.so is missing (no CPU fallback).
"""
__version__ = "0.1.0"
