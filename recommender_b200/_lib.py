"""ctypes binding of librecsys_b200.so (include/recsys_b200.h).

The product path has NO CPU fallback: if the CUDA library is missing or a symbol the header
declares is absent, importing the ops raises.  Tensors cross the boundary as raw device
pointers + sizes (torch is only the carrier of memory and streams).
"""
from __future__ import annotations

import ctypes as C
import os
import re

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RB_LIB_PATH") or os.path.join(PKG, "lib", "librecsys_b200.so")   # RB_LIB_PATH: tuning builds
HEADER = os.path.join(os.path.dirname(PKG), "include", "recsys_b200.h")

RB_MAX_GRAD_SOURCES = 16
RB_MAX_LOOKUP_GROUPS = 8
RB_MAX_DENSE_TENSORS = 32
RB_MAX_RANKS = 8
RB_IPC_HANDLE_BYTES = 64

# enums (recsys_b200.h)
RB_I32, RB_I64 = 0, 1
RB_F32, RB_BF16, RB_BF16_ONES = 0, 1, 2
RB_POOL_SUM, RB_POOL_MEAN, RB_POOL_MASKED_MEAN = 1, 2, 3
RB_OPT_SGD, RB_OPT_ADAGRAD, RB_OPT_ADAM_LAZY, RB_OPT_ADAM_TF_DENSE = 0, 1, 2, 3
RB_SCALE_NONE, RB_SCALE_MEAN, RB_SCALE_MASKED_MEAN, RB_SCALE_MASKED = 0, 1, 2, 3
RB_ROW_CACHE_L2, RB_ROW_CACHE_L1, RB_ROW_CACHE_AUTO = 0, 1, 2
ROW_CACHE_ENUM = {False: RB_ROW_CACHE_L2, True: RB_ROW_CACHE_L1, "auto": RB_ROW_CACHE_AUTO, None: RB_ROW_CACHE_L2}
RB_ACT_NONE, RB_ACT_RELU, RB_ACT_SIGMOID = 0, 1, 2
ACT_ENUM = {None: RB_ACT_NONE, "relu": RB_ACT_RELU, "sigmoid": RB_ACT_SIGMOID}

OPTIMIZER_ENUM = {"sgd": RB_OPT_SGD, "adagrad": RB_OPT_ADAGRAD, "adam_lazy": RB_OPT_ADAM_LAZY,
                  "adam_tf_dense": RB_OPT_ADAM_TF_DENSE}
POOL_ENUM = {"sum": RB_POOL_SUM, "mean": RB_POOL_MEAN, "masked_mean": RB_POOL_MASKED_MEAN}
SCALE_ENUM = {"none": RB_SCALE_NONE, "mean": RB_SCALE_MEAN, "masked_mean": RB_SCALE_MASKED_MEAN, "masked": RB_SCALE_MASKED}


class RbOptParams(C.Structure):
    _fields_ = [("optimizer", C.c_int32), ("step", C.c_int32), ("lr", C.c_float), ("beta_1", C.c_float),
                ("beta_2", C.c_float), ("epsilon", C.c_float), ("alpha_t_dev", C.c_void_p)]


class RbDinHistory(C.Structure):
    _fields_ = [("table0", C.c_void_p), ("rows0", C.c_int64), ("D0", C.c_int32), ("idx0", C.c_void_p),
                ("table1", C.c_void_p), ("rows1", C.c_int64), ("D1", C.c_int32), ("idx1", C.c_void_p),
                ("idx_type", C.c_int32), ("mask", C.c_void_p), ("mask_type", C.c_int32), ("B", C.c_int64), ("L", C.c_int32)]


class RbGradSource(C.Structure):
    _fields_ = [("num_src", C.c_int32), ("scale_mode", C.c_int32),
                ("src", C.c_void_p * RB_MAX_GRAD_SOURCES),
                ("bag_stride", C.c_int64 * RB_MAX_GRAD_SOURCES),
                ("pos_stride", C.c_int64 * RB_MAX_GRAD_SOURCES),
                ("mask_idx", C.c_void_p), ("count", C.c_void_p), ("fm_g", C.c_void_p), ("fm_s", C.c_void_p)]


class RbLookupGroup(C.Structure):
    _fields_ = [("idx", C.c_void_p), ("idx_type", C.c_int32), ("L", C.c_int32), ("n", C.c_int64),
                ("field_row_offset", C.c_void_p), ("hash_mod", C.c_int64), ("grad", RbGradSource)]


class RbDenseSlot(C.Structure):
    _fields_ = [("param", C.c_void_p), ("state0", C.c_void_p), ("state1", C.c_void_p), ("grad", C.c_void_p), ("n", C.c_int64),
                ("shadow_bf16", C.c_void_p)]


_p, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol include/recsys_b200.h declares
SIGNATURES = {
    "rb_version": (C.c_int, []),
    "rb_last_error": (C.c_char_p, []),
    "rb_kernel_launches": (C.c_uint64, []),
    "rb_set_pdl": (None, [_i32]),
    "rb_adam_alpha_t": (C.c_float, [_f, _f, _f, _i32]),
    "rb_gather_fwd": (C.c_int, [_p, _i64, _i32, _p, _i32, _i64, _i32, _p, _i64, _p, _i64, _p, _p]),
    "rb_check_indices": (C.c_int, [_p, _i32, _i64, _i32, _p, _i64, _p, _p]),
    "rb_bag_pool_fwd": (C.c_int, [_p, _i64, _i32, _p, _i32, _i64, _i32, _p, _i64, _i32, _p, _p, _i64, _p, _p, _p]),
    "rb_gather_fm_fwd": (C.c_int, [_p, _i64, _i32, _p, _i32, _i64, _i32, _p, _i64, _p, _p, _p, _p, _p]),
    "rb_gather_fm_deep_fwd": (C.c_int, [_p, _i64, _i32, _p, _i32, _i64, _i32, _p, _i64, _p, _i32, _i64, _p, _i32, _p, _p, _p, _p, _p]),
    "rb_dot_interaction_fwd": (C.c_int, [_p, _p, _i64, _p, _i32, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i64, _i32, _p, _p]),
    "rb_dot_interaction_bwd": (C.c_int, [_p, _p, _i64, _p, _i32, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i64, _p, _p, _i32, _p,
                                         _p]),
    "rb_sparse_bwd_update_workspace_bytes": (C.c_size_t, [_i64, _i32, _i64]),
    "rb_sparse_bwd_update": (C.c_int, [_p, _p, _p, _i64, _i32, _p, _i32, _i64, _i32, _p, _i64,
                                       C.POINTER(RbGradSource), C.POINTER(RbOptParams), _p, C.c_size_t, _p, _p]),
    "rb_sparse_bwd_update_groups": (C.c_int, [_p, _p, _p, _i64, _i32, C.POINTER(RbLookupGroup), _i32,
                                              C.POINTER(RbOptParams), _p, C.c_size_t, _p, _p]),
    "rb_sparse_bwd_prepare": (C.c_int, [_i64, _i32, C.POINTER(RbLookupGroup), _i32, _p, C.c_size_t, _p, C.POINTER(C.c_int32), _p, _p]),
    "rb_sparse_bwd_apply": (C.c_int, [_p, _p, _p, _i64, _i32, C.POINTER(RbLookupGroup), _i32, C.POINTER(RbOptParams), _p,
                                      C.c_size_t, _i32, _p]),
    "rb_sparse_bwd_apply_ex": (C.c_int, [_p, _p, _p, _i64, _i32, C.POINTER(RbLookupGroup), _i32, C.POINTER(RbOptParams), _p,
                                         C.c_size_t, _i32, _i32, _p]),
    "rb_sparse_bwd_mark_singletons": (C.c_int, [_i64, _i32, _i64, _p, C.c_size_t, C.POINTER(C.c_int32), _p, _p]),
    "rb_dot_interaction_bwd_update": (C.c_int, [_p, _i64, _p, _i32, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i64, _p, _p,
                                                _p, _p, _p, C.POINTER(RbOptParams), _i32, _p, _p]),
    "rb_replicated_rows_update": (C.c_int, [_p, _p, _p, _i64, _i32, _p, _i32, _i64, C.POINTER(RbOptParams), _p, _p]),
    "rb_sparse_bwd_dedup": (C.c_int, [_i64, _i32, _p, _i32, _i64, _i32, _p, _i64, C.POINTER(RbGradSource),
                                      _p, _p, _p, _p, C.c_size_t, _p, _p]),
    "rb_dense_opt_step": (C.c_int, [C.POINTER(RbDenseSlot), _i32, C.POINTER(RbOptParams), _p]),
    "rb_colsum_workspace_bytes": (C.c_size_t, [_i64, _i32]),
    "rb_colsum": (C.c_int, [_p, _i32, _i64, _i32, _i64, _p, _p, C.c_size_t, _p]),
    "rb_shared_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), _p]),
    "rb_shared_free": (C.c_int, [_p]),
    "rb_ipc_open": (C.c_int, [_p, C.POINTER(C.c_void_p)]),
    "rb_ipc_close": (C.c_int, [_p]),
    "rb_enable_peer_access": (C.c_int, [_i32]),
    "rb_dot_interaction_fwd_sharded": (C.c_int, [_p, _i32, _i64, _p, _i32, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i64, _p,
                                                 _p, _p]),
    "rb_dot_interaction_fwd_sharded_rep": (C.c_int, [_p, _i32, _i64, _p, _i32, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i64, _p,
                                                     _p, _i64, _p, _p, _p]),
    "rb_dot_interaction_bwd_sharded": (C.c_int, [_p, _i32, _i64, _p, _i32, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i64,
                                                 _p, _p, _p, _p]),
    "rb_dot_interaction_bwd_sharded_split": (C.c_int, [_p, _i32, _i64, _p, _i32, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p, _i32, _i64,
                                                       _p, _p, _p, _p, _p, _i32, _p]),
    "rb_p2p_collect_keys": (C.c_int, [_i32, _i32, _i64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _i64, _i64,
                                      _p, C.c_size_t, _i32, _p, _p, _p]),
    "rb_sparse_bwd_prepare_collected": (C.c_int, [_i64, _i32, _i64, _p, C.c_size_t, C.POINTER(C.c_int32), _p]),
    "rb_sparse_bwd_apply_p2p": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i64, _i32, C.POINTER(C.c_void_p), _i64, _p,
                                          C.POINTER(RbOptParams), _p, C.c_size_t, _i32, _p, _p]),
    "rb_bce_workspace_bytes": (C.c_size_t, [_i64]),
    "rb_bce_clipped": (C.c_int, [_p, _p, _i32, _i64, _p, _p, _p, C.c_size_t, _p]),
    "rb_dense_fwd": (C.c_int, [_p, _i64, _i32, _i64, _p, _i32, _i64, _p, _i32, _p, _i32, _i64, _p]),
    "rb_dense_bwd_input": (C.c_int, [_p, _i64, _i32, _i64, _p, _i32, _i64, _p, _i64, _p]),
    "rb_dense_bwd_weight_workspace_bytes": (C.c_size_t, [_i64, _i32, _i32]),
    "rb_dense_bwd_weight": (C.c_int, [_p, _i64, _i32, _i64, _p, _i32, _i64, _p, _i64, _p, C.c_size_t, _p]),
    "rb_dense_head_fwd": (C.c_int, [_p, _i64, _i32, _i64, _p, _p, _i32, _p, _p]),
    "rb_dense_head_bwd_workspace_bytes": (C.c_size_t, [_i64, _i32]),
    "rb_dense_head_bwd": (C.c_int, [_p, _p, _i32, _p, _i64, _i32, _i64, _p, _p, _i64, _p, _p, _p, _p, C.c_size_t, _p]),
    "rb_dense_act_bwd": (C.c_int, [_p, _p, _i32, _i64, _p, _p]),
    "rb_dense_head_bce_workspace_bytes": (C.c_size_t, [_i64, _i32]),
    "rb_dense_head_bce": (C.c_int, [_p, _i64, _i32, _i64, _p, _p, _p, _i32, _p, _p, _p, _i64, _p, _p, _p, _p, C.c_size_t, _p]),
    "rb_dense_act_bwd_bf16": (C.c_int, [_p, _p, _i32, _i64, _p, _p]),
    "rb_din_workspace_bytes": (C.c_size_t, [_i64]),
    "rb_din_offsets": (C.c_int, [C.POINTER(RbDinHistory), _p, _p, C.c_size_t, _p]),
    "rb_din_build_features": (C.c_int, [C.POINTER(RbDinHistory), _p, _p, _p, _i64, _p]),
    "rb_din_pool_fwd": (C.c_int, [C.POINTER(RbDinHistory), _p, _p, _p, _p]),
    "rb_din_pool_bwd_weights": (C.c_int, [C.POINTER(RbDinHistory), _p, _p, _p, _p]),
    "rb_din_feature_bwd": (C.c_int, [C.POINTER(RbDinHistory), _p, _p, _p, _i64, _p, _p, _p, _p, _i32, _p]),
    "rb_dense_pack_input": (C.c_int, [_p, _i64, _i32, _i64, _p, _i32, _i32, _p]),
    "rb_dense_debug_stats": (C.c_int, [_p]),
    "rb_hash_ids": (C.c_int, [_p, _i32, _i64, _i64, _i32, _p, _p, _p, _p]),
    "rb_bucket_by_owner_workspace_bytes": (C.c_size_t, [_i64, _i32]),
    "rb_bucket_by_owner": (C.c_int, [_p, _i32, _i64, _i32, _p, _i64, _i32, _p, _p, _p, _p, _p, C.c_size_t, _p]),
    "rb_bucket_by_owner_skip": (C.c_int, [_p, _i32, _i64, _i32, _p, _i64, _i32, _i64, _p, _p, _p, _p, _p, C.c_size_t, _p]),
    "rb_criteo_index_workspace_bytes": (C.c_size_t, [_i64]),
    "rb_criteo_index_lines": (C.c_int, [_p, _i64, _i64, _p, _p, _p, C.c_size_t, _p]),
    "rb_criteo_parse": (C.c_int, [_p, _i64, _p, _i64, _p, _p, _p, _p, _p, _p, _i64, _p, _p]),
    "rb_vocab_build_workspace_bytes": (C.c_size_t, [_i64]),
    "rb_vocab_build": (C.c_int, [_p, _i64, _i32, _p, _i64, _p, _p, C.c_size_t, _p]),
    "rb_vocab_table_build": (C.c_int, [_p, _i64, _p, _p, _i64, _p]),
    "rb_vocab_lookup": (C.c_int, [_p, _i64, _p, _p, _i64, _p, _p]),
}


def header_symbols(path: str = HEADER):
    """Function names declared in the public header (used by the CPU-side export test)."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", text)))


class _Lib:
    def __init__(self):
        self._dll = None

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m recommender_b200.build` "
                "(there is no CPU fallback for the CUDA hot path)")
        dll = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(dll, name, None)
            if fn is None:
                raise RuntimeError(f"librecsys_b200.so does not export {name}; rebuild the library")
            fn.restype = res
            fn.argtypes = args
        self._dll = dll
        return dll

    def __getattr__(self, name):
        return getattr(self.load(), name)


lib = _Lib()


class RecsysError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib.rb_last_error()
        raise RecsysError(f"{what or 'recsys_b200 call'} failed with status {rc}: {msg.decode() if msg else ''}")
