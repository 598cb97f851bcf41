"""The C-ABI shared library on a machine WITHOUT a GPU: it builds, loads, exports every symbol
include/recsys_b200.h declares, and its argument validation (which runs before any CUDA call)
returns the documented status codes.  No kernel is launched here."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from oracle import ctr_oracle as O
from recommender_b200 import _lib


def test_every_header_symbol_is_exported_and_bound(cuda_lib):
    declared = _lib.header_symbols()
    assert len(declared) >= 15
    dll = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(dll, name), f"{name} declared in recsys_b200.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table and header disagree"


def test_library_targets_sm_100a(cuda_lib):
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_struct_layout_matches_header():
    # plain C layout: 2*int32 + 16 ptr + 16 i64 + 16 i64 + 4 ptr ; group = ptr,2*i32,i64,ptr,i64 + grad
    assert C.sizeof(_lib.RbGradSource) == 8 + 16 * 8 * 3 + 4 * 8
    assert C.sizeof(_lib.RbLookupGroup) == 8 + 8 + 8 + 8 + 8 + C.sizeof(_lib.RbGradSource)
    assert C.sizeof(_lib.RbOptParams) == 32     # 2*int32 + 4*float + 1 pointer


def test_version_and_alpha_t(cuda_lib):
    assert cuda_lib.rb_version() == 100
    for step in (1, 2, 10, 1000, 100000):
        a = cuda_lib.rb_adam_alpha_t(1e-3, 0.9, 0.999, step)
        b = float(O.adam_alpha_t(step))
        assert np.float32(a) == np.float32(b), (step, a, b)   # both sides: libm powf in fp32, as TF-CPU


def test_argument_validation_without_a_gpu(cuda_lib):
    L = cuda_lib
    buf = (C.c_float * 64)()
    idx = (C.c_int64 * 4)()
    p = C.addressof(buf)
    # null table
    assert L.rb_gather_fwd(None, 10, 16, C.addressof(idx), _lib.RB_I64, 4, 1, None, 0, p, 16, None, None) == -1
    assert b"table" in L.rb_last_error()
    # unsupported D (D/vec > 32)
    assert L.rb_gather_fwd(p, 10, 132, C.addressof(idx), _lib.RB_I64, 4, 1, None, 0, p, 132, None, None) == -2
    # misaligned table for vec=4
    assert L.rb_gather_fwd(p + 4, 10, 16, C.addressof(idx), _lib.RB_I64, 4, 1, None, 0, p, 16, None, None) == -3
    # bad index type
    assert L.rb_gather_fwd(p, 10, 16, C.addressof(idx), 7, 4, 1, None, 0, p, 16, None, None) == -1
    # empty problems succeed without touching the device
    assert L.rb_gather_fwd(p, 10, 16, None, _lib.RB_I64, 0, 1, None, 0, None, 16, None, None) == 0
    assert L.rb_bag_pool_fwd(p, 10, 16, None, _lib.RB_I32, 0, 5, None, 0, _lib.RB_POOL_SUM, None, None, 16, None, None, None) == 0
    # interaction: too many features / bad D
    assert L.rb_dot_interaction_fwd(p, None, 0, None, 0, None, None, 1, 33, 16, 0, 1, 0, p, 0, 33 * 33, 0, None, None) == -2
    assert L.rb_dot_interaction_fwd(p, None, 0, None, 0, None, None, 1, 27, 24, 0, 1, 0, p, 0, 729, 0, None, None) == -2
    # tail without a dense vector
    assert L.rb_dot_interaction_fwd(p, None, 0, None, 0, None, None, 1, 26, 16, 0, 1, 1, p, 0, 800, 0, None, None) == -1
    # sparse update: workspace too small is reported with the needed size
    need = L.rb_sparse_bwd_update_workspace_bytes(1000, 16, 5000)
    assert need > 1000 * 16
    assert L.rb_sparse_bwd_update_workspace_bytes(-1, 16, 10) == 0
    gs = _lib.RbGradSource()
    gs.num_src = 1
    gs.src[0] = p
    opt = _lib.RbOptParams(_lib.RB_OPT_ADAM_LAZY, 1, 1e-3, 0.9, 0.999, 1e-7, None)
    rc = L.rb_sparse_bwd_update(p, p, p, 5000, 16, C.addressof(idx), _lib.RB_I64, 4, 1, None, 0, C.byref(gs), C.byref(opt),
                                p, 16, None, None)
    assert rc == -4 and b"workspace" in L.rb_last_error()
    # Adam without state
    rc = L.rb_sparse_bwd_update(p, None, None, 5000, 16, C.addressof(idx), _lib.RB_I64, 4, 1, None, 0, C.byref(gs),
                                C.byref(opt), p, need, None, None)
    assert rc == -1
    assert L.rb_bucket_by_owner_workspace_bytes(1000, 8) > 0
    assert L.rb_bucket_by_owner(C.addressof(idx), _lib.RB_I64, 4, 1, None, 0, 0, p, p, p, p, p, 1 << 20, None) == -1


def test_argument_validation_of_the_dense_and_peer_memory_entry_points(cuda_lib):
    """Status codes of the later additions, all decided before any CUDA call."""
    L = cuda_lib
    buf = (C.c_float * 64)()
    p = C.addressof(buf)
    opt = _lib.RbOptParams(_lib.RB_OPT_ADAM_LAZY, 1, 1e-3, 0.9, 0.999, 1e-7, None)
    # dense step: too many tensors / Adam without state
    slots = (_lib.RbDenseSlot * 1)()
    slots[0].param, slots[0].grad, slots[0].n = p, p, 8
    assert L.rb_dense_opt_step(slots, _lib.RB_MAX_DENSE_TENSORS + 1, C.byref(opt), None) == -1
    assert L.rb_dense_opt_step(slots, 1, C.byref(opt), None) == -1 and b"Adam" in L.rb_last_error()
    assert L.rb_dense_opt_step(slots, 0, C.byref(opt), None) == 0
    # column sum: cols must be a multiple of 8; workspace is checked
    assert L.rb_colsum(p, _lib.RB_F32, 4, 12, 12, p, p, 1 << 20, None) == -2
    assert L.rb_colsum_workspace_bytes(65536, 512) >= 512 * 4
    assert L.rb_colsum(p, _lib.RB_F32, 4, 16, 16, p, p, 8, None) == -4
    # loss head
    assert L.rb_bce_clipped(None, p, 0, 8, p, None, p, 1 << 20, None) == -1
    assert L.rb_bce_clipped(p, p, 7, 8, p, None, p, 1 << 20, None) == -1
    # interaction: RB_BF16_ONES needs a pad column
    assert L.rb_dot_interaction_fwd(p, None, 0, None, 0, None, None, 1, 27, 16, 0, 1, 0, p, _lib.RB_BF16_ONES, 729, 0, None, None) == -1
    assert L.rb_dot_interaction_fwd(p, None, 0, None, 0, None, None, 1, 27, 16, 0, 1, 0, p, 0, 729, 7, None, None) == -1      # bad row_cache
    # peer-memory path: world size and null checks
    ptrs = (C.c_void_p * 9)(*([p] * 9))
    nv = (C.c_int32 * 1)()
    assert L.rb_p2p_collect_keys(9, 0, 100, ptrs, ptrs, ptrs, 50, 200, p, 1 << 30, 16, nv, None, None) == -1
    assert L.rb_p2p_collect_keys(2, 5, 100, ptrs, ptrs, ptrs, 50, 200, p, 1 << 30, 16, nv, None, None) == -1
    assert L.rb_p2p_collect_keys(2, 0, 100, ptrs, ptrs, ptrs, 50, 200, p, 16, 16, nv, None, None) == -4
    sel = C.c_int32(0)
    assert L.rb_sparse_bwd_prepare_collected(50, 16, 200, p, 16, C.byref(sel), None) == -4
    assert L.rb_sparse_bwd_apply_p2p(p, p, p, 50, 16, 9, 100, 4, ptrs, 200, nv, C.byref(opt), p, 1 << 30, 0, None, None) == -1
    assert L.rb_sparse_bwd_apply_p2p(p, p, p, 50, 16, 2, 100, 3, ptrs, 200, nv, C.byref(opt), p, 1 << 30, 0, None, None) == -1   # n_local % L
    assert L.rb_dot_interaction_fwd_sharded(None, 2, 100, p, 1, None, None, 1, 26, 16, 0, 1, 0, p, 0, 676, None, None, None) == -1
    assert L.rb_ipc_open(None, None) == -1 and L.rb_shared_alloc(0, None, None) == -1


def test_argument_validation_of_the_id_check_and_the_launch_mode_switch(cuda_lib):
    L = cuda_lib
    idx = (C.c_int64 * 8)()
    rows = (C.c_int64 * 4)(3, 3, 3, 3)
    flag = (C.c_int32 * 1)()
    i, r, f = C.addressof(idx), C.addressof(rows), C.addressof(flag)
    assert L.rb_check_indices(i, _lib.RB_I64, 8, 0, r, 0, f, None) == -1          # L must be positive
    assert L.rb_check_indices(i, 7, 8, 4, r, 0, f, None) == -1                    # index type
    assert L.rb_check_indices(None, _lib.RB_I64, 8, 4, r, 0, f, None) == -1 and b"null" in L.rb_last_error()
    assert L.rb_check_indices(i, _lib.RB_I64, 8, 4, r, 0, None, None) == -1       # the flag is the result: required
    assert L.rb_check_indices(i, _lib.RB_I64, 7, 4, r, 0, f, None) == -2          # n is a whole number of samples
    assert L.rb_check_indices(i, _lib.RB_I64, 0, 4, r, 0, f, None) == 0           # empty batch: nothing launched
    for mode in (0, 1, -1):                                                       # host-side flag only; -1 = back to RB_PDL
        L.rb_set_pdl(mode)


def test_ops_refuse_cpu_tensors():
    import torch
    from recommender_b200 import ops
    W = torch.zeros(8, 16)
    idx = torch.zeros(4, dtype=torch.int64)
    with pytest.raises(_lib.RecsysError):
        ops.gather_fwd(W, idx)
    with pytest.raises(_lib.RecsysError):
        ops.dot_interaction_fwd(E=torch.zeros(2, 27, 16))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    fresh = _lib._Lib()
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fresh.load()


def test_argument_validation_of_the_input_side_entry_points(cuda_lib):
    """rb_criteo_* / rb_vocab_*: bad arguments come back as status codes before anything is launched."""
    L = cuda_lib
    buf = (C.c_uint8 * 4096)()
    p = (C.addressof(buf) + 255) & ~255                      # 256-byte aligned inside the buffer
    out = (C.c_int64 * 64)()
    q = C.addressof(out)
    # line index
    assert L.rb_criteo_index_workspace_bytes(1 << 20) > 0
    assert L.rb_criteo_index_workspace_bytes(1 << 31) == 0                       # chunks stay below 2 GiB
    assert L.rb_criteo_index_lines(p, 1 << 31, 8, q, q, p, 1 << 20, None) == -1
    assert L.rb_criteo_index_lines(p, 100, 8, q, None, p, 1 << 20, None) == -1   # num_lines_dev is required
    assert L.rb_criteo_index_lines(p + 1, 100, 8, q, q, p, 1 << 20, None) == -3  # text must be 16-byte aligned
    assert L.rb_criteo_index_lines(p, 100, 8, q, q, p, 16, None) == -4 and b"workspace" in L.rb_last_error()
    assert L.rb_criteo_index_lines(p, 100, 8, q, q, p + 8, 1 << 20, None) == -3  # workspace alignment
    # parse
    assert L.rb_criteo_parse(p, 100, q, 0, q, q, q, None, None, None, 0, None, None) == 0          # no lines: nothing to do
    assert L.rb_criteo_parse(p, 100, q, 2, q, q, None, None, None, None, 0, None, None) == -1      # neither keys nor ids asked for
    assert L.rb_criteo_parse(p, 100, q, 2, q, q, None, q, None, None, 0, None, None) == -1         # ids need the table
    assert L.rb_criteo_parse(p, 100, q, 2, q, q, None, q, q, q, 1000, None, None) == -1            # capacity not a power of two
    assert L.rb_criteo_parse(p + 4, 100, q, 2, q, q, q, None, None, None, 0, None, None) == -3
    assert L.rb_criteo_parse(None, 100, q, 2, q, q, q, None, None, None, 0, None, None) == -1
    # dictionary
    assert L.rb_vocab_build_workspace_bytes(1000) > 1000 * 8 * 2
    assert L.rb_vocab_build_workspace_bytes(1 << 31) == 0
    assert L.rb_vocab_build(q, 1000, 10, q, 10, q, p, 64, None) == -4
    assert L.rb_vocab_build(q, 1000, -1, q, 10, q, p, 1 << 30, None) == -1
    assert L.rb_vocab_build(None, 1000, 10, q, 10, q, p, 1 << 30, None) == -1
    assert L.rb_vocab_table_build(q, 10, q, q, 10, None) == -1                    # capacity must be a power of two ...
    assert L.rb_vocab_table_build(q, 16, q, q, 16, None) == -1                    # ... and leave an empty slot
    assert L.rb_vocab_lookup(q, 0, q, q, 16, q, None) == 0
    assert L.rb_vocab_lookup(q, 4, q, q, 12, q, None) == -1
