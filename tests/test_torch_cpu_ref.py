"""The torch-CPU restatement timed by `bench.py --impl reference` against the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import ctr_oracle as O
from oracle.torch_cpu_ref import DLRMRef


@pytest.mark.parametrize("adam", ["tf_dense", "lazy"])
def test_dlrm_ref_matches_numpy_oracle(adam):
    V, D, B = 500, 16, 64
    ref = DLRMRef([32, 16], [32, 1], D, V, seed=1, adam=adam)
    params = dict(table=ref.table.numpy().copy(), bottom=[(W.numpy().copy(), b.numpy().copy()) for W, b in ref.bottom],
                  top=[(W.numpy().copy(), b.numpy().copy()) for W, b in ref.top])
    st = dict(m=np.zeros_like(params["table"]), v=np.zeros_like(params["table"]))
    for step in (1, 2, 3):
        cat, dense_x, label = O.synth_batch(B, V, seed=step, dist="zipf")
        loss, prob = ref.train_step(torch.tensor(cat), torch.tensor(dense_x), torch.tensor(label))
        if step == 1:
            oprob, cache = O.dlrm_forward(params, cat, dense_x)
            oloss, dprob = O.bce_clipped(oprob, label)
            np.testing.assert_allclose(prob.numpy(), oprob, rtol=1e-5, atol=1e-6)
            assert abs(loss - float(oloss)) < 1e-5
            grads = O.dlrm_backward(params, cache, dprob)
            O.sparse_backward_update(params["table"], st, cat, grads["dE"], "adam_" + adam, 1)
            np.testing.assert_allclose(ref.m.numpy(), st["m"], rtol=1e-3, atol=1e-9)
            big = np.abs(st["m"]) > 1e-7
            np.testing.assert_allclose(ref.table.numpy()[big], params["table"][big], rtol=0, atol=2e-5)
    assert ref.step == 3


def test_multi_table_offsets():
    ref = DLRMRef([16], [8, 1], 16, 50, num_tables=26, seed=2, adam="lazy")
    assert ref.table.shape == (1300, 16)
    cat, dense_x, label = O.synth_batch(8, 50, seed=3)
    before = ref.table.clone()
    ref.train_step(torch.tensor(cat), torch.tensor(dense_x), torch.tensor(label))
    touched = np.unique(cat + np.arange(26)[None] * 50)
    changed = torch.nonzero((ref.table != before).any(1)).reshape(-1).numpy()
    assert set(changed) <= set(touched)
