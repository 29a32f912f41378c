"""bench.py's CPU legs (the NumPy restatement of the reference timed on host cores) must run
without a GPU: the forward + KL step of the headline metric and the LeNet train step of
BASELINE configs[0]."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_cpu_forward_kl_step_runs():
    import bench

    step = bench.cpu_step_factory("rnvp", 20)
    probs, kl = step()
    assert probs.shape == (20, 10) and np.all(np.isfinite(probs))
    np.testing.assert_allclose(probs.sum(1), 1.0, rtol=1e-5)
    assert np.isfinite(kl)


def test_cpu_train_step_runs_and_is_stochastic():
    import bench

    step = bench.cpu_train_step_factory("rnvp", 8)
    loss0, kl0, grads = step()
    loss1, kl1, _ = step()
    assert np.isfinite(loss0) and np.isfinite(kl0) and np.isfinite(loss1) and loss0 != loss1
    g1, g2, g3, g4 = grads
    assert g1["mean_W"].shape == (5, 5, 1, 20) and g2["mean_W"].shape == (5, 5, 20, 50)
    assert g3["mean_W"].shape == (800, 500) and g4["mean_W"].shape == (500, 10)
    for g in grads:
        assert all(np.all(np.isfinite(v)) for k, v in g.items() if isinstance(v, np.ndarray))
