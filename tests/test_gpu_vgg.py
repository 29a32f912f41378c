"""BASELINE.json configs[4]: the MNF VGG-style conv net -- forward, one
training step, and the data-parallel gradient rule (on NCCL when the box has >= 2 GPUs)."""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vgg_full_size_forward_and_kl():
    import tf_mnf_b200 as M
    from tf_mnf_b200.models import MNFVGG

    M.set_seed(0)
    ctx = M.default_context()
    ctx.set_math("tf32x3")
    try:
        model = MNFVGG()
        x = np.random.RandomState(0).uniform(size=(16, 32, 32, 3)).astype(np.float32)
        p = model(x).numpy()
        kl = float(model.kl_div().numpy()[0])
    finally:
        ctx.set_math("tf32x3")
    assert p.shape == (16, 10) and np.all(np.isfinite(p))
    np.testing.assert_allclose(p.sum(1), 1.0, rtol=1e-5)
    assert np.isfinite(kl) and kl > 0
    shapes = [l.mean_W.shape for l in model.mnf_layers]
    assert shapes == [(3, 3, 3, 64), (3, 3, 64, 64), (3, 3, 64, 128), (3, 3, 128, 128), (3, 3, 128, 256),
                      (3, 3, 256, 256), (4096, 512), (512, 10)]
    n_w = sum(int(np.prod(s)) for s in shapes)
    assert n_w == 3_246_784  # mean_W elements; with log_std_W, biases, q0/r0 and flows: 9.87 M


def test_vgg_tensorcore_matches_ffma():
    """Same Philox noise under both arithmetic modes: the tensor-core forward of the whole net
    agrees with the FFMA forward to fp32 accuracy."""
    import tf_mnf_b200 as M
    from tf_mnf_b200.models import MNFVGG

    ctx = M.default_context()
    x = np.random.RandomState(1).uniform(size=(8, 32, 32, 3)).astype(np.float32)
    out = {}
    for mode in ("fp32", "tf32x3"):
        M.set_seed(3)
        ctx.set_math(mode)
        try:
            model = MNFVGG(filters=(16, 32, 64), n_hidden=128)
            model.presample_first = False
            out[mode] = model(x).numpy()
        finally:
            ctx.set_math("tf32x3")
    np.testing.assert_allclose(out["tf32x3"], out["fp32"], rtol=2e-4, atol=1e-6)


def test_vgg_train_step_single_process():
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dist_train_check.py")], check=True, timeout=600,
                   env={**os.environ, "WORLD_SIZE": "1", "RANK": "0", "LOCAL_RANK": "0"})


def test_vgg_data_parallel_nccl_world2():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29511",
                        os.path.join(ROOT, "tests", "dist_train_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_TRAIN_OK world=2" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
