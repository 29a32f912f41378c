"""CPU-only checks: the C-ABI library builds, loads and exports every symbol the header
declares; the host logic (initialisers, MADE masks, shape rules, noise plans, DLPack view,
error behaviour) works; and the product fails LOUDLY without a GPU (no CPU fallback)."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import pytest
from conftest import ROOT

from oracle import mnf_oracle as O


@pytest.fixture(scope="module")
def lib():
    from tf_mnf_b200 import _build, _lib

    _build.build()
    return _lib.load()


def test_library_exports_every_header_symbol(lib):
    from tf_mnf_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "mnf_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(mnf_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 50
    for name in declared:
        assert hasattr(lib, name), f"libmnf_b200.so does not export {name}"
    assert declared == set(_lib.ALL_SYMBOLS), declared ^ set(_lib.ALL_SYMBOLS)
    assert lib.mnf_version() == 100


def test_struct_layouts_match_header():
    """ctypes struct sizes == what nvcc/gcc lay out for include/mnf_b200.h (compiled here)."""
    import subprocess
    import tempfile

    from tf_mnf_b200 import _lib as L

    src = r'''
#include <stdio.h>
#include "mnf_b200.h"
int main(){printf("%zu %zu %zu %zu %zu %zu\n", sizeof(mnf_noise), sizeof(mnf_flow_step), sizeof(mnf_conv_shape),
  sizeof(mnf_kl_args), sizeof(mnf_kl_grads), sizeof(mnf_tensor_view)); return 0;}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        sizes = [int(v) for v in subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(s) for s in (L.Noise, L.FlowStep, L.ConvShape, L.KlArgs, L.KlGrads, L.TensorView)]


def test_no_gpu_fails_loudly(lib):
    """Without a CUDA device a context cannot be created: nothing silently runs on the CPU."""
    n = C.c_int(0)
    rc = lib.mnf_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    import tf_mnf_b200 as M
    from tf_mnf_b200._lib import MnfError
    from tf_mnf_b200.layers import MNFDense

    with pytest.raises(MnfError, match="no CUDA device|no CPU fallback"):
        M.Context(0)
    with pytest.raises(MnfError):
        MNFDense(4)(np.zeros((2, 3), np.float32))


def test_argument_errors_do_not_need_a_gpu(lib):
    from tf_mnf_b200 import _lib as L

    with pytest.raises(ValueError, match="NULL"):
        L.check(lib.mnf_dense_forward(None, 1, 1, 1, None, None, None, None, None, None, 1.0, None, None, None))
    s = L.ConvShape(B=1, H=3, W=3, C=1, kh=5, kw=5, F=2, stride=1, same=0)
    ho, wo = C.c_int64(), C.c_int64()
    with pytest.raises(ValueError, match="smaller than kernel"):
        L.check(lib.mnf_conv_out_hw(C.byref(s), C.byref(ho), C.byref(wo)))
    s.same = 1
    L.check(lib.mnf_conv_out_hw(C.byref(s), C.byref(ho), C.byref(wo)))
    assert (ho.value, wo.value) == (3, 3)


def test_dlpack_view_rejects_cpu_tensor(lib):
    torch = pytest.importorskip("torch")
    from tf_mnf_b200 import runtime

    cap = torch.arange(6, dtype=torch.float32).reshape(2, 3).__dlpack__()
    with pytest.raises(ValueError, match="not CUDA|no CPU path"):
        runtime.dlpack_view(cap)


def test_dlpack_wrap_roundtrip_struct(lib):
    """mnf_dlpack_wrap builds a valid DLManagedTensor (checked through mnf_dlpack_view)."""
    from tf_mnf_b200 import _lib as L

    released = []
    cb = L.RELEASE_FN(lambda arg: released.append(arg))
    shape = (C.c_int64 * 3)(2, 3, 4)
    out = C.c_void_p()
    L.check(lib.mnf_dlpack_wrap(C.c_void_p(0x1000), 3, shape, 0, cb, C.c_void_p(7), C.byref(out)))
    tv = L.TensorView()
    L.check(lib.mnf_dlpack_view(out, C.byref(tv)))
    assert (tv.ndim, list(tv.shape[:3]), tv.device_type, tv.dtype_code, tv.dtype_bits, tv.contiguous) == (3, [2, 3, 4], 2, 2, 32, 1)
    assert tv.data == 0x1000
    L.check(lib.mnf_dlpack_delete(out))
    assert released == [7]


def test_glorot_initialisers():
    from tf_mnf_b200 import init

    rng = np.random.RandomState(0)
    w = init.glorot_normal(rng, (5, 5, 20, 50))
    fi, fo = 5 * 5 * 20, 5 * 5 * 50
    sigma = np.sqrt(2.0 / (fi + fo))
    assert w.dtype == np.float32 and abs(w.std() / sigma - 1) < 0.02
    assert np.abs(w).max() <= 2 * sigma / 0.87962566103423978 + 1e-7
    assert init.fans((800,)) == (800, 800) and init.fans((800, 500)) == (800, 500)
    u = init.glorot_uniform(rng, (800, 50))
    assert np.abs(u).max() <= np.sqrt(6 / 850) and abs(u.std() / np.sqrt(2 / 850) - 1) < 0.02
    # same draw as the oracle's initialiser for the same RandomState
    a = init.glorot_normal(np.random.RandomState(3), (7, 9))
    b = O.glorot_normal(np.random.RandomState(3), (7, 9), np.float32)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("D,h", [(20, (50,)), (800, (50,)), (12, (6, 5)), (10, (50,))])
def test_host_made_masks_match_oracle(D, h):
    from tf_mnf_b200.flows.made import build_masks

    masks, deg = build_masks(D, h, 2)
    ref = O.made_masks(D, h, 2, "autoregressive")
    for a, b in zip(masks, ref):
        np.testing.assert_array_equal(a, b)
    assert masks[-1].shape == (h[-1], 2 * D)


def test_constructor_contract():
    from tf_mnf_b200.flows import IAF, RNVP
    from tf_mnf_b200.layers import MNFConv2D, MNFDense
    from tf_mnf_b200.models import MNFFeedForward, MNFLeNet

    d = MNFDense(7)
    assert (d.n_out, d.n_flows_q, d.n_flows_r, d.learn_p, d.use_z, d.flow_h_sizes, d.max_std, d.std_init) == (7, 2, 2, False, True, (50,), 1.0, 1.0)
    c = MNFConv2D(20, 5, padding="VALID")
    assert c.kernel_size == [5, 5] and c.strides == 1 and c.padding == "VALID"
    assert MNFConv2D(4, (3, 2)).kernel_size == [3, 2]
    with pytest.raises(TypeError, match="unexpected keyword"):
        MNFDense(3, bogus=1)
    with pytest.raises(ValueError):
        MNFDense(3, flow_type="glow")
    assert IAF(0, h_sizes=(6, 5)).h_sizes == (6, 5) and RNVP(4, h_sizes=(7, 4), activation="relu").activation == "relu"
    with pytest.raises(NotImplementedError, match="hidden layers"):
        IAF(0, h_sizes=(6, 5, 4, 3, 2))
    with pytest.raises(NotImplementedError):
        RNVP(4, activation="swish")
    m = MNFLeNet(max_std=1, flow_h_sizes=[50], std_init=10.0)
    assert [type(l).__name__ for l in m.mnf_layers] == ["MNFConv2D", "MNFConv2D", "MNFDense", "MNFDense"]
    assert [l.n_filters for l in m.mnf_layers[:2]] == [20, 50] and [l.n_out for l in m.mnf_layers[2:]] == [500, 10]
    ff = MNFFeedForward()
    assert [l.n_out for l in ff.mnf_layers] == [100, 50, 10]


def test_conv_output_shape_rule():
    from tf_mnf_b200.layers.mnf_conv import conv_out_hw

    assert conv_out_hw(28, 28, 5, 5, 1, "VALID") == (24, 24)
    assert conv_out_hw(12, 12, 5, 5, 1, "VALID") == (8, 8)
    assert conv_out_hw(32, 32, 3, 3, 1, "SAME") == (32, 32)
    assert conv_out_hw(11, 10, 3, 2, 2, "SAME") == (6, 5)
    with pytest.raises(ValueError):
        conv_out_hw(3, 3, 5, 5, 1, "VALID")
    x = np.zeros((1, 11, 10, 2))
    assert O.conv2d(x, np.zeros((3, 2, 2, 4)), 2, "SAME").shape[1:3] == (6, 5)


def test_noise_plan_streams_are_distinct():
    from tf_mnf_b200 import noise

    ids = set()
    for layer in (1, 2, 300):
        for draw in [noise.CALL_Z, noise.CALL_OUT, noise.KL_Z, noise.KL_A, noise.KL_B,
                     *[noise.BERN_Q_CALL + k for k in range(4)], *[noise.BERN_Q_KL + k for k in range(4)],
                     *[noise.BERN_R_KL + k for k in range(4)], noise.FLOW_BERN]:
            ids.add(noise.stream_id(layer, draw))
    assert len(ids) == 3 * 18
    n = noise.make(seed=5, call_index=3, layer_id=2, draw=noise.CALL_OUT, row_offset=10240)
    assert n.seed == 5 | (3 << 32) and n.row_offset == 10240 and n.stream == (2 << 8) | 1 and not n.injected


def test_philox_reference_vector():
    """Known-answer test of the NumPy Philox used to check the device words: Random123's
    published vector philox4x32-10(counter=0, key=0) and the all-ones vector."""
    from gpu_util import philox4x32_10

    out = philox4x32_10(np.zeros((1, 4), np.uint32), (0, 0))[0]
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    out = philox4x32_10(np.full((1, 4), 0xFFFFFFFF, np.uint32), (0xFFFFFFFF, 0xFFFFFFFF))[0]
    assert [hex(v) for v in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_tf_glue_is_import_guarded():
    """The package never imports TensorFlow; tf_glue imports cleanly and only fails when used."""
    import importlib
    import sys

    sys.modules.pop("tensorflow", None)
    mod = importlib.import_module("tf_mnf_b200.tf_glue")
    with pytest.raises(ImportError, match="TensorFlow"):
        mod.make_keras_layers()


def test_nccl_binds_at_run_time():
    """mnf_comm_* resolve NCCL with dlopen (csrc/comm.cu): the version query needs no GPU."""
    import ctypes as C

    from tf_mnf_b200 import _lib as L

    v = C.c_int()
    L.check(L.load().mnf_comm_nccl_version(C.byref(v)))
    assert v.value >= 20000
