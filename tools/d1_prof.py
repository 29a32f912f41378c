#!/usr/bin/env python
"""One MNFDense layer call (default: MNF-LeNet d1, 10240 x 800 -> 500, ReLU fused; `10240 500 10` = d2 + Softmax): device time of pack + GEMM; MNF_DENSE_PROF=1
prints where the CTA-pair GEMM's MMA warp waits.   python tools/d1_prof.py [rows] [K] [N]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tf_mnf_b200 as M  # noqa: E402
from tf_mnf_b200 import _lib as L  # noqa: E402
from tf_mnf_b200.layers import MNFDense  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 10240
KIN = int(sys.argv[2]) if len(sys.argv) > 2 else 800
NOUT = int(sys.argv[3]) if len(sys.argv) > 3 else 500
ACT = "relu" if NOUT > 16 else "softmax"
M.set_seed(0)
ctx = M.default_context()
lyr = MNFDense(NOUT, flow_type="rnvp", std_init=10.0)
x = ctx.to_device(np.random.RandomState(0).uniform(size=(B, KIN)).astype(np.float32))
lyr.build(x.shape)
ctx.freeze_weights(True)
for _ in range(3):
    lyr.call(x, activation=ACT)
ctx.sync()
ks, ke = [], []
for i in range(10):
    a, b = C.c_void_p(), C.c_void_p()
    L.check(ctx.lib.mnf_event_create(C.byref(a))), L.check(ctx.lib.mnf_event_create(C.byref(b)))
    L.check(ctx.lib.mnf_ctx_probe(ctx.handle, 2, 0, a, b))
    lyr.call(x, activation=ACT)
    ks.append(a), ke.append(b)
ctx.sync()
ms = C.c_float()
t = []
for a, b in zip(ks, ke):
    L.check(ctx.lib.mnf_event_elapsed_ms(a, b, C.byref(ms)))
    t.append(ms.value * 1e3)
print(f"{L.last_path()} B={B}: pack + GEMM {np.mean(t):.1f} us (min {np.min(t):.1f})")
