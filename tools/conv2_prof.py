#!/usr/bin/env python
"""One MNF-LeNet conv2 layer call at the bench's size (10240 rows, fused ReLU + pool): device time and, with
MNF_CONV_PROF=1, where the CTA-pair kernel's MMA warp waits.   python tools/conv2_prof.py [rows]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tf_mnf_b200 as M  # noqa: E402
from tf_mnf_b200 import _lib as L  # noqa: E402
from tf_mnf_b200.layers import MNFConv2D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 10240
M.set_seed(0)
ctx = M.default_context()
lyr = MNFConv2D(50, 5, flow_type="rnvp", std_init=10.0)
x = ctx.to_device(np.random.RandomState(0).uniform(size=(B, 12, 12, 20)).astype(np.float32))
lyr.build(x.shape)
ctx.freeze_weights(True)
z, _ = lyr.sample_z(B)
for _ in range(3):
    lyr.call(x, fuse_relu_pool=True)
ctx.sync()
ev = []
for _ in range(2):
    p = C.c_void_p()
    L.check(ctx.lib.mnf_event_create(C.byref(p)))
    ev.append(p)
n = 10
ks, ke = [], []
for i in range(n):
    a, b = C.c_void_p(), C.c_void_p()
    L.check(ctx.lib.mnf_event_create(C.byref(a))), L.check(ctx.lib.mnf_event_create(C.byref(b)))
    L.check(ctx.lib.mnf_ctx_probe(ctx.handle, 1, 0, a, b))
    lyr.call(x, fuse_relu_pool=True)
    ks.append(a), ke.append(b)
ctx.sync()
ms = C.c_float()
t = []
for a, b in zip(ks, ke):
    L.check(ctx.lib.mnf_event_elapsed_ms(a, b, C.byref(ms)))
    t.append(ms.value * 1e3)
print(f"{L.last_path()} B={B}: kernel {np.mean(t):.1f} us (min {np.min(t):.1f}) -> {2 * 2 * 64 * 500 * 50 * B / np.mean(t) / 1e6:.0f} algorithmic TFLOP/s")
