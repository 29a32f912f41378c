#!/usr/bin/env python
"""The z chain of one MNF-LeNet dense layer at the bench's size (sample_z(10240), D = 800 by default, two RNVP
steps): device time of each flow step.  MNF_NO_FLOW_F16=1 selects the 3xTF32 kernel (flows_umma.cu).
   python tools/flow_prof.py [rows] [D]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tf_mnf_b200 as M  # noqa: E402
from tf_mnf_b200 import _lib as L  # noqa: E402
from tf_mnf_b200.layers import MNFDense  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 10240
D = int(sys.argv[2]) if len(sys.argv) > 2 else 800
M.set_seed(0)
ctx = M.default_context()
lyr = MNFDense(500, flow_type="rnvp", std_init=10.0)
lyr.build((B, D))
ctx.freeze_weights(True)
for _ in range(3):
    lyr.sample_z(B)
ctx.sync()
for step in (0, 1):
    ks, ke = [], []
    for i in range(10):
        a, b = C.c_void_p(), C.c_void_p()
        L.check(ctx.lib.mnf_event_create(C.byref(a))), L.check(ctx.lib.mnf_event_create(C.byref(b)))
        L.check(ctx.lib.mnf_ctx_probe(ctx.handle, 3, step, a, b))
        lyr.sample_z(B)
        ks.append(a), ke.append(b)
    ctx.sync()
    ms = C.c_float()
    t = []
    for a, b in zip(ks, ke):
        L.check(ctx.lib.mnf_event_elapsed_ms(a, b, C.byref(ms)))
        t.append(ms.value * 1e3)
    print(f"{L.last_path()} B={B} D={D} step {step + 1}: {np.mean(t):.1f} us (min {np.min(t):.1f})")
a, b = C.c_void_p(), C.c_void_p()
L.check(ctx.lib.mnf_event_create(C.byref(a))), L.check(ctx.lib.mnf_event_create(C.byref(b)))
L.check(ctx.lib.mnf_event_record(ctx.handle, a))
for _ in range(20):
    lyr.sample_z(B)
L.check(ctx.lib.mnf_event_record(ctx.handle, b))
ctx.sync()
ms = C.c_float()
L.check(ctx.lib.mnf_event_elapsed_ms(a, b, C.byref(ms)))
print(f"whole chain (z0 + 2 steps): {ms.value * 1e3 / 20:.1f} us per sample_z({B})")
