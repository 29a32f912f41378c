#!/usr/bin/env python
"""MNF-LeNet conv1 -> ReLU/pool -> conv2 -> ReLU/pool under mc_predict at the bench's size: device time of the
pair with conv1's rows expanded by their own kernel (default) or generated inside conv2's kernel (--fuse).
   python tools/conv12_prof.py [--fuse]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tf_mnf_b200 as M  # noqa: E402
from tf_mnf_b200 import _lib as L  # noqa: E402
from tf_mnf_b200.layers import MNFConv2D  # noqa: E402

fuse = "--fuse" in sys.argv
n, imgs = 1024, 10
M.set_seed(0)
ctx = M.default_context()
c1 = MNFConv2D(20, 5, flow_type="rnvp", std_init=10.0)
c2 = MNFConv2D(50, 5, flow_type="rnvp", std_init=10.0)
x = ctx.to_device(np.random.RandomState(0).uniform(size=(imgs, 28, 28, 1)).astype(np.float32))
c1.build(x.shape), c2.build((imgs * n, 12, 12, 20))
ctx.freeze_weights(True)


def run():
    a = c1.call_mc(x, n, fuse_relu_pool=True, lazy=fuse)
    return c2.call(a, fuse_relu_pool=True)


for _ in range(3):
    run()
ctx.sync()
ev = []
for _ in range(2):
    p = C.c_void_p()
    L.check(ctx.lib.mnf_event_create(C.byref(p)))
    ev.append(p)
N = 10
L.check(ctx.lib.mnf_event_record(ctx.handle, ev[0]))
for _ in range(N):
    run()
L.check(ctx.lib.mnf_event_record(ctx.handle, ev[1]))
ctx.sync()
ms = C.c_float()
L.check(ctx.lib.mnf_event_elapsed_ms(ev[0], ev[1], C.byref(ms)))
print(f"{L.last_path()}: conv1 (maps + z chains + expansion) + conv2 on {imgs * n} rows: {ms.value * 1e3 / N:.1f} us per call")
