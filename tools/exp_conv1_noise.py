"""Experiment (round 1): would drawing conv1's eps ahead of time on a side stream pay?  conv1 with in-kernel
Philox vs pre-generated (injected) noise, and the generator alone.  Measured on a B200: 0.377 ms vs
0.339 ms per layer call, generator 0.259 ms for the 118 M normals -- no: the generator costs more than it
saves.  (What did pay: sharing the contractions between the samples of an image, csrc/conv_mc.cu.)"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tf_mnf_b200 as M
from tf_mnf_b200 import _lib as L
from tf_mnf_b200.layers import MNFConv2D
M.set_seed(0)
ctx = M.default_context(0); ctx.set_math("tf32x3"); lib = ctx.lib
B = 10240
x = ctx.to_device(np.random.RandomState(0).uniform(size=(B, 28, 28, 1)).astype(np.float32))
lyr = MNFConv2D(20, 5, padding="VALID", flow_type="rnvp", std_init=10.0, flow_h_sizes=[50])
lyr.build(x.shape); ctx.freeze_weights(True)
buf = ctx.empty((B, 24, 24, 20))
def ev():
    p = C.c_void_p(); L.check(lib.mnf_event_create(C.byref(p))); return p
def timeit(fn, n=10):
    for _ in range(3): fn()
    ctx.sync()
    s, e = ev(), ev()
    L.check(lib.mnf_event_record(ctx.handle, s))
    for _ in range(n): fn()
    L.check(lib.mnf_event_record(ctx.handle, e)); ctx.sync()
    ms = C.c_float(); L.check(lib.mnf_event_elapsed_ms(s, e, C.byref(ms))); return ms.value / n
print("philox in kernel  :", timeit(lambda: lyr.call(x, fuse_relu_pool=True)), L.last_path())
print("injected eps_out  :", timeit(lambda: lyr.call(x, noise=dict(eps_out=buf), fuse_relu_pool=True)), L.last_path())
gen = lambda: L.check(lib.mnf_philox_normal(ctx.handle, 1, 2, 0, B, 576, 20, C.c_void_p(buf.ptr)))
print("generator alone   :", timeit(gen))
