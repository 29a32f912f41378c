"""Diagnostic (run by hand on a GPU box: python tests/exp_accumulation_drift.py): error of the flow chains and
of the dense contraction against the float64 oracle at large D / K, FFMA vs tcgen05 -- the numbers quoted in
DESIGN.md section 4 and csrc/common.cuh.  Lives in tests/ because it uses the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tf_mnf_b200 as M
from oracle import mnf_oracle as O
from gpu_util import make_flows
from tf_mnf_b200.flows import NormalizingFlow
from tf_mnf_b200 import _lib as L
ctx = M.default_context(0)
f32 = lambda a: np.asarray(a, np.float32)
for kind in ("iaf", "rnvp"):
    for D, B in ((800, 130), (2048, 40), (4096, 33)):
        rng = np.random.RandomState(D + B)
        steps = O.make_flow(rng, kind, D, 3, (50,), "ones", np.float64)
        for st in steps:
            for k, v in st.items():
                if k.startswith("b"):
                    for b in v if isinstance(v, list) else [v]:
                        b[...] = 0.1 * rng.standard_normal(b.shape)
            if kind == "iaf":
                st["W"][1] *= 0.25
        z = rng.standard_normal((B, D))
        bern = [(rng.uniform(size=(B, D)) <= 0.5).astype(np.float64) for _ in range(3)]
        ref_states, ref_ld, caches = O.flow_forward(z, steps, bern)
        # the same chain evaluated by NumPy in float32 (what any fp32 implementation, TF's included, is up against)
        s32 = O.cast_params(steps, np.float32)
        st32, ld32, _ = O.flow_forward(f32(z), s32, [f32(b) for b in bern])
        for mode in ("fp32", "tf32x3"):
            ctx.set_math(mode)
            nf = NormalizingFlow(make_flows(steps, D))
            states, ld = nf.forward(f32(z), bern=[f32(b) for b in bern] if kind == "rnvp" else None, record=True)
            e_s = max(np.max(np.abs(a.numpy() - b)) / np.max(np.abs(b)) for a, b in zip(states, ref_states))
            e_l = np.max(np.abs(ld.numpy() - ref_ld)) / np.max(np.abs(ref_ld))
            print(kind, D, mode, L.last_path(), f"states {e_s:.2e} ld {e_l:.2e}", end=" | ")
        e_s = max(np.max(np.abs(a - b)) / np.max(np.abs(b)) for a, b in zip(st32, ref_states))
        e_l = np.max(np.abs(ld32 - ref_ld)) / np.max(np.abs(ref_ld))
        print(f"numpy-f32 states {e_s:.2e} ld {e_l:.2e}")
ctx.set_math("fp32")
# dense forward at long K: 3xTF32 (one TMEM accumulation chain of K terms) vs FFMA vs the float64 oracle
from gpu_util import make_layer
for K, N, B in ((800, 500, 256), (4096, 512, 256)):
    rng = np.random.RandomState(K)
    P = O.make_layer_params(rng, (K, N), flow="iaf", std_init=10.0, dtype=np.float64)
    x = rng.standard_normal((B, K)); eps_z, eps = rng.standard_normal((B, K)), rng.standard_normal((B, N))
    ref, _ = O.dense_call(P, x, eps_z, eps, 1.0)
    for mode in ("fp32", "tf32x3"):
        ctx.set_math(mode)
        lyr = make_layer(P, x.shape)
        y = lyr(f32(x), noise=dict(eps_out=f32(eps), eps_z=f32(eps_z))).numpy()
        print("dense", K, N, mode, L.last_path(), f"{np.max(np.abs(y - ref)) / np.max(np.abs(ref)):.2e}", end=" | ")
    print()
ctx.set_math("fp32")
