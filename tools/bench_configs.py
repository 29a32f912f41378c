#!/usr/bin/env python
"""Secondary BASELINE.json configs, timed like bench.py (CUDA events on the launching stream,
L2 flushed between timed calls, warm-up first).  bench.py's JSON line stays the headline (C3);
these lines document the other configs of SURVEY.md section 8(d).

  python tools/bench_configs.py --config c4 [--math bf16|tf32x3] [--steps K]

c4: one MNFDense(4096) on x ~ N(0,1) [8192, 4096], planar q/r flows x2, bf16 operands / fp32
    accumulate (configs[3]).  A step = layer(x) (sample_z through the planar chain + the paired
    contraction with in-kernel Philox eps).  Roofline: tensor (2 products x 2*M*K*N FLOP against the
    measured dense-bf16 peak of MEASURED_PEAKS.json)."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        for k in ("bf16_tflops", "bf16_dense_tflops", "tensor_bf16_tflops"):
            if k in p:
                return float(p[k]), "measured (MEASURED_PEAKS.json)"
    return 1662.7, "SURVEY.md 8(d) (measured cuBLAS bf16 burst)"


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4", choices=["c4"])
    ap.add_argument("--math", default="bf16", choices=["bf16", "tf32x3", "fp32"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--width", type=int, default=4096)
    args = ap.parse_args()

    import tf_mnf_b200 as M
    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200.layers import MNFDense

    M.set_seed(0)
    ctx = M.default_context(0)
    ctx.set_math(args.math)
    lib = ctx.lib
    B, K, N = args.rows, args.width, args.width
    rng = np.random.RandomState(0)
    x = ctx.to_device(rng.standard_normal((B, K)).astype(np.float32))
    lyr = MNFDense(N, flow_type="planar", std_init=10.0)
    lyr.build((B, K))
    ctx.freeze_weights(True)
    flush = ctx.empty((64 << 20,))

    def ev():
        p = C.c_void_p()
        L.check(lib.mnf_event_create(C.byref(p)))
        return p

    for _ in range(max(args.warmup, 3)):
        lyr(x)
    ctx.sync()
    path = L.last_path()
    n0 = ctx.launches
    lyr(x)
    per_step = ctx.launches - n0
    starts, stops, kst, ksp = [], [], [], []
    for i in range(args.steps):
        L.check(lib.mnf_memset(ctx.handle, C.c_void_p(flush.ptr), 0, flush.nbytes))
        s, e, ks, ke = ev(), ev(), ev(), ev()
        L.check(lib.mnf_ctx_probe(ctx.handle, 2, 0, ks, ke))  # MNF_PROBE_DENSE_GEMM: the contraction kernel alone
        L.check(lib.mnf_event_record(ctx.handle, s))
        lyr(x)
        L.check(lib.mnf_event_record(ctx.handle, e))
        starts.append(s), stops.append(e), kst.append(ks), ksp.append(ke)
    ctx.sync()
    ms = C.c_float()

    def el(a, b):
        L.check(lib.mnf_event_elapsed_ms(a, b, C.byref(ms)))
        return ms.value

    step_ms = float(np.mean([el(a, b) for a, b in zip(starts, stops)]))
    kern_ms = float(np.mean([el(a, b) for a, b in zip(kst, ksp)]))
    flop = 2 * 2.0 * B * K * N
    peak, src = tensor_peak()
    ach = flop / (kern_ms * 1e-3) / 1e12
    print(json.dumps({
        "metric": "MNFDense rows/sec fwd", "value": B / (step_ms * 1e-3), "unit": "rows/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
        "dtype": {"bf16": "bf16 operands, f32 accumulate", "tf32x3": "f32 (3xTF32 split)", "fp32": "f32"}[args.math],
        "data": "synthetic",
        "config": {"workload": f"wide MNFDense {K}->{N}, planar flows x2, batch {B} (BASELINE.json configs[3])",
                   "math": args.math, "kernel_path": path, "weights": "frozen (packed once)",
                   "l2": "flushed between timed calls"},
        "gpu_launches": per_step * args.steps,
        "roofline": {"bound": "tensor", "kernel": path, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                     "frac": ach / peak, "traffic": None, "peak_source": src, "kernel_ms": kern_ms,
                     "algorithmic_flop_per_launch": flop},
    }))


if __name__ == "__main__":
    main()
