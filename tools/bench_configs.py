#!/usr/bin/env python
"""Secondary BASELINE.json configs, timed like bench.py (CUDA events on the launching stream,
L2 flushed between timed calls, warm-up first).  bench.py's JSON line stays the headline (C3);
these lines document the other configs of SURVEY.md section 8(d).

  python tools/bench_configs.py --config c4 [--math bf16|tf32x3] [--steps K]

c1: MNFLeNet train step, batch 128 (configs[0]): forward + softmax cross-entropy + backward +
    kl_div + KL backward + Adam; images/s.
c2: MNFFeedForward (100, 50, 10) on 136 features, IAF/MADE flows, batch 1024 (configs[1]):
    forward + MSE + backward + KL + Adam; rows/s.
c5: MNFVGG on 32x32x3, batch 256 split over the ranks, data-parallel training with one NCCL
    all-reduce over the flat fp32 gradient buffer (configs[4]); run under torchrun for N > 1;
    images/s (whole job), max over ranks.
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


def train_config(args) -> None:
    import torch

    import tf_mnf_b200 as M
    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200.models import MNFFeedForward, MNFLeNet, MNFVGG
    from tf_mnf_b200.train import Trainer

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    M.runtime.set_default_device(local)
    M.set_seed(0)
    ctx = M.default_context(local)
    ctx.set_math(args.math or "tf32x3")
    lib = ctx.lib
    rng = np.random.RandomState(0)
    if args.config == "c1":
        GB, unit = 128, "images/s"
        model = MNFLeNet(max_std=1, flow_h_sizes=[50], std_init=10.0, flow_type="rnvp")
        x = rng.uniform(size=(GB, 28, 28, 1)).astype(np.float32)
        y = np.eye(10, dtype=np.float32)[rng.randint(0, 10, size=GB)]
        loss, klw, name = "xent", 1.0 / 60000, "MNF-LeNet5 (RNVP flows) train step, synthetic 28x28x1 batch 128 (configs[0])"
    elif args.config == "c2":
        GB, unit = 1024, "rows/s"
        model = MNFFeedForward(layer_sizes=(100, 50, 10), std_init=1e-2)
        x = rng.standard_normal((GB, 136)).astype(np.float32)
        y = rng.standard_normal((GB, 10)).astype(np.float32)
        loss, klw, name = "mse", 1.0 / 128, "MNF-MLP regressor 136->100->50->10 (IAF/MADE flows) fwd+bwd+KL, batch 1024 (configs[1])"
    else:
        GB, unit = 256, "images/s"
        model = MNFVGG()
        x = rng.uniform(size=(GB, 32, 32, 3)).astype(np.float32)
        y = np.eye(10, dtype=np.float32)[rng.randint(0, 10, size=GB)]
        loss, klw, name = "xent", 1.0 / 50000, "MNF VGG-style conv net, synthetic 32x32x3 global batch 256, data-parallel (configs[4])"
    lo, hi = rank * GB // world, (rank + 1) * GB // world
    model.set_row_offset(lo)
    xd, yd = ctx.to_device(x[lo:hi]), ctx.to_device(y[lo:hi])
    tr = Trainer(model, loss=loss, kl_weight=klw, world=world)

    def ev():
        p = C.c_void_p()
        L.check(lib.mnf_event_create(C.byref(p)))
        return p

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        tr.step(xd, yd)
    barrier()
    cap = None
    if args.graph and world == 1:
        # the whole step replayed as ONE CUDA graph (Trainer.capture): a LeNet / MLP step is ~170 small launches and
        # the host's enqueue through ctypes costs more than the GPU work
        cap = tr.capture(xd, yd)
        step = lambda: cap.step()
    else:
        step = lambda: tr.step(xd, yd)
    for _ in range(3):
        step()
    barrier()
    def all_launches():  # the main context, the KL branch's forked context and the model's side contexts
        return ctx.launches + (tr.kl_ctx.launches if tr.kl_ctx else 0) + sum(c.launches for c in getattr(model, "_sides", []))

    n0 = all_launches()
    tr.step(xd, yd)   # (eager: counts the kernels of one step; a replay launches the same ones)
    per_step = all_launches() - n0
    barrier()
    import time
    starts, stops = [ev() for _ in range(args.steps)], [ev() for _ in range(args.steps)]
    t0 = time.perf_counter()
    for i in range(args.steps):
        L.check(lib.mnf_event_record(ctx.handle, starts[i]))
        step()
        L.check(lib.mnf_event_record(ctx.handle, stops[i]))
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    ms = C.c_float()
    per = []
    for a, b in zip(starts, stops):
        L.check(lib.mnf_event_elapsed_ms(a, b, C.byref(ms)))
        per.append(ms.value)
    # a step contains host synchronisation (the all-reduce hands the gradient buffer to torch), so the
    # wall clock per step (barrier to barrier) is the honest figure; the event time is reported beside it
    mine = torch.tensor([wall_ms, float(np.mean(per))], device="cuda")
    if dist is not None:
        dist.all_reduce(mine, op=dist.ReduceOp.MAX)
    wall_ms, dev_ms = (float(v) for v in mine.cpu())
    n_params = sum(p.size for _, p, _ in tr._named_params())
    cpu = None
    if rank == 0 and args.config == "c1" and not args.no_cpu:
        import bench  # the CPU leg lives in bench.py (the one place besides tests/ that runs oracle/)

        t_cpu, _ = bench.time_cpu_train("rnvp", GB, steps=5, warmup=1)
        cores = bench.blas_threads()
        cpu = {"value": GB / t_cpu, "unit": unit, "cores": cores, "kind": "port", "ms_per_step": t_cpu * 1e3,
               "sample": f"5 full steps (batch {GB}): forward + cross-entropy + backward + KL + KL backward, NumPy float32 "
                         f"restatement of the reference TF path (oracle/), {cores} BLAS threads; Adam not included"}
    if rank == 0:
        print(json.dumps({
            "cpu_baseline": cpu,
            "metric": "train-step throughput (fwd + loss + bwd + KL + KL bwd + allreduce + Adam)", "value": GB / (wall_ms * 1e-3),
            "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": wall_ms,
            "device_ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "global_batch": GB, "math": ctx.math, "n_params": int(n_params),
                       "allreduce_bytes": int(n_params) * 4 if world > 1 else 0,
                       "launch": "one CUDA-graph replay per step (Trainer.capture)" if cap else "eager enqueue",
                       "timing": "wall clock barrier to barrier, max over ranks (steps contain host syncs)"},
            "gpu_launches": per_step * args.steps}))
    if dist is not None:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4", choices=["c1", "c2", "c4", "c5"])
    ap.add_argument("--math", default=None, choices=["bf16", "tf32x3", "fp32"],
                    help="default: bf16 for c4, tf32x3 (fp32 parity) for the training configs")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="c1: skip the CPU (NumPy restatement) train-step baseline")
    ap.add_argument("--graph", action="store_true", help="c1 / c2 / c5 on one GPU: replay the train step as one CUDA graph")
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--width", type=int, default=4096)
    args = ap.parse_args()
    if args.config != "c4":
        train_config(args)
        return
    args.math = args.math or "bf16"

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
