#!/usr/bin/env python
"""bench.py -- MNF-LeNet5 MC-predictive forward + KL throughput (BASELINE.json metric).

A step = one pass of the hot path over one batch: `probs = model(x); kl = model.kl_div()`
for x = 10 synthetic 28x28x1 images x 1024 posterior samples = 10240 rows per GPU
(BASELINE.json configs[2]; weak scaling: every rank owns 10240 rows of a 10240*N-row global
batch, Philox counters indexed by the GLOBAL row, no data-path collective).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                     # reference algorithm on host cores

The reference arm times the NumPy restatement of the reference's TF path (oracle/, float32,
all BLAS threads) because TensorFlow cannot be installed here; it is a CPU
baseline, not the target -- the roofline fraction is the figure of merit.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MNF-LeNet MC-sample images/sec fwd+KL"
UNIT = "images/s"
N_IMAGES, N_SAMPLES = 10, 1024
ROWS = N_IMAGES * N_SAMPLES
# algorithmic bytes of the dominant kernel (conv2 paired implicit GEMM), SURVEY.md 8(d):
# per row read x 12*12*20*4 + write y 8*8*50*4, plus mean_W + var_W once per launch
# The fused kernel reads x [12,12,20] and writes the ReLU'd, 2x2-pooled output [4,4,50] (the
# un-pooled [8,8,50] of 8(d) never reaches HBM); the launch also reads mean_W and log_std_W once.
CONV2_BYTES_PER_ROW = 12 * 12 * 20 * 4 + 4 * 4 * 50 * 4
CONV2_WEIGHT_BYTES = 2 * 5 * 5 * 20 * 50 * 4
CONV2_FLOP_PER_ROW = 2 * 2 * 64 * 500 * 50   # two products (mean, variance), SURVEY.md 8(d)
# per-row algorithmic work of the other stages (fused outputs where a stock layer is folded in)
STAGES = {
    # name: (bytes per row, weight bytes per launch, FLOP per row)
    "conv1 (mc_expand: mean*z + sd*eps, ReLU, 2x2 pool; 11520 normals per row)": (12 * 12 * 20 * 4 + 20 * 4, 2 * 10 * 24 * 24 * 20 * 4, 4 * 24 * 24 * 20),
    "conv2 (conv_raw_f16x2: CTA pairs, tcgen05 cta_group::2)": (CONV2_BYTES_PER_ROW, CONV2_WEIGHT_BYTES, CONV2_FLOP_PER_ROW),
    # mc_predict's default: conv2's plane producers generate conv1's noisy, pooled rows themselves (one kernel)
    "conv1+conv2 (conv_raw_f16x2<gen>: conv1's per-sample noise expansion inside conv2's CTA-pair tcgen05 kernel; 11520 + 3200 normals per row)":
        (20 * 4 + 4 * 4 * 50 * 4, 2 * 10 * 24 * 24 * 20 * 4 + CONV2_WEIGHT_BYTES, CONV2_FLOP_PER_ROW + 4 * 24 * 24 * 20),
    "d1 (row_pack + paired_f16x3, ReLU fused)": (2 * 800 * 4 + 500 * 4, 2 * 800 * 500 * 4, 2 * 2 * 800 * 500),
    "d2 (dense_small_rows, Softmax fused)": (2 * 500 * 4 + 10 * 4, 2 * 500 * 10 * 4, 2 * 2 * 500 * 10),
}
LENET_BYTES_PER_ROW = 80776 + 5111520 / 10240   # SURVEY.md 8(d), layer-API boundary, B = 10240
LENET_FLOP_PER_ROW = 9.994e6


def peaks():
    """(HBM GB/s, dense bf16 TFLOP/s burst, sustained, source)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return (float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1662.7)), float(p.get("bf16_tflops_sustained", 1402.8)),
                "measured (MEASURED_PEAKS.json)")
    return 6650.0, 1662.7, 1402.8, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int) -> None:
        self.device, self.rows, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self, t0: float = 0.0, t1: float = float("inf")) -> dict:
        """Only samples that arrived inside [t0, t1] (the timed region) are used."""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if not (t0 <= ts <= t1 + 0.03):
                continue
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_step_factory(flow: str, rows: int):
    """Reference algorithm on host cores: NumPy float32 restatement (oracle/), noise drawn
    inside the step like tf.random.normal does in the reference."""
    from oracle import mnf_oracle as O

    Ps = O.make_lenet_params(seed=0, flow=flow, std_init=10.0, dtype=np.float32)
    rng = np.random.RandomState(0)
    imgs = rng.uniform(size=(N_IMAGES, 28, 28, 1)).astype(np.float32)
    reps = -(-rows // N_IMAGES)
    x = np.repeat(imgs, reps, axis=0)[:rows]
    Ds = [20, 50, 800, 500]
    outs = [(rows, 24, 24, 20), (rows, 8, 8, 50), (rows, 500), (rows, 10)]
    kl_shapes = [(25,), (500,), (500,), (10,)]
    f32 = np.float32

    def step():
        noise, bern, kln, klb = [], [], [], []
        for D, o in zip(Ds, outs):
            noise.append((rng.standard_normal((rows, D)).astype(f32), rng.standard_normal(o).astype(f32)))
            bern.append([(rng.uniform(size=(rows, D)) <= 0.5).astype(f32) for _ in range(2)] if flow == "rnvp" else None)
        for i, (D, ks) in enumerate(zip(Ds, kl_shapes)):
            ez, ea = rng.standard_normal((1, D)).astype(f32), rng.standard_normal(ks).astype(f32)
            kln.append((ez, ea, f32(rng.standard_normal())) if i < 2 else (ez, ea))
            b = [(rng.uniform(size=(1, D)) <= 0.5).astype(f32) for _ in range(4)] if flow == "rnvp" else [None] * 4
            klb.append((b[:2], b[2:]) if flow == "rnvp" else (None, None))
        probs = O.lenet_forward(Ps, x, noise, max_std=1.0, bern=bern)
        kl = O.lenet_kl(Ps, kln, bern=klb)
        return probs, kl

    return step


def time_cpu(flow: str, rows: int, steps: int, warmup: int):
    step = cpu_step_factory(flow, rows)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return float(np.mean(ts)), ts



def cpu_train_step_factory(flow: str = "rnvp", batch: int = 128):
    """BASELINE configs[0] on host cores: one MNF-LeNet5 training step (forward, mean softmax
    cross-entropy + kl/60000, backward through every layer and both KL terms -- what
    tape.gradient computes in notebooks/lenet_mnist.py:108-114) with the NumPy float32
    restatement of the reference (oracle/), noise drawn inside the step.  The Adam update
    (1.7 M parameters, < 1 % of the step) is not included."""
    from oracle import mnf_oracle as O

    Ps = O.make_lenet_params(seed=0, flow=flow, std_init=10.0, dtype=np.float32)
    rng = np.random.RandomState(0)
    x = rng.uniform(size=(batch, 28, 28, 1)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.randint(0, 10, size=batch)]
    Ds = [20, 50, 800, 500]
    outs = [(batch, 24, 24, 20), (batch, 8, 8, 50), (batch, 500), (batch, 10)]
    kl_shapes = [(25,), (500,), (500,), (10,)]
    f32 = np.float32

    def pool_fwd(h):  # ReLU + MaxPool2D(2x2, SAME) on even sizes, keeping what the backward needs
        B, H, W, Cc = h.shape
        r = np.maximum(h, 0).reshape(B, H // 2, 2, W // 2, 2, Cc)
        p = r.max(axis=(2, 4))
        return p, (r, p, h.shape)

    def pool_bwd(g, c):
        r, p, shp = c
        mask = (r == p[:, :, None, :, None, :]) & (r > 0)
        mask = mask / np.maximum(mask.sum(axis=(2, 4), keepdims=True), 1)  # ties share the gradient
        return (mask * g[:, :, None, :, None, :]).reshape(shp).astype(f32)

    def step():
        noise, bern = [], []
        for D, o in zip(Ds, outs):
            noise.append((rng.standard_normal((batch, D)).astype(f32), rng.standard_normal(o).astype(f32)))
            bern.append([(rng.uniform(size=(batch, D)) <= 0.5).astype(f32) for _ in range(2)] if flow == "rnvp" else None)
        h, c1 = O.conv_call(Ps[0], x, *noise[0], max_std=1.0, bern_q=bern[0])
        h, p1 = pool_fwd(h)
        h, c2 = O.conv_call(Ps[1], h, *noise[1], max_std=1.0, bern_q=bern[1])
        h, p2 = pool_fwd(h)
        flat_shape = h.shape
        h, c3 = O.dense_call(Ps[2], h.reshape(batch, -1), *noise[2], max_std=1.0, bern_q=bern[2])
        a3 = h
        h, c4 = O.dense_call(Ps[3], np.maximum(h, 0), *noise[3], max_std=1.0, bern_q=bern[3])
        probs = O.softmax(h)
        loss = float(-np.mean(np.log(np.maximum(probs[y > 0], 1e-30))))
        g = ((probs - y) / batch).astype(f32)
        g, g4 = O.dense_call_backward(Ps[3], c4, g)
        g, g3 = O.dense_call_backward(Ps[2], c3, (g * (a3 > 0)).astype(f32))
        g, g2 = O.conv_call_backward(Ps[1], c2, pool_bwd(g.reshape(flat_shape), p2))
        g, g1 = O.conv_call_backward(Ps[0], c1, pool_bwd(g, p1))
        kl = 0.0
        for i, (D, ks) in enumerate(zip(Ds, kl_shapes)):
            ez, ea = rng.standard_normal((1, D)).astype(f32), rng.standard_normal(ks).astype(f32)
            b = [(rng.uniform(size=(1, D)) <= 0.5).astype(f32) for _ in range(4)] if flow == "rnvp" else [None] * 4
            bq, br = (b[:2], b[2:]) if flow == "rnvp" else (None, None)
            if i < 2:
                k, c = O.conv_kl(Ps[i], ez, ea, f32(rng.standard_normal()), bern_q=bq, bern_r=br)
                O.conv_kl_backward(Ps[i], c, G=1.0 / 60000)
            else:
                k, c = O.dense_kl(Ps[i], ez, ea, bern_q=bq, bern_r=br)
                O.dense_kl_backward(Ps[i], c, G=1.0 / 60000)
            kl += float(k)
        return loss, kl, (g1, g2, g3, g4)

    return step


def time_cpu_train(flow: str = "rnvp", batch: int = 128, steps: int = 5, warmup: int = 1):
    step = cpu_train_step_factory(flow, batch)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), ts


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank: int) -> None:
    if rank != 0:
        return
    rows = args.cpu_rows
    t, _ = time_cpu(args.flow, rows, max(args.steps, 1), max(args.warmup, 1))
    val = rows / t
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {**workload_config(args, rows_per_step=rows),
                   "first_layer": "every row convolved, as the reference does (the B200 arm's mc_predict shares conv1's "
                                  "contractions between the samples of an image; same outputs)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{rows} of {ROWS} rows per step (same images, same model), NumPy float32 "
                                   f"restatement of the reference TF path, {cores} BLAS threads of {os.cpu_count()} host cores; "
                                   "TensorFlow not installable here"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, rows_per_step=ROWS) -> dict:
    return {"workload": "MNF-LeNet5 MC predictive: 1024 posterior samples x 10 synthetic 28x28x1 images, "
                        "forward of 2xMNFConv2D + 2xMNFDense (+ReLU/MaxPool/Softmax) + model.kl_div()",
            "rows_per_gpu": rows_per_step, "flow_type": args.flow, "n_flows_q": 2, "n_flows_r": 2, "flow_h_sizes": [50],
            "std_init": 10.0, "max_std": 1.0, "noise": "in-kernel Philox4x32-10", "l2": "flushed between timed steps "
            "(256 MiB memset outside the per-step event pairs)",
            "launch": "value: eager enqueue (the host runs ahead of the device); the same step replayed as one CUDA "
                      "graph is in launch_modes_ms_per_step; e2e: graph replay unless --no-graph (a device-side "
                      "replay counter advances the Philox call indices: replays are bit-identical to eager calls)",
            "weights": "frozen (Sequential.freeze_weights): packed / pre-split weight blocks are kept across steps",
            "first_layer": ("every one of the 10240 rows convolved (--no-share)" if getattr(args, "no_share", False) else
                            "model.mc_predict(images[10], 1024): in EVERY step (nothing is cached between steps) conv1's two "
                            "convolutions are evaluated once per distinct image and the noise (z, eps) is drawn per "
                            "posterior sample -- SURVEY.md 8f-2; mnf_conv.py:190-197 applies the "
                            "noise after the contractions, so outputs equal those of the 10240-row repeated batch (same "
                            "Philox counters; tests/test_gpu_mc.py); resident_tiled_batch times that batch"),
            "streams": "KL terms on 4 sibling streams and the z chains of layers 2-4 on 3 side streams, all forked from "
                       "the step's stream after its start event and joined before its stop event",
            "parallelism": f"row-sharded x{args.gpus}, no collective"}


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args, rank: int, local_rank: int, world: int) -> None:
    import torch

    import tf_mnf_b200 as M
    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200.models import MNFLeNet

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            # these levels print a version banner on STDOUT (NCCL_DEBUG_FILE does not move it); stdout carries
            # exactly one JSON line, so report the version on stderr instead
            del os.environ["NCCL_DEBUG"]
            if rank == 0:
                print("NCCL version " + ".".join(map(str, torch.cuda.nccl.version())), file=sys.stderr, flush=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    M.runtime.set_default_device(local_rank)
    M.set_seed(0)
    ctx = M.default_context(local_rank)
    ctx.set_math(args.math)
    lib = ctx.lib

    model = MNFLeNet(max_std=1, flow_h_sizes=[50], std_init=10.0, flow_type=args.flow)
    rng = np.random.RandomState(0)
    imgs = rng.uniform(size=(N_IMAGES, 28, 28, 1)).astype(np.float32)
    x_host = np.repeat(imgs, N_SAMPLES, axis=0)
    x_dev = ctx.to_device(x_host)
    model.set_row_offset(rank * ROWS)
    flush = ctx.empty((64 << 20,))  # 256 MiB > 126 MB L2

    # kl_div() does not depend on x: the four layers' terms run on sibling streams beside the
    # forward.  They are forked from the step's stream at its start and joined at its end, so all
    # of a step's work lies between its start and stop events.
    kl_lanes = [ctx.fork() for _ in range(4)]
    kl_ctx = kl_lanes[0]
    model(x_dev)  # variables are created on the first call (Keras build protocol)
    model.freeze_weights(True)  # inference: packed weight blocks are built once, not per call

    def step(x):
        for c in kl_lanes:
            c.wait(ctx)
        if os.environ.get("MNF_BENCH_NOKL"):  # diagnostic only: how much the KL streams cost the forward
            return model(x), None
        probs = model(x)                 # the x-dependent chain is enqueued first: it is the critical path
        kl = model.kl_div(ctx=kl_lanes)  # lanes were forked above, so this still starts at the step's start
        ctx.wait(kl_ctx)
        return probs, kl

    # The workload as the API states it: the 10 distinct images (resident in HBM) under 1024 posterior
    # samples each.  mc_predict lets the first layer convolve every distinct image once and draw only
    # the noise per sample (mnf_conv.py:190-197 applies z and eps after the contractions; SURVEY.md
    # 8f-2) -- same rows, same Philox counters as the forward of the 10240-row repeated batch, which
    # is timed too (`resident_tiled_batch`; --no-share makes it the headline again).
    imgs_dev = ctx.to_device(imgs)
    model.mc_share_first = not args.no_share

    def step_mc():
        for c in kl_lanes:
            c.wait(ctx)
        probs = model.mc_predict(imgs_dev, N_SAMPLES)
        kl = model.kl_div(ctx=kl_lanes)
        ctx.wait(kl_ctx)
        return probs, kl

    main_step = (lambda: step(x_dev)) if args.no_share else step_mc

    def side_launches():  # the model's own side context (z of the later layers, drawn ahead)
        return sum(c.launches for c in getattr(model, "_sides", []))

    def ev():
        p = C.c_void_p()
        L.check(lib.mnf_event_create(C.byref(p)))
        return p

    def barrier():
        ctx.sync()
        for c in kl_lanes:
            c.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    clk = ClockSampler(local_rank).__enter__()  # started early: nvidia-smi's NVML init stalls the GPU briefly
    for _ in range(max(args.warmup, 3)):
        L.check(lib.mnf_memset(ctx.handle, C.c_void_p(flush.ptr), 0, flush.nbytes))
        step(x_dev)
    barrier()
    if args.breakdown and rank == 0:  # per-call device times of one (warm) step, for DESIGN.md / profiles
        marks = []

        def mark(name):
            e = ev()
            L.check(lib.mnf_event_record(ctx.handle, e))
            marks.append((name, e))

        for rep in range(3):
            marks.clear()
            L.check(lib.mnf_memset(ctx.handle, C.c_void_p(flush.ptr), 0, flush.nbytes))
            mark("start")
            h = x_dev
            for lyr in model.layers:
                h = lyr(h)
                mark(type(lyr).__name__)
            for lyr in model.mnf_layers:
                lyr.kl_div()
                mark("kl_div:" + type(lyr).__name__)
            ctx.sync()
        parts = []
        for (n0, e0), (n1, e1) in zip(marks, marks[1:]):
            ms_ = C.c_float()
            L.check(lib.mnf_event_elapsed_ms(e0, e1, C.byref(ms_)))
            parts.append((n1, round(ms_.value * 1e3, 1)))
        print("[breakdown us] " + ", ".join(f"{n}={v}" for n, v in parts) + f"  total={sum(v for _, v in parts):.0f}",
              file=sys.stderr, flush=True)
        ctx.sync()
        t0 = time.perf_counter()
        model(x_dev)
        t1 = time.perf_counter()
        model.kl_div()
        t2 = time.perf_counter()
        ctx.sync()
        print(f"[host enqueue us] forward={(t1 - t0) * 1e6:.0f} kl_div={(t2 - t1) * 1e6:.0f}", file=sys.stderr, flush=True)
    if args.hostprof and rank == 0:  # where the host time of one step goes (stderr)
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(20):
            step(x_dev)
        pr.disable()
        barrier()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(28)
    t_dead = time.perf_counter() + 8.0  # nvidia-smi has finished initialising once it prints samples
    while clk.proc is not None and len(clk.rows) < 3 and time.perf_counter() < t_dead:
        time.sleep(0.05)
    for _ in range(2):
        L.check(lib.mnf_memset(ctx.handle, C.c_void_p(flush.ptr), 0, flush.nbytes))
        step(x_dev)
    barrier()

    def all_launches():
        return ctx.launches + sum(c.launches for c in kl_lanes) + side_launches()

    # ---- the kernels of a step timed in place: a probe (mnf_ctx_probe) brackets ONE launch family of
    # one context with CUDA events inside eager steps, side streams running (a captured graph has no
    # place for the probe's events).  One probe per context and step, so the table takes a few rounds.
    n_probe = min(args.steps, 10)
    sides = list(getattr(model, "_sides", []))
    PR_CONV, PR_DENSE, PR_FLOW, PR_KL, PR_EXPAND = 1, 2, 3, 4, 5
    share = not args.no_share
    fused12 = share and bool(getattr(model, "mc_fuse_second", False))  # conv1's expansion runs inside conv2's kernel
    plan = [
        ("conv1+conv2" if fused12 else "conv2", ctx, PR_CONV, 0 if share else 1),
        ("d1", ctx, PR_DENSE, 0),
        ("d2", ctx, PR_DENSE, 1),
    ]
    if not fused12:
        plan.insert(1, ("conv1", ctx, PR_EXPAND if share else PR_CONV, 0))
    for name, si in (("conv2.z", 1), ("d1.z", 2), ("d2.z", 3)):   # the z chains of layers 2-4 (their own streams)
        if si < len(sides):
            plan += [(f"flow step 1 of {name}", sides[si], PR_FLOW, 0), (f"flow step 2 of {name}", sides[si], PR_FLOW, 1)]
    kernel_ms: dict[str, list] = {}
    launches0 = all_launches()
    probe_steps = 0
    pending = list(plan)
    while pending:
        batch, rest, used = [], [], set()
        for item in pending:
            (rest if id(item[1]) in used else batch).append(item)
            used.add(id(item[1]))
        pending = rest
        evs = {name: ([ev() for _ in range(n_probe)], [ev() for _ in range(n_probe)]) for name, *_ in batch}
        for i in range(n_probe):
            L.check(lib.mnf_memset(ctx.handle, C.c_void_p(flush.ptr), 0, flush.nbytes))
            for name, c, kind, skip in batch:
                L.check(lib.mnf_ctx_probe(c.handle, kind, skip, evs[name][0][i], evs[name][1][i]))
            main_step()
            probe_steps += 1
        barrier()
        for name, c, kind, skip in batch:
            L.check(lib.mnf_ctx_probe(c.handle, 0, 0, None, None))
            ms_ = C.c_float()
            vals = []
            for a_, b_ in zip(*evs[name]):
                if lib.mnf_event_elapsed_ms(a_, b_, C.byref(ms_)) == 0:
                    vals.append(ms_.value)
            kernel_ms[name] = vals
    launches_per_step = (all_launches() - launches0) // max(probe_steps, 1)
    kern_ms = kernel_ms.get("conv2") or [float("nan")]

    # ---- timed: K steps, device-resident input, per-step CUDA events, L2 flushed in between.
    # The step is captured once into a CUDA graph (tf_mnf_b200/graph.py) and replayed: one launch
    # per step instead of ~30; a device-side replay counter keeps the Philox call indices advancing,
    # so replay i is bit-identical to eager call i (tests/test_gpu_streams.py).
    graph = None
    if not args.no_graph:
        from tf_mnf_b200.graph import CapturedStep

        try:
            graph = CapturedStep(ctx, main_step, forks=kl_lanes + list(model._sides), layers=model.mnf_layers,
                                 models=[model])
            for _ in range(3):
                graph.launch()
            barrier()
        except Exception as exc:  # never lose the measurement to the optional launch mode
            print(f"[rank {rank}] CUDA-graph capture unavailable ({exc!r}); eager steps only", file=sys.stderr, flush=True)
            graph, args.no_graph = None, True
            barrier()
    def timed_steps(run_step):
        starts, stops = [ev() for _ in range(args.steps)], [ev() for _ in range(args.steps)]
        t0 = time.perf_counter()
        for i in range(args.steps):
            L.check(lib.mnf_memset(ctx.handle, C.c_void_p(flush.ptr), 0, flush.nbytes))
            L.check(lib.mnf_event_record(ctx.handle, starts[i]))
            run_step()
            L.check(lib.mnf_event_record(ctx.handle, stops[i]))
        barrier()
        t1 = time.perf_counter()
        ms = C.c_float()
        per = []
        for a_, b_ in zip(starts, stops):
            L.check(lib.mnf_event_elapsed_ms(a_, b_, C.byref(ms)))
            per.append(ms.value)
        return per, t0, t1

    # both ways of launching the same step are timed (K steps each); the faster one is `value`
    modes = {"eager": timed_steps(main_step)}
    if graph is not None:
        modes["graph"] = timed_steps(graph.launch)
    tiled_ms = float(np.mean(timed_steps(lambda: step(x_dev))[0])) if not args.no_share else None
    launch_mode = "eager"  # `value` is the plain API call; the graph replay of the same step is reported beside it
    step_ms, t_wall0, t_wall1 = modes[launch_mode]
    t_wall = t_wall1 - t_wall0
    clk.__exit__()
    launches = launches_per_step * args.steps  # kernels per timed step (counted on eager steps; a graph replays the same ones)

    def elapsed(a, b):
        ms = C.c_float()
        L.check(lib.mnf_event_elapsed_ms(a, b, C.byref(ms)))
        return ms.value

    my_ms = float(np.mean(step_ms))
    print(f"[rank {rank}] step_ms={np.round(step_ms, 3).tolist()} conv2_ms={np.round(kern_ms, 3).tolist()} "
          f"wall_per_step_ms={t_wall / args.steps * 1e3:.3f} launch={launch_mode} "
          + " ".join(f"{k}_ms={np.mean(v[0]):.3f}" for k, v in modes.items()), file=sys.stderr, flush=True)

    # ---- e2e: host (pinned) input -> public API -> host result, copies inside the timed region.
    # The call a user makes for this workload is model.mc_predict(images, n_samples): the 10
    # distinct images cross PCIe and are repeated on the device (the reference's notebooks call the
    # model n_samples times on one resident batch).  The same step fed with the already-repeated
    # 10240-row batch from the host (32 MB over PCIe per step) is timed too and reported beside it.
    res_host = C.c_void_p()
    L.check(lib.mnf_host_alloc(ROWS * 10 * 4 + 16, C.byref(res_host)))

    def host_buf(arr):
        hp_ = C.c_void_p()
        L.check(lib.mnf_host_alloc(arr.nbytes, C.byref(hp_)))
        C.memmove(hp_.value, arr.ctypes.data, arr.nbytes)
        return hp_

    hp_img, hp_full = host_buf(imgs), host_buf(x_host)
    img_in, x_in = ctx.empty(imgs.shape), ctx.empty(x_host.shape)

    def finish(probs, kl):
        L.check(lib.mnf_memcpy_d2h(ctx.handle, res_host, C.c_void_p(probs.ptr), probs.nbytes))
        L.check(lib.mnf_memcpy_d2h(ctx.handle, C.c_void_p(res_host.value + probs.nbytes), C.c_void_p(kl.ptr), 4))

    def enqueue_mc():
        L.check(lib.mnf_memcpy_h2d(ctx.handle, C.c_void_p(img_in.ptr), hp_img, imgs.nbytes))
        for c in kl_lanes:
            c.wait(ctx)
        probs = model.mc_predict(img_in, N_SAMPLES)
        kl = model.kl_div(ctx=kl_lanes)
        ctx.wait(kl_ctx)
        finish(probs, kl)

    def enqueue_tiled():
        L.check(lib.mnf_memcpy_h2d(ctx.handle, C.c_void_p(x_in.ptr), hp_full, x_host.nbytes))
        probs, kl = step(x_in)
        finish(probs, kl)

    def synced(enqueue):
        def run():
            enqueue()
            ctx.sync()
            for c in kl_lanes:
                c.sync()
        run.__wrapped__ = enqueue
        return run

    e2e_mc, e2e_tiled = synced(enqueue_mc), synced(enqueue_tiled)

    def time_e2e(fn):
        if not args.no_graph:  # the whole step incl. its H2D / D2H copies (pinned memory) as one graph launch
            try:
                g = CapturedStep(ctx, fn.__wrapped__, forks=kl_lanes + list(model._sides), layers=model.mnf_layers,
                                 models=[model])

                def fn():  # noqa: F811
                    g.launch()
                    ctx.sync()
            except Exception as exc:
                print(f"[rank {rank}] e2e graph capture unavailable ({exc!r}); eager", file=sys.stderr, flush=True)
                barrier()
        fn()
        barrier()
        out_ms = []
        for _ in range(args.steps):
            L.check(lib.mnf_memset(ctx.handle, C.c_void_p(flush.ptr), 0, flush.nbytes))
            ctx.sync()
            a, b = ev(), ev()
            L.check(lib.mnf_event_record(ctx.handle, a))
            fn()
            L.check(lib.mnf_event_record(ctx.handle, b))
            ctx.sync()
            out_ms.append(elapsed(a, b))
        barrier()
        return float(np.mean(out_ms))

    my_e2e = time_e2e(e2e_mc)
    my_e2e_tiled = time_e2e(e2e_tiled)

    # ---- BASELINE configs[4] beside the headline: the MNF VGG-style net's data-parallel train step, the one
    # place of the design with a collective (mnf_allreduce_grads: NCCL over NVLink, inside the C ABI)
    c5 = None
    if not args.no_train:
        try:
            c5 = time_c5_train(ctx, rank, world, dist, barrier, ev, elapsed, args.train_steps)
        except Exception as exc:  # never lose the headline to the secondary measurement
            print(f"[rank {rank}] C5 train step unavailable ({exc!r})", file=sys.stderr, flush=True)
            barrier()

    # ---- max over ranks
    names = sorted(kernel_ms)
    mine = [my_ms, my_e2e, my_e2e_tiled] + [float(np.mean(kernel_ms[n])) if kernel_ms[n] else float("nan") for n in names]
    if dist is not None:
        t = torch.tensor(mine, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mine = [float(v) for v in t.tolist()]
    ms, e2e_ms, e2e_tiled_ms = mine[:3]
    kms = dict(zip(names, mine[3:]))
    k_ms = kms.get("conv2", kms.get("conv1+conv2", float("nan")))
    fused12 = "conv1+conv2" in kms

    if rank == 0:
        hbm, tc_burst, tc_sus, which = peaks()
        tc = args.math == "tf32x3"
        # conv2's arithmetic intensity (263 FLOP/B fused) is above the ridge (254), and the fp32-parity mode
        # spends THREE kind::f16 MMAs per product (hi*hi + lo*hi + hi*lo): its roof is the tensor pipe / 3.
        # `peak` is the burst figure (the kernel is timed alone by its probe), per the contract.
        k_flop = CONV2_FLOP_PER_ROW * ROWS
        k_bytes = CONV2_BYTES_PER_ROW * ROWS + CONV2_WEIGHT_BYTES
        ach_tf = k_flop / (k_ms * 1e-3) / 1e12
        ffma_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        peak_tf = tc_burst / 3 if tc else ffma_peak
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if tc and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj.get("conv2_dram_bytes_per_launch")
        table = {}
        for name, (bpr, wb, fpr) in STAGES.items():
            key = name.split(" ")[0]
            if key in kms and kms[key] == kms[key]:
                t_ms = kms[key]
                table[name] = {"ms": t_ms, "share_of_step": t_ms / ms, "algorithmic_GBps": (bpr * ROWS + wb) / (t_ms * 1e-3) / 1e9,
                               "hbm_frac": (bpr * ROWS + wb) / (t_ms * 1e-3) / 1e9 / hbm,
                               "algorithmic_TFLOPs": fpr * ROWS / (t_ms * 1e-3) / 1e12}
        for name in names:
            if name.startswith("flow"):
                table[name] = {"ms": kms[name], "stream": "side (beside the main stream)"}
        step_roof = hbm * 1e9 / LENET_BYTES_PER_ROW
        line = {
            "metric": METRIC, "value": ROWS * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": ROWS * world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(imgs.nbytes),
                    "d2h_bytes_per_step": ROWS * 10 * 4 + 4, "ms_per_step": e2e_ms,
                    "api": "model.kl_div() + model.mc_predict(images[10,28,28,1] from pinned host memory, 1024): the "
                           "distinct images cross PCIe and are repeated on the device; probabilities [10240,10] and "
                           "the KL scalar are read back every step",
                    "host_repeated_batch": {"value": ROWS * world / (e2e_tiled_ms * 1e-3), "ms_per_step": e2e_tiled_ms,
                                            "h2d_bytes_per_step": int(x_host.nbytes),
                                            "note": "same step fed with the 10240-row repeated batch from the host"}},
            "gpu_launches": int(launches),
            "clocks": clk.summary(min(v[1] for v in modes.values()), max(v[2] for v in modes.values())),
            "launch_modes_ms_per_step": {k: float(np.mean(v[0])) for k, v in modes.items()},
            "resident_tiled_batch": None if tiled_ms is None else {
                "value": ROWS / (tiled_ms * 1e-3), "unit": UNIT, "ms_per_step": tiled_ms, "n_gpus": 1,
                "note": "rank 0, eager: model(x) on the resident 10240-row repeated batch + kl_div() -- the first layer "
                        "convolves every row (no sharing between the samples of an image)"},
            "roofline": {"bound": "tensor",
                         "kernel": (("cf::conv_raw_f16x2<gen>" if fused12 else "cf::conv_raw_f16x2") if tc else "paired_gemm_fwd<CONV,64>")
                         + " (conv2: 12x12x20 -> 8x8x50, 5x5, + fused ReLU/2x2 max-pool"
                         + ("; its plane producers also generate the input rows = conv1's per-sample noise expansion + ReLU + pool, "
                            "which is NOT counted in `achieved`" if fused12 else "") + "), the dominant kernel of the step",
                         "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel per launch, from the ncu --set full
                         # capture named in profiles/roofline_traffic.json (null when no capture of this round exists)
                         "traffic": traffic,
                         "peak_source": which + (": dense bf16/fp16 burst / 3 (three MMAs per fp32-parity product)" if tc else
                                                 ": FFMA 148 SM x 128 lanes x 2 x 1.965 GHz"),
                         "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms,
                         "algorithmic_flop_per_launch": k_flop, "algorithmic_bytes_per_launch": k_bytes,
                         "hbm_view": {"achieved_GBps": k_bytes / (k_ms * 1e-3) / 1e9, "peak_GBps": hbm,
                                      "frac": k_bytes / (k_ms * 1e-3) / 1e9 / hbm,
                                      "note": "the same launch against the HBM roof: algorithmic bytes of the FUSED kernel "
                                              "(pooled output); not what binds it"},
                         "pipe": ("tcgen05.mma cta_group::2 kind::f16 (M = 256 over a CTA pair), power-of-two scaled fp16 hi/lo "
                                  "split of both operands, fp32 accumulate (fp32 parity <=1e-5 rel.); kernel_ms is the in-situ "
                                  "time, side streams running"
                                  if tc else "fp32 FFMA (parity path, <=1e-5 rel.)"),
                         "step": {"rows_per_s_per_gpu": ROWS / (ms * 1e-3), "hbm_roof_rows_per_s": step_roof,
                                  "frac": ROWS / (ms * 1e-3) / step_roof,
                                  "tensor_roof_rows_per_s": tc_sus * 1e12 / 3 / LENET_FLOP_PER_ROW,
                                  "frac_of_tensor_roof": ROWS / (ms * 1e-3) / (tc_sus * 1e12 / 3 / LENET_FLOP_PER_ROW),
                                  "no_share_frac": None if tiled_ms is None else ROWS / (tiled_ms * 1e-3) / step_roof,
                                  "note": "whole step against SURVEY.md 8(d): 81,275 B/row at the layer-API boundary -> HBM roof; "
                                          "9.994 MFLOP/row x 3 MMAs per fp32-parity product -> tensor roof (sustained bf16 / 3); "
                                          "no_share = every row convolved by conv1 (resident_tiled_batch)"},
                         "kernels": table},
            "wall_s_timed_region": t_wall,
        }
        if c5 is not None:
            line["c5_train_step"] = c5
        if world == 1 and not args.no_cpu:
            rows = args.cpu_rows
            t_cpu, _ = time_cpu(args.flow, rows, args.cpu_steps, 1)
            cores = blas_threads()
            line["cpu_baseline"] = {
                "value": rows / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{args.cpu_steps} steps of {rows} of {ROWS} rows (same images/model), NumPy float32 "
                          f"restatement of the reference TF path (oracle/), {cores} BLAS threads of {os.cpu_count()} host cores"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def time_c5_train(ctx, rank, world, dist, barrier, ev, elapsed, steps: int = 8):
    """BASELINE.json configs[4]: MNF VGG-style net, synthetic 32x32x3, GLOBAL batch 256 split
    over the ranks (strong scaling), one training step = forward + cross-entropy + backward + kl_div + KL
    backward + ONE in-place NCCL all-reduce of the flat fp32 gradient buffer (mnf_allreduce_grads, C ABI)
    + Adam.  Device time per step (CUDA events on the context's stream), max over ranks."""
    import torch

    import tf_mnf_b200 as M
    from tf_mnf_b200 import _lib as L
    from tf_mnf_b200.models import MNFVGG
    from tf_mnf_b200.train import Trainer

    GB = 256
    M.set_seed(0)  # same parameters and noise keys on every rank
    model = MNFVGG()
    rng = np.random.RandomState(0)
    x = rng.uniform(size=(GB, 32, 32, 3)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.randint(0, 10, size=GB)]
    lo, hi = rank * GB // world, (rank + 1) * GB // world
    model.set_row_offset(lo)
    xd, yd = ctx.to_device(x[lo:hi]), ctx.to_device(y[lo:hi])
    tr = Trainer(model, loss="xent", kl_weight=1.0 / 50000, world=world)
    for _ in range(3):
        tr.step(xd, yd)
    barrier()
    starts, stops = [ev() for _ in range(steps)], [ev() for _ in range(steps)]
    t0 = time.perf_counter()
    for i in range(steps):
        L.check(ctx.lib.mnf_event_record(ctx.handle, starts[i]))
        tr.step(xd, yd)
        L.check(ctx.lib.mnf_event_record(ctx.handle, stops[i]))
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / steps
    dev_ms = float(np.mean([elapsed(a, b) for a, b in zip(starts, stops)]))
    if dist is not None:
        t = torch.tensor([wall_ms, dev_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall_ms, dev_ms = (float(v) for v in t.tolist())
    n_params = int(tr.flat_g.size)
    out = {"workload": "MNF VGG-style conv net (6 x MNFConv2D 3x3 SAME + 2 x MNFDense, IAF flows), synthetic 32x32x3, "
                       "global batch 256, data-parallel train step (BASELINE.json configs[4])",
           "value": GB / (wall_ms * 1e-3), "unit": "images/s", "n_gpus": world, "scaling": "strong", "steps": steps,
           "ms_per_step": wall_ms, "device_ms_per_step": dev_ms, "per_gpu_batch": hi - lo,
           "timing": "wall clock barrier to barrier per step, max over ranks (device-event time beside it)",
           "collective": None if world == 1 else {
               "op": "ncclAllReduce(sum), in place, fp32, on the context's stream (mnf_allreduce_grads)",
               "bytes": n_params * 4, "nccl_version": tr.comm.nccl_version if tr.comm is not None else None}}
    if tr.comm is not None:
        tr.comm.close()
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--flow", default="rnvp", choices=["rnvp", "iaf", "planar"])
    ap.add_argument("--math", default="tf32x3", choices=["tf32x3", "fp32"],
                    help="contraction arithmetic: tcgen05 3xTF32 (fp32 parity) or FFMA")
    ap.add_argument("--cpu-rows", type=int, default=2560, help="rows per CPU-baseline step (bounded sample)")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--hostprof", action="store_true", help="cProfile of 20 steps' host-side enqueue (stderr)")
    ap.add_argument("--no-share", action="store_true",
                    help="forward the 10240-row repeated batch instead of mc_predict(images, 1024): the first layer then "
                         "convolves every row")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--breakdown", action="store_true", help="print per-layer device times of one step (stderr)")
    ap.add_argument("--no-train", action="store_true", help="skip the C5 (MNF-VGG data-parallel train step) measurement")
    ap.add_argument("--train-steps", type=int, default=8)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
