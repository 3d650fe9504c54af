#!/usr/bin/env python
"""bench.py - Jacobians/s of the jacve hot path on the RoeFlux_3d vmap batch (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--order mixed] [--impl reference]

One "step" = one pass of the hot path over one batch of 2^20 synthetic samples per GPU (weak scaling,
no collective on the data path).  `value` is whole-job throughput with inputs resident in HBM, timed
with CUDA events (max over ranks); `e2e` is the same metric through the public API with HOST buffers
(pinned host -> device copies and the device -> host read of the Jacobians inside the timed region);
`roofline` relates the kernel to the measured HBM copy bandwidth over the compulsory 240 B/Jacobian;
`cpu_baseline` / `--impl reference` time the reference on the host cores: graphax's OWN sources (the
`baseline/_ref` install) in the scalable form `jax.vmap(graphax.jacve(f, order, argnums))`, with JAX - which this
image does not have - replaced by the NumPy stand-in `tests/golden/jaxlite`; if that install is missing, the
oracle's NumPy port of the algorithm.  The headline runs the default plan (`rounding="reference"`: bit-identical
to that reference arithmetic); the contracted `rounding="fma"` plan is reported beside it with its parity record.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from graphax_b200.configs import WORKLOADS  # noqa: E402

METRIC = "Jacobians/s, RoeFlux_3d vmap batch"
UNIT = "Jacobians/s"
WORKLOAD = "roe3d"       # headline config; --workload selects the other named configs (not bench lines of record)
TITLES = {"roe3d": "RoeFlux_3d", "roe1d": "RoeFlux_1d", "robotarm": "RobotArm_6DOF", "readme": "README f"}
N_BUFFER_SETS = 4        # 4 x (40 MB in + 200 MB out) = 960 MB >> 126 MB L2: every step touches cold memory


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's NumPy port of the reference algorithm, all host cores
# ------------------------------------------------------------------------------------------------
_CPU_CTX = {}


def _cpu_chunk(bounds):
    from oracle import ve_numpy
    w = WORKLOADS[WORKLOAD]
    lo, hi = bounds
    xs = [x[lo:hi] for x in _CPU_CTX["xs"]]
    jac, _, _ = ve_numpy.jacve(w.jaxpr(), _CPU_CTX["order"], w.argnums, xs, np.float32)
    return float(sum(np.abs(b).sum() for row in jac for b in row))


def cpu_port_throughput(order_key: str, sample: int, reps: int, cores: int):
    """Jacobians/s of the NumPy restatement (vectorised over the batch) on `cores` processes."""
    w = WORKLOADS[WORKLOAD]
    _CPU_CTX["xs"] = w.inputs(sample)
    _CPU_CTX["order"] = w.order(order_key)
    chunk = 4096
    bounds = [(lo, min(lo + chunk, sample)) for lo in range(0, sample, chunk)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_chunk, bounds[:cores])                       # warm-up (imports, page faults)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            pool.map(_cpu_chunk, bounds)
            times.append(time.perf_counter() - t0)
    return sample / float(np.median(times)), times


def _reference_available() -> bool:
    return (ROOT / "baseline" / "_ref" / "graphax" / "core.py").exists() and \
        (ROOT / "tests" / "golden" / "jaxlite" / "jax" / "__init__.py").exists()


def _ref_chunk(bounds):
    """One chunk through the reference's own `jax.vmap(graphax.jacve(RoeFlux, order, argnums))` (worker process)."""
    if "fun" not in _CPU_CTX:
        try:                                   # a box that has the real thing: the reference on JAX / XLA itself
            import jax
            sys.path.insert(0, str(ROOT / "baseline" / "_ref"))
        except ImportError:
            sys.path[:0] = [str(ROOT / "tests" / "golden" / "jaxlite"), str(ROOT / "baseline" / "_ref")]
            import jax
        import jax.numpy as jnp
        import graphax
        from graphax import examples as gex
        assert "_ref" in graphax.__file__
        w = WORKLOADS[WORKLOAD]
        fun = {"roe3d": gex.RoeFlux_3d, "roe1d": gex.RoeFlux_1d, "robotarm": gex.RobotArm_6DOF}[WORKLOAD]
        order = _CPU_CTX["order"]
        _CPU_CTX["fun"] = jax.vmap(graphax.jacve(fun, order=order if isinstance(order, str) else list(order),
                                                 argnums=tuple(w.argnums)))
        _CPU_CTX["jnp"] = jnp
    lo, hi = bounds
    jnp = _CPU_CTX["jnp"]
    out = _CPU_CTX["fun"](*[jnp.array(x[lo:hi]) for x in _CPU_CTX["xs"]])
    return float(sum(np.abs(np.asarray(b)).sum() for row in out for b in row))


def cpu_reference_throughput(order_key: str, sample: int, reps: int, cores: int, warmup: int = 1):
    """Jacobians/s of graphax's own sources on the NumPy stand-in for JAX, `cores` worker processes."""
    w = WORKLOADS[WORKLOAD]
    _CPU_CTX.pop("fun", None)
    _CPU_CTX["xs"] = w.inputs(sample)
    _CPU_CTX["order"] = w.order(order_key)
    chunk = 1 << 15
    bounds = [(lo, min(lo + chunk, sample)) for lo in range(0, sample, chunk)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(max(1, warmup)):
            pool.map(_ref_chunk, bounds, chunksize=1)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            pool.map(_ref_chunk, bounds, chunksize=1)
            times.append(time.perf_counter() - t0)
    return sample / float(np.median(times)), times


def cpu_replay_throughput(order_key: str, sample: int, reps: int, cores: int):
    """Jacobians/s of the oracle's C plan replay (fused, vectorised across samples, pthreads)."""
    from oracle import replay
    w = WORKLOADS[WORKLOAD]
    plan = w.plan(order_key)
    xs = w.inputs(sample)
    replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols, n_threads=cores)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols, n_threads=cores)
        times.append(time.perf_counter() - t0)
    return sample / float(np.median(times))


def run_reference_arm(args) -> None:
    """--impl reference: the reference's CPU implementation of the path on all host cores, same workload, batch,
    order and metric as the GPU arm.  graphax needs JAX, which this image does not have: its own sources (the
    `baseline/_ref` install, unmodified) run on the NumPy stand-in `tests/golden/jaxlite` (kind "reference"); if the
    install is missing, the oracle's NumPy port of the algorithm is timed instead (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    sample = args.batch
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 5))
    t0 = time.perf_counter()
    if _reference_available():
        kind = "reference"
        how = ("graphax's own sources (baseline/_ref, unmodified) as jax.vmap(graphax.jacve(f, order, argnums)), JAX "
               "replaced by the NumPy stand-in tests/golden/jaxlite (JAX itself is not installed in this image)")
        value, times = cpu_reference_throughput(args.order, sample, steps, cores, warmup)
    else:
        kind = "port"
        how = "oracle/ve_numpy.py: NumPy port of the reference algorithm (baseline/_ref not present)"
        value, times = cpu_port_throughput(args.order, sample, steps + warmup, cores)
        times = times[warmup:]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * float(np.median(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{TITLES[WORKLOAD]} jacve order={args.order}, vmap batch {sample}, fp32 on the host CPU",
                   "batch_per_gpu": sample, "order": args.order, "implementation": how},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} samples x {steps} steps, fp32, {cores} processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Polls SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int, period_s: float = 0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def sample_once(self):
        if not self.ok:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self.sample_once()
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measured_peak_gbs():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def recorded_traffic(order_key: str):
    """DRAM bytes per launch from the committed ncu --set full capture, if one exists for this order."""
    path = ROOT / "profiles" / "traffic.json"
    if path.exists():
        try:
            return json.loads(path.read_text()).get(f"{WORKLOAD}:{order_key}")
        except Exception:
            return None
    return None


def parity_record(order_key: str, rounding: str):
    """Full-batch parity of this (order, rounding) against the float32 oracle, from the committed record of
    tests/test_gpu_full_parity.py (profiles/parity_full_r02.json); None if that order was not recorded."""
    path = ROOT / "profiles" / "parity_full_r02.json"
    if not path.exists():
        return None
    try:
        rec = json.loads(path.read_text())
    except Exception:
        return None
    key = f"{WORKLOAD}:{order_key}:reference" if rounding == "reference" else f"{WORKLOAD}:{order_key}:fma:reciprocal"
    return rec.get(key)


def _time_launches(launch, steps: int, stream, world: int, dist, dev) -> float:
    """ms per step of `launch(i)` over `steps` back-to-back steps (CUDA events, max over ranks)."""
    import torch
    for i in range(5):
        launch(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        launch(i)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _graph_time(enqueue, per_graph: int, replays: int, dev) -> float:
    """ms per step when `per_graph` consecutive steps are captured into one CUDA graph and replayed `replays` times."""
    import torch
    cap = torch.cuda.Stream(device=dev)
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            cs = torch.cuda.current_stream()
            for i in range(per_graph):
                enqueue(i, cs.cuda_stream)
    torch.cuda.current_stream().wait_stream(cap)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (replays * per_graph)


def measure_other_configs(local_rank: int, peak_gbs: float) -> dict:
    """The other named configs of BASELINE.json on this GPU (N = 1 only; reported, not the headline):
    config 2 RoeFlux_1d 2^16 'rev', config 4 RobotArm_6DOF 2^22 'rev' and Markowitz, config 5 the MLP
    512 -> 1024 -> 512 weight-Jacobian step at batch 4096 (bf16 operands, fp32 accumulate)."""
    import torch
    import graphax_b200 as gx
    dev = f"cuda:{local_rank}"
    out = {}
    stream = torch.cuda.current_stream()
    for name, wname, order_key in (("RoeFlux_1d 2^16 rev", "roe1d", "rev"), ("RobotArm_6DOF 2^22 rev", "robotarm", "rev"),
                                   ("RobotArm_6DOF 2^22 markowitz", "robotarm", "markowitz")):
        try:
            w = WORKLOADS[wname]
            batch = w.batch
            host = w.inputs(batch)
            rec = {}
            for rounding in ("reference", "fma"):
                jf = gx.jacve(w.fun, order=w.order(order_key), argnums=w.argnums, rounding=rounding)
                plan = jf.lower(w.arg_shapes)
                dp = jf._device_plan(plan, local_rank)
                nsets = 4 if plan.bytes_per_jacobian * batch * 4 < (8 << 30) else 2
                sets = [([torch.from_numpy(x).to(dev) for x in host],
                         torch.empty((batch, plan.n_rows, plan.n_cols), device=dev)) for _ in range(nsets)]

                def enqueue(i, cuda_stream, dp=dp, sets=sets, batch=batch):
                    ins, o = sets[i % len(sets)]
                    dp.launch(batch, [t.data_ptr() for t in ins], o.data_ptr(), None, cuda_stream)
                if plan.bytes_per_jacobian * batch < (64 << 20):      # launch-bound: replay from a CUDA graph
                    ms, how = _graph_time(enqueue, 40, 25, dev), "CUDA graph of 40 steps"
                else:
                    ms = _time_launches(lambda i: enqueue(i, stream.cuda_stream), 100, stream, 1, None, dev)
                    how = "one host call per step"
                gbs = plan.bytes_per_jacobian * batch / (ms * 1e-3) / 1e9
                rec[rounding] = {"value": batch / (ms * 1e-3), "unit": UNIT, "us_per_step": ms * 1e3,
                                 "roofline_frac": gbs / peak_gbs, "achieved_gbs": gbs, "kernel": dp.kernel, "launch": how}
                del sets
                torch.cuda.empty_cache()
            out[name] = dict(rec["reference"], bound="hbm", batch=batch, bytes_per_jacobian=plan.bytes_per_jacobian,
                             fma_rounding=rec["fma"])
        except Exception as e:                               # a side measurement must never take the bench line down
            out[name] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    try:
        from graphax_b200.dense import DenseExecutor, build_dense_plan
        from graphax_b200.examples import MLP_loss
        from graphax_b200.tracer import make_jaxpr
        B, I, H, O = 4096, 512, 1024, 512
        g = torch.Generator(device=dev).manual_seed(0)
        margs = (torch.randn(B, I, device=dev, generator=g).bfloat16(),
                 (torch.randn(H, I, device=dev, generator=g) * I ** -0.5).bfloat16(),
                 torch.randn(H, device=dev, generator=g) * 0.1,
                 (torch.randn(O, H, device=dev, generator=g) * H ** -0.5).bfloat16(),
                 torch.randn(O, device=dev, generator=g) * 0.1,
                 torch.randn(B, O, device=dev, generator=g).bfloat16())
        ex = DenseExecutor(build_dense_plan(make_jaxpr(MLP_loss, [tuple(t.shape) for t in margs]), "rev", (1, 2, 3, 4)),
                           local_rank, precision="bf16")
        for _ in range(3):
            ex.run(margs)
        torch.cuda.synchronize()
        l0 = ex.launches
        ex.run(margs)
        per_step = ex.launches - l0
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            ex.run(margs)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                ex.run(margs)
        for _ in range(5):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 200 * 1e3
        flops = 2 * B * (2 * I * H + 3 * H * O)
        peak_tf, sustained_tf = 1667.8, None
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            peak_tf = float(peaks["bf16_tflops"])
            sustained_tf = float(peaks["bf16_tflops_sustained"]) if "bf16_tflops_sustained" in peaks else None
        except Exception:
            pass
        out["MLP 512-1024-512 weight Jacobians, batch 4096, bf16 operands / fp32 accumulate, rev"] = {
            "value": 1e6 / us, "unit": "steps/s", "us_per_step": us, "tflops": flops / us / 1e6, "bound": "tensor",
            "roofline_frac": flops / us / 1e6 / peak_tf, "peak_tflops": peak_tf, "launches_per_step": per_step,
            "launch": "CUDA graph of one step", "flops_per_step": flops,
            # 200 replays = a 10 ms burst: the burst peak (cuBLAS 8192^3, best of 10) is the denominator; the same
            # library sustains less over seconds (power), given here for scale only
            "frac_of_sustained_peak": (flops / us / 1e6 / sustained_tf) if sustained_tf else None}
    except Exception as e:
        out["MLP 512-1024-512 weight Jacobians"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    return out


def run_gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist
    import graphax_b200 as gx
    from graphax_b200 import runtime
    from graphax_b200.plan import DEFAULT_ROUNDING
    from graphax_b200.sharding import bind_to_local_numa_node, gather_for_verification, init_process_group

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device. graphax_b200 has no CPU fallback (use --impl reference for the CPU arm).")
    # stdout carries exactly one JSON line: libraries that print there (NCCL announces its version on the first
    # collective) are sent to stderr, the result goes to the saved descriptor
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    rank, local_rank, world = init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    numa = bind_to_local_numa_node(local_rank) if world > 1 and os.environ.get("GX_NUMA_BIND", "0") == "1" else "off"
    dev = f"cuda:{local_rank}"
    w = WORKLOADS[WORKLOAD]
    batch = args.batch                                   # per GPU (weak scaling)
    order = w.order(args.order)
    rounding = args.rounding or DEFAULT_ROUNDING
    jacfun = gx.jacve(w.fun, order=order, argnums=w.argnums, rounding=rounding)
    plan = jacfun.lower(w.arg_shapes)
    dp = jacfun._device_plan(plan, local_rank)
    if args.kernel != "auto":
        dp.select_kernel({"interpreter": 0, "aot": 1, "jit": 2}[args.kernel])
    info = dp.info()
    peak, peak_src = measured_peak_gbs()

    # synthetic inputs of this rank's shard; pinned on the host (for e2e) and resident in HBM (for value)
    host_in = [torch.from_numpy(x).pin_memory() for x in w.inputs(batch, start=rank * batch)]
    sets = []
    for s in range(N_BUFFER_SETS):
        ins = [t.to(dev, non_blocking=True) for t in host_in]
        sets.append((ins, torch.empty((batch, plan.n_rows, plan.n_cols), device=dev, dtype=torch.float32)))
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()

    def step(i: int):
        ins, out = sets[i % N_BUFFER_SETS]
        dp.launch(batch, [t.data_ptr() for t in ins], out.data_ptr(), None, stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(args.warmup, 3)):
        step(i)
    sync_all()

    # Launch-bound configurations (a step of a few microseconds: RoeFlux_1d at 2^16 is 6 MB per launch) are
    # bound by the host's launch rate when every step is an individual ctypes call.  They are replayed from a
    # CUDA graph of `per_graph` consecutive steps instead: the same kernels on the same rotating buffers, only
    # enqueued by the driver in one go.
    step_bytes = plan.bytes_per_jacobian * batch
    per_graph = 0
    if args.graph == "on" or (args.graph == "auto" and step_bytes < (64 << 20)):
        per_graph = 40
        while args.steps % per_graph:
            per_graph -= 1
    graph = None
    if per_graph > 1:
        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap):
                cs = torch.cuda.current_stream()
                for i in range(per_graph):
                    ins, out = sets[i % N_BUFFER_SETS]
                    dp.launch(batch, [t.data_ptr() for t in ins], out.data_ptr(), None, cs.cuda_stream)
        torch.cuda.current_stream().wait_stream(cap)
        for _ in range(3):
            graph.replay()
        sync_all()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = dp.info().launches
    sync_all()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    sampler.sample_once()
    # hold the stream for ~0.1 ms while the host enqueues the start event and the first launches: the timed region
    # then begins with the first step already queued behind the event instead of with the host's launch latency
    # (matters only for small --steps; the K timed steps themselves are unchanged)
    with torch.cuda.stream(stream):
        torch.cuda._sleep(200_000)
    ev[0].record(stream)
    if graph is not None:
        for _ in range(args.steps // per_graph):
            graph.replay()
    else:
        for i in range(args.steps):
            step(i)
    ev[1].record(stream)
    sampler.sample_once()
    torch.cuda.synchronize()
    sampler.sample_once()
    elapsed_ms = ev[0].elapsed_time(ev[1])
    launches = args.steps if graph is not None else dp.info().launches - launches0
    # keep the GPU under the same load a little longer so that the clock record has enough samples
    t_probe = time.perf_counter()
    i = 0
    while time.perf_counter() - t_probe < args.clock_probe_s:
        for _ in range(64):
            step(i)
            i += 1
        torch.cuda.synchronize()
    sampler.stop()
    elapsed_ms = max_over_ranks(elapsed_ms)
    sync_all()

    # ---- verification: every rank contributes a window of its shard's result through an NCCL all-gather; rank 0
    # compares all of them with the oracle on the same global sample indices -----------------------------------
    verify = None
    if not args.no_verify:
        rows = 1024
        lo = (batch - rows) // 2 if batch > rows else 0
        n = min(rows, batch)
        ins, out_t = sets[0]
        dp.launch(batch, [t.data_ptr() for t in ins], out_t.data_ptr(), None, stream.cuda_stream)
        torch.cuda.synchronize()
        window = out_t[lo:lo + n].contiguous()
        gathered = gather_for_verification(window, [n] * world)          # NCCL all-gather at world > 1
        torch.cuda.synchronize()
        if rank == 0:
            from oracle import replay as oracle_replay
            from oracle import ve_numpy
            got = gathered.cpu().numpy().reshape(world, n, plan.n_rows, plan.n_cols)
            bit_exact = values_equal_reference = True
            for r in range(world):
                xs = w.inputs(n, start=r * batch + lo)
                ref_j, _ = oracle_replay.replay(plan.blob, xs, plan.n_rows, plan.n_cols)
                bit_exact &= bool(np.array_equal(got[r].view(np.uint32), ref_j.view(np.uint32)))
                if rounding == "reference":
                    j32, _, _ = ve_numpy.jacve(w.jaxpr(), order, w.argnums, xs, np.float32)
                    t32 = np.concatenate([np.concatenate([np.asarray(b).reshape(n, -1, sz) for b, sz in zip(row, w.sizes)], axis=2)
                                          for row in j32], axis=1)
                    values_equal_reference &= bool(np.array_equal(got[r], t32))
            verify = {"ranks": world, "rows": world * n, "window_per_rank": [lo, lo + n],
                      "collective": "NCCL all_gather" if world > 1 else "none (single rank)",
                      "bit_exact": bit_exact, "against": "oracle/replay.c on the plan blob (uint32 compare)"}
            if rounding == "reference":
                verify["equal_to_reference_arithmetic"] = values_equal_reference
                verify["reference_arithmetic"] = "oracle/ve_numpy.py float32 (bit-exact with the reference run, tests/golden)"
        sync_all()

    # ---- e2e: public API, host buffers, H2D + kernel + D2H inside the timed region --------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    host_out = torch.empty((batch, plan.n_rows, plan.n_cols), dtype=torch.float32).pin_memory()  # reused result buffer
    for _ in range(2):
        jacfun.batched(*host_in, out=host_out)                                     # warm-up (staging allocations)
    # three windows of e2e_steps calls, the fastest counts (a host hiccup inside one 100 ms window - seen once in this
    # round as 0.158 G against 0.25 G on every other box - is a property of the box, not of the path); all are reported
    e2e_windows = []
    for _ in range(3):
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = jacfun.batched(*host_in, out=host_out)
        torch.cuda.synchronize()
        e2e_windows.append(max_over_ranks(time.perf_counter() - t0))
    e2e_s = min(e2e_windows)
    e2e_value = world * batch * e2e_steps / e2e_s
    checksum = float(host_out[:4096].double().abs().sum())                         # touch the result on the host
    # the floor of that call on this box: the same bytes as raw pinned copies (H2D of the inputs on one stream, D2H
    # of the result on another), all ranks at the same time, no kernel
    sync_all()
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    d_in = [torch.empty_like(t, device=dev) for t in host_in]
    d_out = sets[0][1]

    def raw_copies():
        with torch.cuda.stream(s_in):
            for h, d in zip(host_in, d_in):
                d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s_out):
            host_out.copy_(d_out, non_blocking=True)
    for _ in range(2):
        raw_copies()
    floor_windows = []
    for _ in range(3):
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            raw_copies()
        torch.cuda.synchronize()
        floor_windows.append(max_over_ranks(time.perf_counter() - t0))
    floor_s = min(floor_windows)
    floor_value = world * batch * e2e_steps / floor_s
    copy_bytes = 4 * (plan.n_in_scalars + plan.n_rows * plan.n_cols) * batch
    del d_in
    sync_all()

    def time_variant(jf2, k2: int) -> float:
        dp2 = jf2._device_plan(jf2.lower(w.arg_shapes), local_rank)

        def launch(i):
            ins, out_t = sets[i % N_BUFFER_SETS]
            dp2.launch(batch, [t.data_ptr() for t in ins], out_t.data_ptr(), None, stream.cuda_stream)
        return _time_launches(launch, k2, stream, world, dist, dev)

    def summary(ms: float) -> dict:
        gbs = plan.bytes_per_jacobian * batch / (ms * 1e-3) / 1e9
        return {"value": world * batch / (ms * 1e-3), "ms_per_step": ms, "roofline_frac": gbs / peak}

    # both roundings of the headline order, with their full-batch parity records
    k2 = max(20, args.steps // 4)
    modes = {}
    if args.other_orders:
        for r in ("reference", "fma"):
            ms = elapsed_ms / args.steps if r == rounding else \
                time_variant(gx.jacve(w.fun, order=order, argnums=w.argnums, rounding=r), k2)
            modes[r] = dict(summary(ms), parity_full_batch=parity_record(args.order, r),
                            what=("the reference's own sequence of rounded float32 operations (separate multiply and add, "
                                  "IEEE divisions): results equal the reference arithmetic" if r == "reference" else
                                  "product-and-add pairs contracted into FMAs, x/y through one reciprocal: fewer "
                                  "instructions, a few ulp per operation away from the reference"))
    # the other elimination orders of the config ("mixed ... vs fwd/rev"), same batch, fewer steps
    other = {}
    if args.other_orders and WORKLOAD == "roe3d":
        for key in w.order_names():
            if key == args.order:
                continue
            rec = {}
            for r in ("reference", "fma"):
                rec[r] = summary(time_variant(gx.jacve(w.fun, order=w.order(key), argnums=w.argnums, rounding=r), k2))
            other[key] = dict(rec[rounding], steps=k2, other_rounding=rec["fma" if rounding == "reference" else "reference"])
        # "+contract_sums": the opt-in fma variant that also folds edge sums into FMA chains (plan._contract_sums)
        ms = time_variant(gx.jacve(w.fun, order=order, argnums=w.argnums, contract_sums=True), k2)
        other[args.order + "+contract_sums (fma)"] = dict(summary(ms), steps=k2)
    sync_all()

    other_configs = None
    if world == 1 and args.other_configs and WORKLOAD == "roe3d":
        del sets
        torch.cuda.empty_cache()
        other_configs = measure_other_configs(local_rank, peak)

    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        value = world * batch / (ms_per_step * 1e-3)
        achieved = plan.bytes_per_jacobian * batch / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{TITLES[WORKLOAD]} ({plan.stats['n_vertices']} vertices) jacve order={args.order}, "
                                   f"vmap batch {batch} per GPU, fp32, argnums all {len(w.argnums)} inputs -> "
                                   f"[{plan.n_rows}x{plan.n_cols}] Jacobian",
                       "batch_per_gpu": batch, "order": args.order, "rounding": rounding,
                       "kernel": runtime.KERNEL_NAMES[info.kernel_kind],
                       "plan_ops": plan.stats["n_ops"], "threads_per_cta": info.threads,
                       "ctas_per_sm": info.ctas_per_sm, "regs_per_thread": info.regs_per_thread,
                       "cache": f"{N_BUFFER_SETS} rotating input/output sets "
                                f"({N_BUFFER_SETS * plan.bytes_per_jacobian * batch / 1e6:.0f} MB) larger than the 126 MB L2",
                       "launch": (f"CUDA graph of {per_graph} steps per replay" if graph is not None else "one host call per step"),
                       "parallelism": f"batch sharded over {world} GPU(s), no collective",
                       "host_affinity": numa},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * plan.n_in_scalars * batch,
                    "d2h_bytes_per_step": 4 * plan.n_rows * plan.n_cols * batch, "steps": e2e_steps,
                    "windows": {"n": 3, "what": "fastest of three windows of `steps` calls",
                                "values": [world * batch * e2e_steps / t for t in e2e_windows]},
                    "api": f"graphax_b200.vmap(jacve({TITLES[WORKLOAD]}, order, argnums))(*pinned_host_arrays, out=pinned_result)",
                    "checksum": checksum,
                    "floor": {"value": floor_value, "unit": UNIT, "ms_per_step": 1e3 * floor_s / e2e_steps,
                              "gbs_per_gpu": copy_bytes * e2e_steps / floor_s / 1e9,
                              "gbs_all_gpus": world * copy_bytes * e2e_steps / floor_s / 1e9,
                              "what": f"raw pinned H2D of the inputs + D2H of the result ({copy_bytes / 1e6:.0f} MB per GPU and "
                                      f"step) on two streams, {world} rank(s) concurrently, no kernel"},
                    "frac_of_floor": e2e_value / floor_value},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": recorded_traffic(f"{args.order}:{rounding}") or
                         (recorded_traffic(args.order) if rounding == "fma" else None), "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "fp32_issue": recorded_traffic(f"{args.order}:{rounding}:fp32"),   # the co-bound
                         "algorithmic_bytes_per_jacobian": plan.bytes_per_jacobian,
                         "kernel_ms": ms_per_step},
            "clocks": sampler.summary(),
        }
        if verify is not None:
            line["verify"] = verify
        if modes:
            line["rounding_modes"] = modes
        if other:
            line["other_orders"] = other
        if other_configs:
            line["other_configs"] = other_configs
        if world == 1 and not args.no_cpu_baseline:
            cores = len(os.sched_getaffinity(0))
            if _reference_available():
                sample = 1 << 20
                v, _ = cpu_reference_throughput(args.order, sample, 5, cores)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "sample": f"{sample} samples x 5 reps, fp32, {cores} processes: graphax's own sources "
                                                  f"(baseline/_ref) as jax.vmap(graphax.jacve(f, order, argnums)) on the NumPy "
                                                  f"stand-in for JAX (tests/golden/jaxlite)"}
            else:
                sample = 1 << 19
                v, _ = cpu_port_throughput(args.order, sample, 3, cores)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"{sample} samples x 3 reps, NumPy port of the reference algorithm, "
                                                  f"fp32, {cores} processes"}
            line["cpu_baseline"]["numpy_port_value"] = cpu_port_throughput(args.order, 1 << 18, 3, cores)[0]
            line["cpu_baseline"]["c_replay_value"] = cpu_replay_throughput(args.order, 1 << 18, 3, cores)
            # generic autodiff on the same GPU (stand-in for the reference's XLA-on-one-B200 run, JAX
            # not being installed): torch.func.vmap(jacrev(f)), eager.  Reported for context only.
            try:
                sys.path.insert(0, str(ROOT / "tools"))
                from torch_baseline import measure as torch_measure
                sets = None
                torch.cuda.empty_cache()
                tb = min(batch, 1 << 18)
                tv, _ = torch_measure(w, tb, reps=3, device="cuda")
                tc, _ = torch_measure(w, min(batch, 1 << 16), reps=3, device="cpu")
                line["gpu_autodiff_baseline"] = {"value": tv, "unit": UNIT, "kind": "torch.func.vmap(jacrev(f)), eager, fp32",
                                                 "sample": f"{tb} samples x 3 reps on cuda:0",
                                                 "same_on_host_cpu": tc}
            except Exception as e:          # a baseline must never take the bench line down
                line["gpu_autodiff_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="graphax_b200", choices=["graphax_b200", "reference"])
    ap.add_argument("--order", default="mixed")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--kernel", default="auto", choices=["auto", "aot", "jit", "interpreter"])
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the timed steps from a CUDA graph (auto: when a step moves < 64 MB)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--clock-probe-s", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-orders", dest="other_orders", action="store_false")
    ap.add_argument("--workload", default="roe3d", choices=list(WORKLOADS))
    ap.add_argument("--rounding", default="", choices=["", "reference", "fma"],
                    help="plan rounding of the headline (default: plan.DEFAULT_ROUNDING = reference)")
    ap.add_argument("--no-other-configs", dest="other_configs", action="store_false",
                    help="skip the other named configs (RoeFlux_1d, RobotArm_6DOF, MLP) reported at N = 1")
    ap.add_argument("--no-verify", action="store_true", help="skip the all-gather verification after the timed region")
    args = ap.parse_args()
    global WORKLOAD, METRIC
    WORKLOAD = args.workload
    METRIC = f"Jacobians/s, {TITLES[WORKLOAD]} vmap batch"
    if args.order not in WORKLOADS[WORKLOAD].order_names():
        args.order = "rev"
    if args.batch == 0:
        args.batch = WORKLOADS[WORKLOAD].batch
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
