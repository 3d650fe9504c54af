#!/usr/bin/env python
"""bench.py - Jacobians/s of the jacve hot path on the RoeFlux_3d vmap batch (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--order mixed] [--impl reference]

One "step" = one pass of the hot path over one batch of 2^20 synthetic samples per GPU (weak scaling,
no collective on the data path).  `value` is whole-job throughput with inputs resident in HBM, timed
with CUDA events (max over ranks); `e2e` is the same metric through the public API with HOST buffers
(pinned host -> device copies and the device -> host read of the Jacobians inside the timed region);
`roofline` relates the kernel to the measured HBM copy bandwidth over the compulsory 240 B/Jacobian;
`cpu_baseline` / `--impl reference` time the oracle's NumPy port of the reference algorithm on the
host cores (the reference itself needs JAX, which this image does not have).
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
    """--impl reference: the reference's CPU implementation of the path. graphax itself cannot run here
    (it imports jax), so this times the oracle's port of it on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    reference_note = ("graphax needs jax, which is not installed in this image; timing the oracle port (bit-exact with a "
                      "run of the reference's own sources on a NumPy stand-in for JAX, tests/golden/reference_run.*)")
    try:  # honour a per-pod reference install if one ever exists
        sys.path.insert(0, str(ROOT / "baseline" / "_ref"))
        import jax  # noqa: F401
        import graphax  # noqa: F401
        reference_note = "jax+graphax importable but not wired up; timing the oracle port"
    except Exception:
        pass
    finally:
        sys.path.pop(0)
    sample = 1 << 18
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    value, times = cpu_port_throughput(args.order, sample, steps + max(0, min(args.warmup, 2)), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * float(np.median(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{TITLES[WORKLOAD]} order={args.order}, NumPy port of the reference algorithm, "
                               f"{sample} samples per step (bounded sample of the 2^20 batch)",
                   "note": reference_note},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
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


def run_gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist
    import graphax_b200 as gx
    from graphax_b200 import runtime
    from graphax_b200.sharding import bind_to_local_numa_node, init_process_group

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
    jacfun = gx.jacve(w.fun, order=order, argnums=w.argnums)
    plan = jacfun.lower(w.arg_shapes)
    dp = jacfun._device_plan(plan, local_rank)
    if args.kernel != "auto":
        dp.select_kernel({"interpreter": 0, "aot": 1, "jit": 2}[args.kernel])
    info = dp.info()

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
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    sync_all()

    # ---- e2e: public API, host buffers, H2D + kernel + D2H inside the timed region --------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    host_out = torch.empty((batch, plan.n_rows, plan.n_cols), dtype=torch.float32).pin_memory()  # reused result buffer
    for _ in range(2):
        jacfun.batched(*host_in, out=host_out)                                     # warm-up (staging allocations)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = jacfun.batched(*host_in, out=host_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * batch * e2e_steps / e2e_s
    checksum = float(host_out[:4096].double().abs().sum())                         # touch the result on the host

    # the other elimination orders of the config ("mixed ... vs fwd/rev"), same batch, fewer steps
    other = {}
    if args.other_orders and WORKLOAD == "roe3d":
        # "+contract_sums": the opt-in plan variant that folds edge sums into FMA chains (plan._contract_sums);
        # compiled at run time with NVRTC, reported for information, never the headline
        for key, contract in [(k, c) for c in (False, True) for k in w.order_names()]:
            if key == args.order and not contract:
                continue
            jf2 = gx.jacve(w.fun, order=w.order(key), argnums=w.argnums, contract_sums=contract)
            dp2 = jf2._device_plan(jf2.lower(w.arg_shapes), local_rank)
            k2 = max(20, args.steps // 4)
            for i in range(5):
                ins, out_t = sets[i % N_BUFFER_SETS]
                dp2.launch(batch, [t.data_ptr() for t in ins], out_t.data_ptr(), None, stream.cuda_stream)
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(k2):
                ins, out_t = sets[i % N_BUFFER_SETS]
                dp2.launch(batch, [t.data_ptr() for t in ins], out_t.data_ptr(), None, stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / k2
            if world > 1:
                t = torch.tensor([ms2], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms2 = float(t.item())
            other[key + ("+contract_sums" if contract else "")] = {"value": world * batch / (ms2 * 1e-3), "ms_per_step": ms2, "steps": k2,
                          "roofline_frac": plan.bytes_per_jacobian * batch / (ms2 * 1e-3) / 1e9 / measured_peak_gbs()[0]}
            sync_all()

    if rank == 0:
        ms_per_step = elapsed_ms / args.steps
        value = world * batch / (ms_per_step * 1e-3)
        peak, peak_src = measured_peak_gbs()
        achieved = plan.bytes_per_jacobian * batch / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{TITLES[WORKLOAD]} ({plan.stats['n_vertices']} vertices) jacve order={args.order}, "
                                   f"vmap batch {batch} per GPU, fp32, argnums all {len(w.argnums)} inputs -> "
                                   f"[{plan.n_rows}x{plan.n_cols}] Jacobian",
                       "batch_per_gpu": batch, "order": args.order, "kernel": runtime.KERNEL_NAMES[info.kernel_kind],
                       "plan_ops": plan.stats["n_ops"], "threads_per_cta": info.threads,
                       "ctas_per_sm": info.ctas_per_sm, "regs_per_thread": info.regs_per_thread,
                       "cache": f"{N_BUFFER_SETS} rotating input/output sets "
                                f"({N_BUFFER_SETS * plan.bytes_per_jacobian * batch / 1e6:.0f} MB) larger than the 126 MB L2",
                       "launch": (f"CUDA graph of {per_graph} steps per replay" if graph is not None else "one host call per step"),
                       "parallelism": f"batch sharded over {world} GPU(s), no collective",
                       "host_affinity": numa},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * plan.n_in_scalars * batch,
                    "d2h_bytes_per_step": 4 * plan.n_rows * plan.n_cols * batch, "steps": e2e_steps,
                    "api": f"graphax_b200.vmap(jacve({TITLES[WORKLOAD]}, order, argnums))(*pinned_host_arrays, out=pinned_result)",
                    "checksum": checksum},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": recorded_traffic(args.order), "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "fp32_issue": recorded_traffic(args.order + ":fp32"),   # the co-bound
                         "algorithmic_bytes_per_jacobian": plan.bytes_per_jacobian,
                         "kernel_ms": ms_per_step},
            "clocks": sampler.summary(),
        }
        if other:
            line["other_orders"] = other
        if world == 1 and not args.no_cpu_baseline:
            cores = len(os.sched_getaffinity(0))
            sample = 1 << 19
            v, _ = cpu_port_throughput(args.order, sample, 3, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} samples x 3 reps, NumPy port of the reference algorithm, "
                                              f"fp32, {cores} processes",
                                    "c_replay_value": cpu_replay_throughput(args.order, 1 << 18, 3, cores)}
            # generic autodiff on the same GPU (stand-in for the reference's XLA-on-one-B200 run, JAX
            # not being installed): torch.func.vmap(jacrev(f)), eager.  Reported for context only.
            try:
                sys.path.insert(0, str(ROOT / "tools"))
                from torch_baseline import measure as torch_measure
                del sets
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
