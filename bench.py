#!/usr/bin/env python3
"""Headline benchmark: SSA reaction events/s for the Vilar oscillator ensemble.

Workload (BASELINE.json configs[3]): Vilar circadian oscillator (9 species, 16 reactions,
benchmarks/benches/vilar/vilar.rs:6-46 of the reference), t = 0..200 sampled every 1 time unit
(201 x 9 samples per trajectory), 10^7 trajectories sharded over 8 GPUs = 1.25e6 per GPU, weak
scaling (every GPU always simulates 1.25e6 trajectories; trajectory n uses seed n).

A "step" is one pass of the hot path over one shard: every trajectory of the shard from t = 0 to
t = 200 in ONE kernel launch, samples left in HBM as int32 [step][species][trajectory].

  value  device-resident: seeds derived on the device (seed = base + n), samples stay in HBM.
  e2e    the C-ABI call a host makes, with HOST buffers: per step the per-trajectory seeds and x0 are
         copied from pinned host memory, the kernel runs, and the full sample tensor is copied back
         into a pinned host buffer, all inside the timed region.
  roofline  the kernel is FP64/INT issue bound, not HBM bound: `achieved` = events x F / kernel time
         with F = O + 2R + 9 = 60 non-fused FP64 operations per event, `peak` = the
         non-fused FP64 issue rate measured on this GPU in this run; the `hbm` sub-object reports
         the sample stores against MEASURED_PEAKS.json.
  cpu_baseline  the oracle's define_system!-style straight-line Vilar code on all host cores, on a
         bounded sample of the same workload (N=1, rank 0 only).

  parity_check  inside the same run: the first 256 trajectories of the last end-to-end step (host buffer)
         are recomputed by the oracle from the same seeds and compared bit for bit; a mismatch makes the
         run exit non-zero, so a printed number always belongs to a checked result.
  configs  the other BASELINE.json configurations, each in its stated form, as short runs (1 warm-up + 3
         steps): C1 SIR through the function API arithmetic, C2 Dimers in define_system! arithmetic, C3
         Michaelis-Menten with an expression rate through the Python `Gillespie.run` (trajectories/s, full
         samples and reduce=True), C5 the synthetic 100 x 500 network, 10^6 trajectories sharded over the GPUs.

`--impl reference` times the reference's CPU algorithm (the oracle port; the Rust crate cannot be
compiled in this image) on the host cores and prints the same JSON line shape.  It does not import the
product package (no CUDA library is mapped into that process).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402



def load_models():
    """rebop_b200/models.py as a stand-alone module: plain data, no import of the package (whose __init__ loads the
    CUDA library)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("rebop_b200_models_data", os.path.join(ROOT, "rebop_b200", "models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


METRIC = "ssa_reaction_events_per_sec"
UNIT = "events/s"
TRAJ_PER_GPU = 1_250_000  # 10^7 over 8 GPUs


def fp64_ops_per_event(model):
    """F = O + 2R + 9 non-fused FP64 operations per event: O = total reactant order,
    R = reactions, 9 = guard, Exp1 fast path (3), divide, t +=, overshoot test, uniform scale, total * u.
    Vilar 60, SIR 16, Dimers 25."""
    order = sum(e for _, terms, _ in model["reactions"] for _, e in terms)
    return order + 2 * len(model["reactions"]) + 9



def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def workload_config(args, n_gpus):
    return {
        "workload": f"{args.model} ensemble, define_system! arithmetic, tmax={args.tmax:g}, nb_steps={args.nb_steps}, "
                    f"{args.traj_per_gpu} trajectories per GPU ({args.traj_per_gpu * n_gpus} total), all species saved "
                    f"(--impl reference times a bounded sample of this workload per step on the host cores: its size is in "
                    f"`sample` and `cpu_baseline.sample` of that line)",
        "trajectories_per_gpu": args.traj_per_gpu,
        "trajectories_total": args.traj_per_gpu * n_gpus,
        "tmax": args.tmax,
        "nb_steps": args.nb_steps,
        "sharding": f"trajectories x{n_gpus}, no data-path collective",
        "l2": "sample tensor written per step exceeds L2 (%.2f GB); inputs are seeds only"
              % (args.traj_per_gpu * (args.nb_steps + 1) * args.n_species * 4 / 1e9),
    }


# --------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm on the host cores
# --------------------------------------------------------------------------------------
def cpu_sample(model, n_traj, tmax, nb_steps, threads, seed_first=0):
    """Oracle (define_system! form) on `threads` host threads -> (events, seconds)."""
    from oracle import oracle as O  # checker / CPU baseline only

    seeds = np.arange(seed_first, seed_first + n_traj, dtype=np.uint64)
    t0 = time.perf_counter()
    try:  # hand-expanded define_system! code of the benchmark systems (what the macro compiles to on the CPU)
        _, _, events = O.run_batch_macro(model["name"], model["params"], model["x0"], seeds, tmax, nb_steps,
                                         threads=threads, want_out=True, want_events=False)
    except KeyError:  # any other network: the function-API engine in its sparse form
        net = O.Network(len(model["species"]), [("lma", k, terms, diff) for k, terms, diff in model["reactions"]], arith=1)
        _, _, events = net.run_batch(model["x0"], seeds, tmax, nb_steps, threads=threads, want_out=True, want_events=False)
    return events, time.perf_counter() - t0


def run_reference(args, model):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    per_step = max(cores, args.ref_traj_per_core * cores)
    for i in range(args.warmup):
        cpu_sample(model, per_step, args.tmax, args.nb_steps, cores, seed_first=i * per_step)
    events = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        ev, _ = cpu_sample(model, per_step, args.tmax, args.nb_steps, cores, seed_first=(args.warmup + i) * per_step)
        events += ev
    dt = time.perf_counter() - t0
    value = events / dt
    sample = f"{per_step} trajectories per step ({args.ref_traj_per_core} per core) of the same workload"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "sample": sample, "trajectories_per_step": per_step,
        "trajectories_per_s": per_step * args.steps / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle port of the reference's define_system! code path (the Rust crate cannot be built here: no cargo/rustc); "
                "built with gcc -O3 -ffp-contract=off, without -march=native (the .so travels from the build container to a "
                "different host: oracle/Makefile)",
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled during the timed region.

    NVML is queried directly from a thread of this process (pynvml).  Polling through the nvidia-smi
    binary was measured to perturb short workloads: attaching to the driver takes about a second and every
    poll delays kernel launches by tens of milliseconds (a 110 ms step became 147 ms)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, device, period=0.25):
        self.device = device
        self.period = period
        self.samples = []
        self.thread = None
        self.stop_flag = threading.Event()
        self.error = None
        self.skip = 0

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.device < len(ids) and ids[self.device].isdigit():
                return int(ids[self.device])
        return self.device

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001 - NVML missing or refusing: the run goes on without clock samples
            self.error = f"NVML unavailable: {e}"
            return

        def loop():
            while not self.stop_flag.is_set():
                try:
                    sm = float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))
                    power = pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0
                    try:
                        reasons = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle))
                    except AttributeError:
                        reasons = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
                    self.samples.append((sm, power, reasons))
                except Exception as e:  # noqa: BLE001
                    self.error = f"NVML query failed: {e}"
                    return
                self.stop_flag.wait(self.period)

        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def mark(self):
        """Samples taken so far (warm-up) are not part of the report."""
        self.skip = len(self.samples)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.error or "sampler not started"]}
        self.stop_flag.set()
        self.thread.join(timeout=5)
        samples = self.samples[self.skip:] or self.samples[-1:]
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.error or "no samples"]}
        seen = 0
        for _, _, r in samples:
            seen |= r
        return {"sm_mhz": float(np.median([s[0] for s in samples])), "sm_max_mhz": self.sm_max,
                "power_w_max": max(s[1] for s in samples), "samples": len(samples),
                "reasons": sorted(name for name, bit in self.REASONS.items() if seen & bit)}


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def oracle_samples(model, arith, seeds, tmax, nb_steps, threads, save_idx=None):
    """The checker: oracle samples [nb_steps+1][n_save][len(seeds)] for a model dict in the given arithmetic."""
    from oracle import oracle as O  # checker only

    if arith == 1 and save_idx is None:
        try:
            return O.run_batch_macro(model["name"], model["params"], model["x0"], seeds, tmax, nb_steps, threads=threads,
                                     want_events=False)[0]
        except KeyError:
            pass
    net = O.Network(len(model["species"]), [("lma", k, terms, diff) for k, terms, diff in model["reactions"]], arith=arith)
    return net.run_batch(model["x0"], seeds, tmax, nb_steps, save_idx=save_idx, threads=threads, want_events=False)[0]


def oracle_parity(model, arith, seeds, tmax, nb_steps, got, threads, save_idx=None):
    """Compare `got` [nb_steps+1][n_save][len(seeds)] with the oracle on the same seeds, bit for bit."""
    t0 = time.perf_counter()
    ref = oracle_samples(model, arith, seeds, tmax, nb_steps, threads, save_idx)
    equal = bool(np.array_equal(np.asarray(got).astype(np.int64), ref.astype(np.int64)))
    return {"n": int(len(seeds)), "equal": equal, "rows": int(ref.shape[0] * ref.shape[1]),
            "oracle_seconds": round(time.perf_counter() - t0, 3),
            "what": "samples of the first trajectories of the last timed end-to-end step vs the CPU oracle on the same seeds"}


class _HostBuffer:
    """Page-locked host array when the box grants it, pageable otherwise (stated in the line)."""

    def __init__(self, ffi, shape, dtype):
        self.kind = "pinned"
        try:
            self._pin = ffi.PinnedBuffer(shape, dtype)
            self.array = self._pin.array
        except ffi.RebopError:
            self._pin = None
            self.array = np.empty(shape, dtype=dtype)
            self.kind = "pageable (page-locking the buffer failed)"

    def close(self):
        self.array = None
        if self._pin is not None:
            self._pin.close()


def bench_config(tag, model, arith, n, sample_dtype, ctx, steps, warmup, save_idx=None, parity_n=64, scaling="weak",
                 fast_kernel=False):
    """One BASELINE configuration as a short run on this rank's GPU: device-resident and end-to-end through the C ABI
    (host buffers in the timed region), FP64-issue roofline fraction, lane efficiency and an oracle check of the first
    trajectories of the last end-to-end step.  ctx: rank/world/local, reducers, ffi/models modules, fp64 peak."""
    _ffi, models = ctx["ffi"], ctx["models"]
    rank, world, local = ctx["rank"], ctx["world"], ctx["local"]
    S = len(model["species"])
    n_save = S if save_idx is None else len(save_idx)
    nb, tmax = model["nb_steps"], model["tmax"]
    net = models.build_network(model, arith)
    dtype = np.dtype(sample_dtype)

    def base(i):
        return (i * world + rank) * n

    b = _ffi.Batch(net, n, model["x0"], seeds=None, seed_base=base(0), device=local)
    b.set_sample_dtype(dtype)
    x0 = np.asarray(model["x0"], dtype=np.int64)

    def dev_step(i):
        b.set_species(x0)
        b.set_time(0.0)
        b.seed(None, base(i))
        b.run_grid(tmax, nb, save_idx=save_idx)
        return b.events()[1], b.last_kernel_ms, b.last_finish_ms, b.lane_slots

    for i in range(warmup):
        dev_step(i)
    ctx["barrier"]()
    events = loop_ms = fin_ms = slots = 0
    w0 = time.perf_counter()
    for i in range(steps):
        ev, ms, fm, sl = dev_step(warmup + i)
        events += ev; loop_ms += ms; fin_ms += fm; slots += sl
    ctx["barrier"]()
    wall_dev = ctx["max"](time.perf_counter() - w0)
    dev_ms = ctx["max"](loop_ms + fin_ms)  # device time of the kernels (ensemble loop + sample finishing), max over ranks
    tot_events = ctx["sum"](float(events))
    kernel_used, schedule_used = b.kernel_used, b.schedule_used

    # end to end: seeds and x0 from host memory, samples into a host buffer, all inside the timed region
    host_seeds = _HostBuffer(_ffi, (n,), np.uint64)
    host_out = _HostBuffer(_ffi, (nb + 1, n_save, n), dtype)

    def e2e_step(i):
        host_seeds.array[:] = np.arange(base(i), base(i) + n, dtype=np.uint64)
        b.set_species(x0)
        b.set_time(0.0)
        b.seed(host_seeds.array)
        b.run_grid(tmax, nb, save_idx=save_idx, host_out=host_out.array)
        return b.events()[1]

    for i in range(warmup):
        e2e_step(i)
    ctx["barrier"]()
    w0 = time.perf_counter()
    e2e_events = 0
    for i in range(steps):
        e2e_events += e2e_step(warmup + i)
    ctx["barrier"]()
    e2e_s = ctx["max"](time.perf_counter() - w0)
    e2e_total = ctx["sum"](float(e2e_events))
    last = warmup + steps - 1
    n_chk = min(parity_n, n)
    parity = oracle_parity(model, arith, np.arange(base(last), base(last) + n_chk, dtype=np.uint64), tmax, nb,
                           host_out.array[:, :, :n_chk], host_cores(), save_idx)
    parity["equal"] = bool(ctx["min"](1.0 if parity["equal"] else 0.0) == 1.0)
    host_mem = host_out.kind
    # ---- optional: the opt-in partial-propensity kernel on the same workload (tier-2 parity: it is the direct method
    # with differently ordered floating-point sums, so it is checked here through the ensemble mean of the last sample
    # row against the bit-exact run above -- different seeds would do as well -- and through tests/test_pdm.py)
    fast = None
    if fast_kernel:
        exact_last = host_out.array[-1].astype(np.float64)        # last e2e step of the bit-exact kernel, seeds base(last)
        fb = _ffi.Batch(net, n, model["x0"], seeds=None, seed_base=base(0), device=local, kernel=_ffi.KERNEL_PDM)
        fb.set_sample_dtype(dtype)
        fev = fms = 0
        for i in range(warmup + steps):
            fb.set_species(x0)
            fb.set_time(0.0)
            fb.seed(None, base(10**6 + i))                        # seeds the bit-exact runs did not use
            fb.run_grid(tmax, nb, save_idx=save_idx)
            if i >= warmup:
                fev += fb.events()[1]
                fms += fb.last_kernel_ms + fb.last_finish_ms
        fast_last = fb.samples()[-1].astype(np.float64)
        fb.close()
        se = np.sqrt(exact_last.var(axis=1, ddof=1) / n + fast_last.var(axis=1, ddof=1) / n)
        zmax = float(np.max(np.abs(exact_last.mean(axis=1) - fast_last.mean(axis=1)) / np.maximum(se, 1e-300)))
        fast = {"kernel": "REBOP_KERNEL_PDM (partial propensities; statistically exact, opt-in)",
                "value": ctx["sum"](float(fev)) / (ctx["max"](fms) * 1e-3), "unit": UNIT, "ms_per_step": fms / steps,
                "speedup_vs_bit_exact": None,
                "check": {"what": "ensemble means of the last sample row vs the bit-exact kernel (independent seeds), in standard errors",
                          "max_z": zmax, "ok": bool(ctx["min"](1.0 if zmax < 5.0 else 0.0) == 1.0)}}
    host_out.close()
    host_seeds.close()
    b.close()
    F = fp64_ops_per_event(model)
    achieved = events * F / (loop_ms * 1e-3)
    return {
        "workload": f"{tag}: {model['name']}, {'define_system!' if arith == 1 else 'function-API'} arithmetic, tmax={tmax:g}, "
                    f"nb_steps={nb}, {n} trajectories per GPU x {world} GPU(s), {n_save} of {S} species saved as {dtype.name}",
        "value": tot_events / (dev_ms * 1e-3), "unit": UNIT, "scaling": scaling,
        "ms_per_step": dev_ms / steps, "loop_ms_per_step": loop_ms / steps, "finish_ms_per_step": fin_ms / steps,
        "wall_ms_per_step": wall_dev / steps * 1e3,
        "trajectories_per_s": n * world * steps / (dev_ms * 1e-3),
        "e2e": {"value": e2e_total / e2e_s, "unit": UNIT, "ms_per_step": e2e_s / steps * 1e3,
                "trajectories_per_s": n * world * steps / e2e_s, "h2d_bytes_per_step": int(n * 8 + S * 8),
                "d2h_bytes_per_step": int((nb + 1) * n_save * n * dtype.itemsize), "host_memory": host_mem},
        "frac": achieved / ctx["fp64_peak"], "ops_per_event": F, "lane_efficiency": events / slots if slots else None,
        "events_per_trajectory": events / (n * steps), "kernel": kernel_used, "schedule": schedule_used, "steps": steps,
        "warmup": warmup, "parity_check": parity,
        **({"fast_kernel": dict(fast, speedup_vs_bit_exact=fast["value"] / (tot_events / (dev_ms * 1e-3)))} if fast else {}),
    }


def bench_python_mm(n, ctx, steps, warmup):
    """BASELINE config 3: examples/mm.py of the reference (expression rate V * A / (Km + A)) as a batched
    `Gillespie.run` through the Python binding: trajectories/s end to end, with the samples returned to the caller
    (int16, the narrowest type the binding offers; counts are <= 100) and with reduce=True (mean/variance only)."""
    import rebop_b200

    rank, world, local = ctx["rank"], ctx["world"], ctx["local"]
    mm = rebop_b200.Gillespie()
    mm.add_reaction("V * A / (Km + A)", ["A"], ["P"])
    kw = dict(tmax=250, nb_steps=100, params={"V": 1, "Km": 20}, n_trajectories=n, device=local)

    def timed(**extra):
        for i in range(warmup):
            mm.run({"A": 100}, rng=1000 * rank + i, **kw, **extra)
        ctx["barrier"]()
        t0 = time.perf_counter()
        events = kernel_ms = 0
        ds = None
        for i in range(steps):
            ds = mm.run({"A": 100}, rng=1000 * rank + warmup + i, **kw, **extra)
            events += mm.last_events
            kernel_ms += mm.last_kernel_ms
        ctx["barrier"]()
        dt = ctx["max"](time.perf_counter() - t0)
        return ds, dt, ctx["sum"](float(events)), kernel_ms

    ds, dt, events, kernel_ms = timed(dtype=np.int16)
    # parity of the last returned samples: the same seeds through the oracle's expression engine
    from oracle import oracle as O  # checker only

    n_chk = min(64, n)
    seeds = np.random.default_rng(1000 * rank + warmup + steps - 1).integers(np.iinfo(np.uint64).max, size=n, dtype=np.uint64)[:n_chk]
    # V * A / (Km + A) with V = 1, Km = 20, post-order as Expr::eval walks it (src/expr.rs:24-38)
    prog = [("const", 0, 1.0), ("species", 0, 0), ("mul", 0, 0), ("const", 0, 20.0), ("species", 0, 0), ("add", 0, 0),
            ("div", 0, 0)]
    try:
        ref = O.Network(2, [("expr", prog, [-1, 1])]).run_batch([100, 0], seeds, 250.0, 100, threads=host_cores(),
                                                               want_events=False)[0]
        got = np.stack([np.asarray(ds["A"])[:, :n_chk], np.asarray(ds["P"])[:, :n_chk]], axis=1)
        equal = bool(np.array_equal(got.astype(np.int64), ref.astype(np.int64)))
    except Exception as e:  # noqa: BLE001 - a failing checker is reported, it does not hide the measurement
        equal = f"checker error: {e}"
    if isinstance(equal, bool):
        equal = bool(ctx["min"](1.0 if equal else 0.0) == 1.0)
    _, dt_r, _, kernel_ms_r = timed(reduce=True)
    return {
        "workload": f"C3: Michaelis-Menten A -> P at rate 'V * A / (Km + A)' (examples/mm.py), rebop_b200.Gillespie.run("
                    f"n_trajectories={n}) per GPU x {world} GPU(s), tmax=250, nb_steps=100, both species returned as int16",
        "value": n * world * steps / dt, "unit": "trajectories/s", "events_per_s": events / dt,
        "ms_per_call": dt / steps * 1e3, "kernel_ms_per_call": kernel_ms / steps,
        "d2h_bytes_per_call": int(101 * 2 * n * 2),
        "reduce": {"value": n * world * steps / dt_r, "unit": "trajectories/s", "ms_per_call": dt_r / steps * 1e3,
                   "kernel_ms_per_call": kernel_ms_r / steps,
                   "what": "reduce=True: ensemble mean and variance per sample time, samples never leave the GPU"},
        "steps": steps, "warmup": warmup, "parity_check": {"n": n_chk, "equal": equal},
    }


def run_gpu(args, model):
    import torch
    import torch.distributed as dist

    from rebop_b200 import _ffi, models

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if _ffi.device_count() <= local:
        raise SystemExit("bench.py needs a CUDA device per rank (rebop_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def min_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    n = args.traj_per_gpu
    S = len(model["species"])
    nb = args.nb_steps
    arith = _ffi.ARITH_MACRO  # BASELINE config 4: the Vilar oscillator through define_system!
    net = models.build_network(model, arith)
    kernel = {"auto": _ffi.KERNEL_AUTO, "table": _ffi.KERNEL_TABLE, "nvrtc": _ffi.KERNEL_NVRTC,
              "prebuilt": _ffi.KERNEL_PREBUILT}[args.kernel]

    def shard_base(step_index):  # distinct trajectories every step and every rank
        return (step_index * world + rank) * n

    batch = _ffi.Batch(net, n, model["x0"], seeds=None, seed_base=shard_base(0), device=local, kernel=kernel)
    stream = torch.cuda.ExternalStream(batch.stream, device=torch.device("cuda", local))

    def device_step(i):
        batch.set_species(model["x0"])
        batch.set_time(0.0)
        batch.seed(None, shard_base(i))
        batch.run_grid(args.tmax, nb)
        return batch.events()[1], batch.last_kernel_ms

    lane_slots = 0

    # ---- value: device-resident -------------------------------------------------------
    # the clock sampler (nvidia-smi polling) starts before the warm-up so that its start-up cost -- it
    # briefly contends for the driver -- is not inside the timed region; it keeps sampling through it
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(args.warmup):
        device_step(i)
    launches0 = _ffi.kernel_launches()
    barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    events = 0
    kernel_ms = 0.0
    w0 = time.perf_counter()
    for i in range(args.steps):
        ev, ms = device_step(args.warmup + i)
        events += ev
        kernel_ms += ms
        lane_slots += batch.lane_slots
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - w0
    clocks = sampler.stop()
    launches = _ffi.kernel_launches() - launches0
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    total_events = sum_over_ranks(float(events))
    value = total_events / (dev_ms * 1e-3)
    kernel_used = batch.kernel_used

    # ---- ensemble statistics: K4 sums, all-reduced over the GPUs with NCCL (north_star's only collective)
    stats_ms = None
    t0 = time.perf_counter()
    from rebop_b200 import ensemble
    sums, rows = ensemble.device_sums_as_tensor(batch, local)  # K4 on this GPU's shard
    ensemble.allreduce_sums(sums)                               # NCCL, int64 sum (no-op at N=1)
    sums_host = sums.cpu().numpy()
    stats_ms = (time.perf_counter() - t0) * 1e3
    n_total = n * world
    mean_last = (sums_host[:rows].reshape(nb + 1, S)[-1] / n_total).tolist()

    # ---- e2e: host buffers through the C ABI ---------------------------------------------
    e2e = None
    parity = None
    if not args.no_e2e:
        host_seeds = _ffi.PinnedBuffer((n,), np.uint64)
        host_memory = "pinned"
        try:
            host_out = _ffi.PinnedBuffer((nb + 1, S, n), np.int32)
        except _ffi.RebopError:  # the box refuses to page-lock that much per rank: pageable host memory instead

            class _Pageable:
                def __init__(self, shape):
                    self.array = np.empty(shape, dtype=np.int32)

                def close(self):
                    self.array = None

            host_out = _Pageable((nb + 1, S, n))
            host_memory = "pageable (page-locking the result buffer failed)"
        x0 = np.asarray(model["x0"], dtype=np.int64)

        e2e_kernel_ms = []

        def e2e_step(i):
            host_seeds.array[:] = np.arange(shard_base(i), shard_base(i) + n, dtype=np.uint64)
            batch.set_species(x0)
            batch.set_time(0.0)
            batch.seed(host_seeds.array)
            batch.run_grid(args.tmax, nb, host_out=host_out.array)
            e2e_kernel_ms.append(batch.last_kernel_ms)
            return batch.events()[1]

        for i in range(args.warmup):
            e2e_step(i)
        barrier()
        w0 = time.perf_counter()
        e2e_events = 0
        for i in range(args.steps):
            e2e_events += e2e_step(args.warmup + i)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - w0)
        e2e_total = sum_over_ranks(float(e2e_events))
        checksum = int(host_out.array[-1].astype(np.int64).sum())
        # ---- parity of the very result that was timed: the oracle recomputes the first trajectories of the last step
        n_chk = min(args.parity_n, n)
        if n_chk:
            last = args.warmup + args.steps - 1
            chk_seeds = np.arange(shard_base(last), shard_base(last) + n_chk, dtype=np.uint64)
            parity = oracle_parity(model, arith, chk_seeds, args.tmax, nb, host_out.array[:, :, :n_chk], host_cores())
            parity["equal"] = bool(min_over_ranks(1.0 if parity["equal"] else 0.0) == 1.0)  # every rank checks its own shard
        e2e = {"value": e2e_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(n * 8 + S * 4),
               "d2h_bytes_per_step": int((nb + 1) * S * n * 4), "ms_per_step": e2e_s / args.steps * 1e3,
               "kernel_ms_per_step": float(np.mean(e2e_kernel_ms[-args.steps:])),
               "trajectories_per_s": n * world * args.steps / e2e_s, "last_row_checksum": checksum,
               "api": "rebop_batch_seed + rebop_batch_run_grid(host_out) with host buffers", "host_memory": host_memory}
        host_out.close()
        host_seeds.close()

    # ---- roofline of the dominant kernel (this rank) ---------------------------------------
    fp64_peak, fp64_mhz = _ffi.measure_fp64_rate(local)
    F = fp64_ops_per_event(model)
    ms_per_launch = kernel_ms / args.steps
    ev_per_launch = events / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    sample_bytes = (nb + 1) * S * n * 4 + n * (S * 4 + 8 + 32) * 2
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tr.get("model") == model["name"]:  # measured per (trajectory, sample row); scaled to this launch
            traffic = tr["dram_bytes_per_trajectory_row"] * n * (nb + 1)
    except (OSError, KeyError, ValueError):
        pass
    achieved = ev_per_launch * F / (ms_per_launch * 1e-3) / 1e9 if F else None
    roofline = {
        "kernel": {_ffi.KERNEL_TABLE: "rb_ssa_table_kernel (K1, table-driven)", _ffi.KERNEL_NVRTC: "rb_ssa_jit (K2, NVRTC)",
                   _ffi.KERNEL_PREBUILT: "rb_ssa_sys_* (K2, built by rebop_sysgen + nvcc)"}.get(kernel_used, str(kernel_used)),
        "bound": "fp64_issue", "achieved": achieved, "peak": fp64_peak / 1e9, "unit": "GFLOP/s (non-fused f64 ops)",
        "frac": achieved / (fp64_peak / 1e9) if achieved else None,
        "peak_source": "measured in this run (rebop_b200_measure_fp64_rate): 8 independent non-fused DADD/DMUL chains per "
                       "thread on all SMs; MEASURED_PEAKS.json has no FP64 entry",
        "peak_sm_mhz": fp64_mhz, "peak_lanes_per_sm_clk": fp64_peak / (fp64_mhz * 1e6) / torch.cuda.get_device_properties(local).multi_processor_count
        if fp64_mhz else None,
        "ops_per_event": F, "events_per_launch": ev_per_launch, "ms_per_launch": ms_per_launch,
        "lane_efficiency": events / lane_slots if lane_slots else None,
        "traffic": traffic,
        "hbm": {"bound": "hbm", "achieved": sample_bytes / (ms_per_launch * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": sample_bytes / (ms_per_launch * 1e-3) / 1e9 / hbm_peak, "bytes_per_launch": sample_bytes,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }

    # ---- CPU baseline (rank 0, N = 1) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        n_cpu = max(cores, args.cpu_traj_per_core * cores)
        cpu_sample(model, cores, args.tmax, args.nb_steps, cores)  # warm the threads and caches
        ev, dt = cpu_sample(model, n_cpu, args.tmax, args.nb_steps, cores, seed_first=10**9)
        cpu = {"value": ev / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_cpu} trajectories of the same workload in {dt:.1f} s (oracle, define_system! form, {cores} threads)"}

    batch.close()

    # ---- the other BASELINE configurations, each in its stated form (short runs)
    configs = None
    if not args.no_configs:
        ctx = {"ffi": _ffi, "models": models, "rank": rank, "world": world, "local": local, "barrier": barrier,
               "max": max_over_ranks, "sum": sum_over_ranks, "min": min_over_ranks, "fp64_peak": fp64_peak}
        cs, cw = args.config_steps, args.config_warmup
        n1 = args.config_traj
        configs = {
            # C1: SIR through the Gillespie function API (its arithmetic, its reaction choice), counts <= 1000 -> int16
            "C1_sir_api": bench_config("C1", models.sir(), _ffi.ARITH_API, n1, np.int16, ctx, cs, cw),
            # C2: Dimers as define_system! writes it, final state only (nb_steps = 1)
            "C2_dimers_macro": bench_config("C2", models.dimers(), _ffi.ARITH_MACRO, n1, np.int32, ctx, cs, cw),
            # C3: Python binding, expression rate, 10^7 trajectories (split over the GPUs of the run)
            "C3_mm_python": bench_python_mm(max(1, args.config_traj_mm // world), ctx, cs, cw),
            # C5: synthetic 100 x 500 network, function-API arithmetic, 10^6 trajectories sharded over the GPUs (strong
            # scaling); the first 10 species are returned (all 100 would be 40 GB of int32 per step)
            "C5_synthetic_api": bench_config("C5", models.synthetic(), _ffi.ARITH_API, max(1, n1 // world), np.int32, ctx, cs, cw,
                                             save_idx=list(range(10)), parity_n=16, scaling="strong", fast_kernel=True),
        }
    all_equal = all(v is True for v in [parity and parity["equal"]] +
                    [c["parity_check"]["equal"] for c in (configs or {}).values()] +
                    [c["fast_kernel"]["check"]["ok"] for c in (configs or {}).values() if "fast_kernel" in c] if v is not None)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "trajectories_per_s": n * world * args.steps / (dev_ms * 1e-3),
            "events_per_step": total_events / args.steps,
            "wall_ms_per_step": wall / args.steps * 1e3,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "kernel": roofline["kernel"], "roofline": roofline, "cpu_baseline": cpu,
            "parity_check": parity, "configs": configs,
            "ensemble_stats": {"ms": stats_ms, "collective": "nccl all_reduce(int64 sum)" if world > 1 else "none (1 GPU)",
                               "mean_at_tmax": dict(zip(model["species"], mean_last))},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not all_equal:
        print("bench.py: PARITY CHECK FAILED: the timed result differs from the oracle", file=sys.stderr)
        return 3
    return 0


def run_single_process(args, model):
    """`--single-process`: ONE process drives all N GPUs through the C ABI's ensemble handle (rebop_ensemble_*: one
    batch and one host worker thread per device, contiguous trajectory ranges; the ensemble sums are all-reduced over
    an ncclCommInitAll communicator inside the library).  Same workload and keys as the torchrun mode; the device time
    of a step is the slowest device's kernel time (rebop_ensemble_last_kernel_ms)."""
    from rebop_b200 import _ffi, models

    G = args.gpus
    if _ffi.device_count() < G:
        raise SystemExit(f"--single-process --gpus {G}: only {_ffi.device_count()} CUDA device(s) visible")
    n, nb, S = args.traj_per_gpu * G, args.nb_steps, len(model["species"])
    arith = _ffi.ARITH_MACRO
    net = models.build_network(model, arith)
    ens = _ffi.Ensemble(net, n, model["x0"], list(range(G)), seeds=None, seed_base=0)
    x0 = np.asarray(model["x0"], dtype=np.int64)
    samplers = [ClockSampler(g) for g in range(G)]
    for sm in samplers:
        sm.start()

    def step(i, host_out=None, host_seeds=None):
        ens.set_species(x0)
        ens.set_time(0.0)
        if host_seeds is None:
            ens.seed(None, i * n)
        else:
            host_seeds[:] = np.arange(i * n, (i + 1) * n, dtype=np.uint64)
            ens.seed(host_seeds)
        ens.run_grid(args.tmax, nb, host_out=host_out)
        return ens.events()[1], ens.last_kernel_ms

    for i in range(args.warmup):
        step(i)
    for sm in samplers:
        sm.mark()
    launches0 = _ffi.kernel_launches()
    events = kernel_ms = 0
    w0 = time.perf_counter()
    for i in range(args.steps):
        ev, ms = step(args.warmup + i)
        events += ev
        kernel_ms += ms
    wall = time.perf_counter() - w0
    launches = _ffi.kernel_launches() - launches0
    t0 = time.perf_counter()
    mean, var = ens.stats()          # K4 per device -> ncclAllReduce(int64) -> finalisation kernel
    stats_ms = (time.perf_counter() - t0) * 1e3
    clocks = [sm.stop() for sm in samplers]

    e2e = parity = None
    if not args.no_e2e:
        host_out = np.empty((nb + 1, S, n), dtype=np.int32)
        host_seeds = np.empty(n, dtype=np.uint64)
        for i in range(args.warmup):
            step(i, host_out, host_seeds)
        w1 = time.perf_counter()
        e2e_events = 0
        for i in range(args.steps):
            e2e_events += step(args.warmup + i, host_out, host_seeds)[0]
        e2e_s = time.perf_counter() - w1
        e2e = {"value": e2e_events / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(n * 8 + S * 8),
               "d2h_bytes_per_step": int(host_out.nbytes), "ms_per_step": e2e_s / args.steps * 1e3,
               "host_memory": "pageable", "api": "rebop_ensemble_seed + rebop_ensemble_run_grid(host_out)"}
        last = args.warmup + args.steps - 1
        # every device's first trajectories of the last step against the oracle
        equal, n_chk = True, 0
        for dev, first, count in ens.shards():
            k = min(args.parity_n // G or 1, count)
            chk = oracle_parity(model, arith, np.arange(last * n + first, last * n + first + k, dtype=np.uint64), args.tmax, nb,
                                host_out[:, :, first:first + k], host_cores())
            equal, n_chk = equal and chk["equal"], n_chk + k
        parity = {"n": n_chk, "equal": bool(equal), "what": "first trajectories of every device's shard, last e2e step, vs the oracle"}
    ens.close()
    dev_s = kernel_ms * 1e-3
    line = {
        "metric": METRIC, "value": events / dev_s, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": kernel_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, G),
        "mode": "single process, rebop_ensemble_* (one host thread per device; NCCL all-reduce of the ensemble sums in the library)",
        "wall_ms_per_step": wall / args.steps * 1e3, "value_by_wall_clock": events / wall,
        "trajectories_per_s": n * args.steps / dev_s, "clocks": clocks[0], "clocks_per_device": clocks, "e2e": e2e,
        "gpu_launches": launches, "parity_check": parity,
        "ensemble_stats": {"ms": stats_ms, "collective": "ncclAllReduce(int64 sum) over ncclCommInitAll" if G > 1 else "none (1 GPU)",
                           "mean_at_tmax": dict(zip(model["species"], mean[-1].tolist()))},
    }
    print(json.dumps(line), flush=True)
    if parity is not None and not parity["equal"]:
        print("bench.py: PARITY CHECK FAILED: the timed result differs from the oracle", file=sys.stderr)
        return 3
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="vilar")
    ap.add_argument("--traj-per-gpu", type=int, default=TRAJ_PER_GPU)
    ap.add_argument("--tmax", type=float, default=None)
    ap.add_argument("--nb-steps", type=int, default=None)
    ap.add_argument("--kernel", default="auto", choices=["auto", "table", "nvrtc", "prebuilt"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--single-process", action="store_true",
                    help="one process drives all --gpus devices through the C ABI's ensemble handle (no torchrun)")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE configurations")
    ap.add_argument("--parity-n", type=int, default=256, help="trajectories of the last e2e step recomputed by the oracle")
    ap.add_argument("--config-steps", type=int, default=3)
    ap.add_argument("--config-warmup", type=int, default=1)
    ap.add_argument("--config-traj", type=int, default=1_000_000, help="trajectories of C1/C2 per GPU and of C5 in total")
    ap.add_argument("--config-traj-mm", type=int, default=10_000_000, help="trajectories of C3 in total")
    ap.add_argument("--cpu-traj-per-core", type=int, default=1024, help="cpu_baseline sample size per host core")
    ap.add_argument("--ref-traj-per-core", type=int, default=512, help="--impl reference: trajectories per core per step")
    args = ap.parse_args()

    if not os.path.exists(os.path.join(ROOT, "rebop_b200", "librebop_b200.so")) and int(os.environ.get("LOCAL_RANK", "0")) == 0:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "rebop_b200", "csrc"), "-j8"], stdout=sys.stderr)
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=sys.stderr)
    models = load_models()  # plain data; the product package (and its CUDA library) is imported by the GPU arm only
    model = models.MODELS[args.model]()
    args.tmax = model["tmax"] if args.tmax is None else args.tmax
    args.nb_steps = model["nb_steps"] if args.nb_steps is None else args.nb_steps
    args.n_species = len(model["species"])

    if args.impl == "reference":
        return run_reference(args, model)
    if args.single_process:
        return run_single_process(args, model)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_gpu(args, model)


if __name__ == "__main__":
    sys.exit(main())
