"""Sharding an ensemble over GPUs, and the ensemble mean/variance reduction.

Trajectories are independent (the reference has no exchange step at all), so an ensemble of N
trajectories shards into contiguous index ranges, one per GPU, with no data-path collective; every
trajectory carries its own seed, so results do not depend on the number of GPUs.  The only
collective is optional: the per-(time, species) sums and sums of squares (exact integers, K4 in
csrc/engine.cu) are all-reduced with NCCL and turned into mean and variance.  Integer sums make the
result independent of the reduction order and of the GPU count.

Two ways to use several GPUs:
  * one process per GPU under torch.distributed (bench.py): `shard_range` + `allreduce_sums`;
  * one process, several devices: `run_sharded` drives one host thread per device (the C ABI calls
    release the GIL), which is what `Gillespie.run(..., devices=[...])` uses.
"""
from __future__ import annotations

import threading

import numpy as np

from rebop_b200 import _ffi

__all__ = ("shard_range", "shard_ranges", "finalize_stats", "allreduce_sums", "run_sharded")


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """(first, count) of the contiguous trajectory range owned by `rank`: [rank*N/W, (rank+1)*N/W)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank must be in [0, world)")
    lo = n_total * rank // world
    hi = n_total * (rank + 1) // world
    return lo, hi - lo


def shard_ranges(n_total: int, world: int) -> list[tuple[int, int]]:
    return [shard_range(n_total, r, world) for r in range(world)]


def finalize_stats(sums, sumsq, n: int):
    """Exact integer sums -> (mean, unbiased variance) as float64 arrays.

    The subtraction n*sum(x^2) - (sum x)^2 is done in unbounded integers, so there is no
    cancellation error however large the ensemble is.
    """
    s = [int(v) for v in np.asarray(sums).ravel()]
    q = [int(v) for v in np.asarray(sumsq).ravel()]
    mean = np.array([v / n for v in s], dtype=np.float64)
    if n > 1:
        var = np.array([(n * b - a * a) / (n * (n - 1)) for a, b in zip(s, q)], dtype=np.float64)
    else:
        var = np.full(len(s), np.nan)
    shape = np.asarray(sums).shape
    return mean.reshape(shape), var.reshape(shape)


def allreduce_sums(sums):
    """Sum a torch int64 tensor of [2][rows] partial sums over the process group (NCCL on GPUs,
    gloo in the CPU tests).  A no-op without an initialised process group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def device_sums_as_tensor(batch: "_ffi.Batch", device: int):
    """Zero-copy torch view of a batch's device-resident K4 sums (int64 [2 * rows])."""
    import torch

    dptr, rows = batch.sample_sums_device()
    batch.synchronize()
    holder = type("RebopSums", (), {"__cuda_array_interface__": {
        "shape": (2 * rows,), "typestr": "<i8", "data": (dptr, False), "version": 3}})()
    return torch.as_tensor(holder, device=torch.device("cuda", device)), rows


def run_sharded(net: "_ffi.Network", n_total: int, x0, seeds, tmax: float, nb_steps: int, save_idx, devices,
                kernel: int = _ffi.KERNEL_AUTO, out: np.ndarray | None = None, want_samples: bool = True):
    """Simulate n_total trajectories split over `devices`, one host thread per device.

    seeds: uint64 [n_total].  Returns (samples int32 [nb_steps+1][n_save][n_total] or None,
    sums int64 [rows], sumsq uint64 [rows], events, kernel_ms_max).
    """
    devices = list(devices)
    ranges = shard_ranges(n_total, len(devices))
    n_save = net.n_species if save_idx is None else len(save_idx)
    rows = (nb_steps + 1) * n_save
    if want_samples and out is None:
        out = np.empty((nb_steps + 1, n_save, n_total), dtype=np.int32)
    results: list = [None] * len(devices)

    def work(i):
        lo, cnt = ranges[i]
        if cnt == 0:
            results[i] = (np.zeros(rows, np.int64), np.zeros(rows, np.uint64), 0, 0.0)
            return
        try:
            b = _ffi.Batch(net, cnt, x0, seeds=seeds[lo:lo + cnt], device=devices[i], kernel=kernel)
            try:
                b.run_grid(tmax, nb_steps, save_idx=save_idx)
                if want_samples and rows:
                    b.samples_into(out, lo)  # straight into this shard's columns of the result
                s1, s2 = b.sample_sums() if rows else (np.zeros(0, np.int64), np.zeros(0, np.uint64))
                results[i] = (s1, s2, b.events()[1], b.last_kernel_ms)
            finally:
                b.close()
        except BaseException as e:  # noqa: BLE001 - re-raised in the caller's thread
            results[i] = e

    if len(devices) == 1:
        work(0)
    else:
        threads = [threading.Thread(target=work, args=(i,)) for i in range(len(devices))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    for r in results:
        if isinstance(r, BaseException):
            raise r
    sums = sum(r[0] for r in results)
    sumsq = sum(r[1] for r in results)
    events = sum(r[2] for r in results)
    return out if want_samples else None, sums, sumsq, events, max(r[3] for r in results)
