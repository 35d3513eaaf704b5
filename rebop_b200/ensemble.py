"""Sharding an ensemble over GPUs, and the ensemble mean/variance reduction.

Trajectories are independent (the reference has no exchange step at all), so an ensemble of N
trajectories shards into contiguous index ranges, one per GPU, with no data-path collective; every
trajectory carries its own seed, so results do not depend on the number of GPUs.  The only
collective is optional: the per-(time, species) sums and sums of squares (exact integers, K4 in
csrc/engine.cu) are all-reduced with NCCL and turned into mean and variance.  Integer sums make the
result independent of the reduction order and of the GPU count.

Two ways to use several GPUs:
  * one process per GPU under torch.distributed (bench.py): `shard_range` + `allreduce_sums`;
  * one process, several devices: `run_sharded` drives the C ABI's ensemble handle (rebop_ensemble_*: one host
    worker thread per device inside the library, NCCL all-reduce of the sums), which is what
    `Gillespie.run(..., devices=[...])` uses.
"""
from __future__ import annotations

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
                kernel: int = _ffi.KERNEL_AUTO, out: np.ndarray | None = None, want_samples: bool = True,
                want_sums: bool = True, dtype=np.int32):
    """Simulate n_total trajectories split over `devices` through the C ABI's ensemble handle
    (rebop_ensemble_*: one batch and one host worker thread per device, contiguous trajectory ranges).

    seeds: uint64 [n_total].  Returns (samples [nb_steps+1][n_save][n_total] of `dtype` or None,
    sums int64 [rows] or None, sumsq uint64 [rows] or None, events, kernel_ms_max); the sums are the exact integer
    row sums, all-reduced over the devices with NCCL inside the library when there is more than one.
    """
    devices = list(devices)
    n_save = net.n_species if save_idx is None else len(save_idx)
    rows = (nb_steps + 1) * n_save
    dtype = np.dtype(dtype)
    if want_samples and out is None:
        out = np.empty((nb_steps + 1, n_save, n_total), dtype=dtype)
    e = _ffi.Ensemble(net, n_total, x0, devices, seeds=seeds, kernel=kernel, dtype=dtype)
    try:
        e.run_grid(tmax, nb_steps, save_idx=save_idx, host_out=out if (want_samples and rows) else None)
        sums = sumsq = None
        if want_sums and rows:
            sums, sumsq = e.sums()
        return out if want_samples else None, sums, sumsq, e.events()[1], e.last_kernel_ms
    finally:
        e.close()
