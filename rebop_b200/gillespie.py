"""Python front end: the reference's `rebop.Gillespie`, extended with a batch of trajectories.

Mirrors python/rebop/gillespie.py:22-160 and the pyo3 class behind it
(src/pyo3_gillespie.rs:65-253): same method names, argument meaning, species ordering,
seed derivation, warnings and exception messages.  The simulation itself runs on the GPU
through the C ABI (rebop_b200/_ffi.py); there is no CPU fallback.

New relative to the reference: ``run(..., n_trajectories=N)`` simulates N independent
trajectories in one launch and returns variables with dims ``("time", "trajectory")``.
Trajectory n uses the seed the reference would have derived for the n-th successive
``run`` call on the same numpy generator (python/rebop/gillespie.py:139-140).
"""
from __future__ import annotations

import warnings
from collections.abc import Mapping, Sequence

import numpy as np

from rebop_b200 import _ffi, ensemble
from rebop_b200.dataset import make_dataset

__all__ = ("Gillespie",)

_KERNELS = {"auto": _ffi.KERNEL_AUTO, "table": _ffi.KERNEL_TABLE, "nvrtc": _ffi.KERNEL_NVRTC, "prebuilt": _ffi.KERNEL_PREBUILT,
            "fast": _ffi.KERNEL_PDM}
_U64_MAX = np.iinfo(np.uint64).max


class _Rate:
    """PRate (src/pyo3_gillespie.rs:23-62): a mass-action constant or a parsed expression."""

    def __init__(self, rate, reactants):
        if isinstance(rate, str):
            try:
                self.expr = _ffi.PExpr(rate)
            except _ffi.RebopError as e:
                if e.status == _ffi.ERR_PARSE:
                    raise ValueError("Rate expression not understood") from None
                raise
            self.k = None
        else:
            self.expr = None
            self.k = float(rate)
        self.reactants = list(reactants)

    def __str__(self):
        if self.expr is not None:
            return str(self.expr)
        return f"LMA({_format_f64(self.k)})"


def _format_f64(v: float) -> str:
    """Rust's `{}` for f64 (shortest round-trip digits, never scientific notation)."""
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    if v == int(v) and abs(v) < 1e16:
        return str(int(v)) if not (v == 0 and np.signbit(v)) else "-0"
    return np.format_float_positional(v, unique=True, trim="-")


class Gillespie:
    """Reaction system composed of species and reactions."""

    def __init__(self) -> None:
        self._species: dict[str, int] = {}  # name -> index, insertion order (src/pyo3_gillespie.rs:76-80)
        self._init: dict[str, int] = {}
        self._reactions: list[tuple[_Rate, list[str], list[str]]] = []

    # -- model building ---------------------------------------------------
    def add_species(self, species: str) -> None:
        """Register a species to the model."""
        if species not in self._species:
            self._species[species] = len(self._species)

    def nb_species(self) -> int:
        """Number of species currently in the system."""
        return len(self._species)

    def nb_reactions(self) -> int:
        """Number of reactions currently in the system."""
        return len(self._reactions)

    def add_reaction(self, rate, reactants: Sequence[str], products: Sequence[str], reverse_rate=None) -> None:
        """Add a reaction to the system.

        A numeric `rate` is the constant of a Law of Mass Action; a string is the expression of
        an arbitrary rate over species and `params`.  `reverse_rate` adds the reverse reaction.
        (src/pyo3_gillespie.rs:86-113: the rate is parsed before any species is registered.)
        """
        reactants = [str(r) for r in reactants]
        products = [str(p) for p in products]
        prate = _Rate(rate, reactants)
        for name in reactants:
            self.add_species(name)
        for name in products:
            self.add_species(name)
        self._reactions.append((prate, reactants, products))
        if reverse_rate is not None:
            self.add_reaction(reverse_rate, products, reactants, None)

    def set_init(self, init: Mapping[str, int]) -> None:
        """Set the initial count of species; unknown names become species and raise UserWarning
        *after* the values are stored (src/pyo3_gillespie.rs:119-134)."""
        checked = {}
        for name, value in init.items():
            v = int(value)
            if v < 0:
                raise OverflowError("can't convert negative int to unsigned")
            checked[str(name)] = v
        warning = False
        for name in checked:
            if name not in self._species:
                warning = True
                self.add_species(name)
        self._init = checked
        if warning:
            raise UserWarning(
                "Some species are not involved in any reactions. You should probably instead use parameters.")

    def __str__(self) -> str:
        s = f"{len(self._species)} species and {len(self._reactions)} reactions\n"
        for rate, reactants, products in self._reactions:
            s += " + ".join(reactants) + " --> " + " + ".join(products) + f" @ {rate}\n"
        return s

    # -- lowering -----------------------------------------------------------
    def _lower(self, params: Mapping[str, float], arith: int) -> _ffi.Network:
        """Names -> indices, reactant multiset -> exponents, jump = -reactants + products
        (src/pyo3_gillespie.rs:46-62,180-196)."""
        names = list(self._species)
        S = len(names)
        net = _ffi.Network(S, arith)
        for rate, reactants, products in self._reactions:
            diff = [0] * S
            for r in reactants:
                diff[self._species[r]] -= 1
            for p in products:
                diff[self._species[p]] += 1
            if rate.expr is None:
                exps = [0] * S
                for r in reactants:
                    exps[self._species[r]] += 1
                net.add_reaction_lma(rate.k, exps, diff)
            else:
                try:
                    program = rate.expr.lower(names, dict(params))
                except _ffi.RebopError as e:
                    if e.status == _ffi.ERR_MISSING_PARAM:
                        raise ValueError(str(e)) from None
                    raise
                net.add_reaction_expr(program, diff)
        return net

    def _run_events(self, net, n, batched, x0, seeds, tmax, names, save_names, device, kernel, dtype, reduce):
        """nb_steps = 0 (src/pyo3_gillespie.rs:209-223): every reaction is returned, ending with the first that
        happens at or after tmax.  One Dataset per trajectory: their lengths differ."""
        if reduce:
            raise ValueError("reduce=True needs a time grid (nb_steps > 0): event times differ between trajectories")
        if not names:
            empty = make_dataset({}, np.array([0.0, float("inf")]), batched=False)
            return [empty] * n if batched else empty
        uniq = sorted({self._species[v] for v in save_names})
        row = {idx: j for j, idx in enumerate(uniq)}
        b = _ffi.Batch(net, n, x0, seeds=seeds, device=device, kernel=kernel)
        try:
            offsets, times, samples = b.run_events(tmax, uniq)
            self.last_events, self.last_kernel_ms = b.events()[1], b.last_kernel_ms
        finally:
            b.close()
        samples = samples.astype(np.dtype(dtype), copy=False)
        out = []
        for i in range(n):
            lo, hi = int(offsets[i]), int(offsets[i + 1])
            out.append(make_dataset({v: samples[row[self._species[v]], lo:hi] for v in save_names}, times[lo:hi],
                                    batched=False))
        return out if batched else out[0]

    # -- simulation -----------------------------------------------------------
    def run(  # noqa: PLR0913
        self,
        init: Mapping[str, int],
        tmax: float,
        nb_steps: int,
        *,
        params: Mapping[str, float] | None = None,
        rng=None,
        sparse: bool | None = None,
        var_names: Sequence[str] | None = None,
        n_trajectories: int | None = None,
        device: int = 0,
        devices: Sequence[int] | None = None,
        kernel: str = "auto",
        dtype=np.int64,
        reduce: bool = False,
    ):
        """Run the system until `tmax` with `nb_steps` steps.

        Parameters are those of the reference (python/rebop/gillespie.py:96-137); `sparse` is
        accepted for compatibility and has no effect on results (the reference guarantees dense
        and sparse runs are identical, tests/test_rebop.py:55-65).

        n_trajectories : int | None
            Number of independent trajectories simulated in one GPU launch.  None reproduces the
            reference call: one trajectory, variables with the single dim ``time``.
        device, devices, kernel, dtype
            CUDA device index, or a list of devices the trajectories are sharded over (contiguous
            ranges, one host thread per device; results do not depend on the sharding);
            "auto" | "nvrtc" | "table"; dtype of the returned counts: np.int64 like the reference, np.int32, or
            np.int16 (the samples are produced in that type on the device, so a narrower type shrinks the
            device-to-host copy; a count that does not fit int16 raises instead of wrapping).
        reduce : bool
            Return the ensemble mean and variance of every saved species at every sample time
            (variables ``<name>_mean`` and ``<name>_var``, dim ``time``) instead of the
            trajectories; the samples never leave the GPUs.

        nb_steps = 0 returns every reaction (one row per event, the last at or after `tmax`); with
        `n_trajectories` the result is then a list of Datasets, one per trajectory (their lengths differ).

        Returns an `xarray.Dataset` when xarray is importable, otherwise a small stand-in with
        the same access patterns (`ds.S`, `ds["S"]`, `ds.time`).
        """
        params = {} if params is None else params
        rng_ = np.random.default_rng(rng)
        n = 1 if n_trajectories is None else int(n_trajectories)
        if n < 1:
            raise ValueError("n_trajectories must be at least 1")
        # one u64 per trajectory, exactly the draws N successive reference runs would make
        seeds = rng_.integers(_U64_MAX, size=n, dtype=np.uint64)
        try:
            self.set_init(init)
        except UserWarning as e:
            warnings.warn(e, stacklevel=2)
        for p in params:
            if p in self._species:
                raise ValueError(f"Species {p} cannot also be a parameter.")
        names = list(self._species)
        x0 = [0] * len(names)
        for name, value in self._init.items():
            if name in self._species:
                x0[self._species[name]] = value
        if var_names is None:
            save_names = names
        else:
            save_names = [str(v) for v in var_names]
            for v in save_names:
                if v not in self._species:
                    raise KeyError(v)  # the reference panics on an unknown name (src/pyo3_gillespie.rs:175)
        nb_steps = int(nb_steps)
        if nb_steps < 0:
            raise OverflowError("can't convert negative int to unsigned")
        net = self._lower(params, _ffi.ARITH_API)
        tmax = float(tmax)
        if nb_steps == 0:
            return self._run_events(net, n, n_trajectories is not None, x0, seeds, tmax, names, save_names, device,
                                    _KERNELS[kernel], dtype, reduce)
        times = np.array([tmax * float(i) / float(nb_steps) for i in range(nb_steps + 1)], dtype=np.float64)

        values = {}
        if names:
            devs = [device] if devices is None else list(devices)
            # rows come back in the order requested; duplicates are allowed like in the reference
            uniq = sorted({self._species[v] for v in save_names})
            sample_dtype = np.dtype(dtype) if np.dtype(dtype) in (np.dtype(np.int16), np.dtype(np.int32), np.dtype(np.int64)) \
                else np.dtype(np.int32)
            samples, sums, sumsq, self.last_events, self.last_kernel_ms = ensemble.run_sharded(
                net, n, x0, seeds, tmax, nb_steps, uniq, devs, kernel=_KERNELS[kernel], want_samples=not reduce,
                want_sums=reduce, dtype=sample_dtype)
            row = {idx: j for j, idx in enumerate(uniq)}
            if reduce:
                mean, var = ensemble.finalize_stats(sums, sumsq, n)
                mean = mean.reshape(nb_steps + 1, len(uniq))
                var = var.reshape(nb_steps + 1, len(uniq))
                for v in save_names:
                    values[v + "_mean"] = mean[:, row[self._species[v]]]
                    values[v + "_var"] = var[:, row[self._species[v]]]
                return make_dataset(values, times, batched=False)
            if uniq:
                samples = samples.astype(np.dtype(dtype), copy=False)
            for v in save_names:
                col = samples[:, row[self._species[v]], :]
                values[v] = col[:, 0] if n_trajectories is None else col
        return make_dataset(values, times, batched=n_trajectories is not None)
