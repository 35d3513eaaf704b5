"""`define_system!` for the GPU engine (src/gillespie_macro.rs:49-129 of the reference).

    Dimers = rebop_b200.define_system('''
        rtx rtl rdi rdm rdp;
        Dimers { gene, mRNA, protein, dimer }
        transcription   : gene      => gene + mRNA      @ rtx
        translation     : mRNA      => mRNA + protein   @ rtl
        dimerization    : 2 protein => dimer            @ rdi
        decay_mRNA      : mRNA      =>                  @ rdm
        decay_prot      : protein   =>                  @ rdp
    ''')
    dimers = Dimers.with_parameters(25., 1000., 0.001, 0.1, 1.)
    dimers.gene = 1
    dimers.seed(0)
    dimers.advance_until(1.)
    print(dimers.t, dimers.dimer)

The object mirrors the struct the macro generates: one attribute per species (0 after `new`),
one per parameter (NaN after `new`: "If a NAN remains at the time of the simulation, no reaction
will happen"), `t`, `seed(u64)` and `advance_until(tmax)`.  The extension is the ensemble:
`new(n_trajectories=N)` makes every species attribute and `t` an array of N values, trajectory n
seeded with `seed + n` (or with `seeds[n]`).  Systems listed under rebop_b200/systems/ were turned
into CUDA at build time (rebop_sysgen + nvcc, the `build.rs` analogue) and run without NVRTC; any
other text is specialised at run time.
"""
from __future__ import annotations

import os

import numpy as np

from rebop_b200 import _ffi

__all__ = ("define_system", "DefinedSystem", "SystemState")


class DefinedSystem:
    """What `define_system!` defines: a type with `new` and `with_parameters` constructors."""

    def __init__(self, text: str):
        self._sys = _ffi.System(text)
        self.name = self._sys.name
        self.params = list(self._sys.params)
        self.species = list(self._sys.species)
        self.reactions = list(self._sys.reactions)

    def new(self, n_trajectories: int | None = None, device: int = 0, kernel: int = _ffi.KERNEL_AUTO) -> "SystemState":
        """`Name::new()`: species 0, parameters NaN, t = 0, RNG seeded from OS entropy."""
        return SystemState(self, [float("nan")] * len(self.params), n_trajectories, device, kernel)

    def with_parameters(self, *params, n_trajectories: int | None = None, device: int = 0,
                        kernel: int = _ffi.KERNEL_AUTO) -> "SystemState":
        """`Name::with_parameters(p...)`."""
        if len(params) != len(self.params):
            raise TypeError(f"{self.name}.with_parameters() takes {len(self.params)} parameters ({len(params)} given)")
        return SystemState(self, [float(p) for p in params], n_trajectories, device, kernel)

    def __repr__(self) -> str:
        return f"<define_system {self.name}: {len(self.species)} species, {len(self.reactions)} reactions>"


class SystemState:
    """An instance of the generated struct, resident on one GPU once it has been advanced."""

    def __init__(self, system: DefinedSystem, params, n_trajectories, device, kernel):
        d = object.__setattr__
        d(self, "_system", system)
        d(self, "_scalar", n_trajectories is None)
        d(self, "_n", 1 if n_trajectories is None else int(n_trajectories))
        d(self, "_params", dict(zip(system.params, params)))
        d(self, "_x", np.zeros((self._n, len(system.species)), dtype=np.int64))
        d(self, "_t", np.zeros(self._n, dtype=np.float64))
        d(self, "_seeds", None)
        d(self, "_seed_base", int.from_bytes(os.urandom(8), "little"))  # rand::make_rng()
        d(self, "_batch", None)
        d(self, "_device", device)
        d(self, "_kernel", kernel)
        d(self, "_dirty_x", True)
        d(self, "_dirty_t", False)
        d(self, "_dirty_k", False)
        d(self, "_dirty_seed", False)

    # ---- struct fields ----
    def __getattr__(self, name):
        sysm = object.__getattribute__(self, "_system")
        if name in sysm.species:
            col = self._x[:, sysm.species.index(name)]
            return int(col[0]) if self._scalar else col.copy()
        if name in sysm.params:
            return self._params[name]
        if name == "t":
            return float(self._t[0]) if self._scalar else self._t.copy()
        raise AttributeError(f"no field `{name}` on type `{sysm.name}`")

    def __setattr__(self, name, value):
        sysm = self._system
        if name in sysm.species:
            self._x[:, sysm.species.index(name)] = value
            object.__setattr__(self, "_dirty_x", True)
        elif name in sysm.params:
            self._params[name] = float(value)
            object.__setattr__(self, "_dirty_k", True)
        elif name == "t":
            self._t[:] = value
            object.__setattr__(self, "_dirty_t", True)
        else:
            raise AttributeError(f"no field `{name}` on type `{sysm.name}`")

    def seed(self, seed) -> None:
        """`seed(u64)`; an array gives every trajectory its own seed, a scalar seeds trajectory n with seed + n."""
        if np.ndim(seed) == 0:
            object.__setattr__(self, "_seeds", None)
            object.__setattr__(self, "_seed_base", int(seed))
        else:
            seeds = np.ascontiguousarray(seed, dtype=np.uint64)
            if seeds.shape != (self._n,):
                raise ValueError("one seed per trajectory expected")
            object.__setattr__(self, "_seeds", seeds)
        object.__setattr__(self, "_dirty_seed", True)

    def advance_until(self, tmax: float) -> None:
        """Simulates every trajectory until t = tmax (src/gillespie_macro.rs:98-126)."""
        sysm = self._system
        pv = [self._params[p] for p in sysm.params]
        b = self._batch
        if b is None:
            net = sysm._sys.network(pv)
            b = _ffi.Batch(net, self._n, self._x, seeds=self._seeds, seed_base=self._seed_base, device=self._device,
                           kernel=self._kernel)
            object.__setattr__(self, "_batch", b)
            object.__setattr__(self, "_dirty_x", False)
            object.__setattr__(self, "_dirty_k", False)
            object.__setattr__(self, "_dirty_seed", False)
        if self._dirty_x:
            b.set_species(self._x)
        if self._dirty_t:
            if not np.all(self._t == self._t[0]):
                raise ValueError("t can only be set to one value for the whole ensemble")
            b.set_time(float(self._t[0]))
        if self._dirty_k:
            b.set_rates(sysm._sys.rates(pv))
        if self._dirty_seed:
            b.seed(self._seeds, self._seed_base)
        for flag in ("_dirty_x", "_dirty_t", "_dirty_k", "_dirty_seed"):
            object.__setattr__(self, flag, False)
        b.advance_until(float(tmax))
        self._x[:] = b.species()
        self._t[:] = b.times()

    @property
    def events(self) -> int:
        """Reactions applied so far, over all trajectories."""
        return 0 if self._batch is None else self._batch.events()[0]

    @property
    def kernel_used(self) -> int:
        return _ffi.KERNEL_AUTO if self._batch is None else self._batch.kernel_used


def define_system(text: str) -> DefinedSystem:
    return DefinedSystem(text)
