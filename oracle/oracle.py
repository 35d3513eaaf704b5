"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package
``rebop_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libssa_oracle.so")

ARITH_API, ARITH_MACRO = 0, 1
OP = dict(const=0, species=1, neg=2, add=3, sub=4, mul=5, div=6, pow=7, max=8, min=9, exp=10)


class ExprOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("index", C.c_int32), ("value", C.c_double)]


class Rng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4)]


class State(C.Structure):
    _fields_ = [("x", C.POINTER(C.c_int64)), ("t", C.c_double), ("rng", Rng), ("events", C.c_uint64)]


def build(force: bool = False) -> str:
    """Compile libssa_oracle.so with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32p, i64p, u64p, f64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.POINTER(C.c_double)
        L.ora_rng_seed.argtypes = [C.POINTER(Rng), C.c_uint64]
        L.ora_rng_next_u64.argtypes = [C.POINTER(Rng)]
        L.ora_rng_next_u64.restype = C.c_uint64
        L.ora_rng_uniform.argtypes = [C.POINTER(Rng)]
        L.ora_rng_uniform.restype = C.c_double
        L.ora_rng_exp1.argtypes = [C.POINTER(Rng)]
        L.ora_rng_exp1.restype = C.c_double
        L.ora_network_new.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ora_network_new.restype = vp
        L.ora_network_free.argtypes = [vp]
        L.ora_network_add_lma.argtypes = [vp, C.c_double, i32p, i32p, C.c_int, i64p]
        L.ora_network_add_expr.argtypes = [vp, C.POINTER(ExprOp), C.c_int, i64p]
        L.ora_rate.argtypes = [vp, C.c_int, i64p]
        L.ora_rate.restype = C.c_double
        L.ora_expr_eval.argtypes = [C.POINTER(ExprOp), C.c_int, i64p]
        L.ora_expr_eval.restype = C.c_double
        L.ora_advance_until.argtypes = [vp, C.POINTER(State), C.c_double]
        L.ora_advance_one_reaction.argtypes = [vp, C.POINTER(State)]
        L.ora_run_grid.argtypes = [vp, i64p, C.c_uint64, C.c_double, C.c_int, i32p, C.c_int, i64p, f64p]
        L.ora_run_grid.restype = C.c_uint64
        L.ora_run_events.argtypes = [vp, i64p, C.c_uint64, C.c_double, i32p, C.c_int, i64p, f64p, C.c_size_t]
        L.ora_run_events.restype = C.c_size_t
        L.ora_run_batch.argtypes = [vp, i64p, C.c_size_t, u64p, C.c_size_t, C.c_double, C.c_int, i32p, C.c_int, i32p, u64p, C.c_int]
        L.ora_run_batch.restype = C.c_uint64
        L.ora_run_batch_macro.argtypes = [C.c_char_p, f64p, i64p, u64p, C.c_size_t, C.c_double, C.c_int, i32p, u64p, C.c_int]
        L.ora_run_batch_macro.restype = C.c_uint64
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def make_prog(ops: Sequence[tuple]) -> "C.Array[ExprOp]":
    """ops: sequence of (opname_or_code, index, value) in post-order."""
    arr = (ExprOp * max(1, len(ops)))()
    for i, (op, idx, val) in enumerate(ops):
        arr[i].op = OP[op] if isinstance(op, str) else int(op)
        arr[i].index = int(idx)
        arr[i].value = float(val)
    return arr


class RngStream:
    """SmallRng::seed_from_u64 + draws, for the known-answer tests."""

    def __init__(self, seed: int):
        self._r = Rng()
        lib().ora_rng_seed(C.byref(self._r), C.c_uint64(seed))

    @property
    def state(self):
        return [int(v) for v in self._r.s]

    def next_u64(self) -> int:
        return int(lib().ora_rng_next_u64(C.byref(self._r)))

    def uniform(self) -> float:
        return float(lib().ora_rng_uniform(C.byref(self._r)))

    def exp1(self) -> float:
        return float(lib().ora_rng_exp1(C.byref(self._r)))


class Network:
    """A reaction network in the oracle.

    reactions: list of ("lma", k, [(species_index, exponent), ...], diff[S])
               or      ("expr", [(op, index, value), ...], diff[S]).
    """

    def __init__(self, n_species: int, reactions=(), arith: int = ARITH_API, dense: bool = False):
        self.n_species = int(n_species)
        self.arith = arith
        self._h = lib().ora_network_new(self.n_species, arith, 1 if dense else 0)
        self.n_reactions = 0
        for r in reactions:
            if r[0] == "lma":
                self.add_lma(r[1], r[2], r[3])
            else:
                self.add_expr(r[1], r[2])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_network_free(self._h)
            self._h = None

    def _diff(self, diff):
        d = np.ascontiguousarray(diff, dtype=np.int64)
        assert d.shape == (self.n_species,)
        return d

    def add_lma(self, k, terms, diff):
        idx = np.ascontiguousarray([t[0] for t in terms], dtype=np.int32)
        ex = np.ascontiguousarray([t[1] for t in terms], dtype=np.int32)
        d = self._diff(diff)
        rc = lib().ora_network_add_lma(self._h, float(k), _p(idx, C.c_int32), _p(ex, C.c_int32), len(terms), _p(d, C.c_int64))
        if rc != 0:
            raise IndexError("reactant index out of range")
        self.n_reactions += 1

    def add_expr(self, ops, diff):
        d = self._diff(diff)
        prog = make_prog(ops)
        lib().ora_network_add_expr(self._h, prog, len(ops), _p(d, C.c_int64))
        self.n_reactions += 1

    def rate(self, r: int, x) -> float:
        xa = np.ascontiguousarray(x, dtype=np.int64)
        return float(lib().ora_rate(self._h, r, _p(xa, C.c_int64)))

    def step_one(self, x0, seed: int, n_steps: int, t0: float = 0.0):
        """n_steps calls of Gillespie::advance_one_reaction (src/gillespie.rs:270-297) -> (t, x[S], events)."""
        x = np.ascontiguousarray(x0, dtype=np.int64).copy()
        st = State()
        st.x = _p(x, C.c_int64)
        st.t = float(t0)
        st.events = 0
        lib().ora_rng_seed(C.byref(st.rng), C.c_uint64(int(seed)))
        for _ in range(n_steps):
            lib().ora_advance_one_reaction(self._h, C.byref(st))
        return float(st.t), x, int(st.events)

    def run_grid(self, x0, seed: int, tmax: float, nb_steps: int, save_idx=None):
        """pyo3 grid loop for one trajectory -> (times[nb+1], out[nb+1][n_save], events)."""
        x0a = np.ascontiguousarray(x0, dtype=np.int64)
        save = np.arange(self.n_species, dtype=np.int32) if save_idx is None else np.ascontiguousarray(save_idx, dtype=np.int32)
        out = np.zeros((nb_steps + 1, len(save)), dtype=np.int64)
        times = np.zeros(nb_steps + 1, dtype=np.float64)
        ev = lib().ora_run_grid(self._h, _p(x0a, C.c_int64), C.c_uint64(int(seed)), float(tmax), int(nb_steps),
                                _p(save, C.c_int32), len(save), _p(out, C.c_int64), _p(times, C.c_double))
        return times, out, int(ev)

    def run_events(self, x0, seed: int, tmax: float, save_idx=None, cap: int = 1 << 20):
        """nb_steps == 0 path -> (times[n], out[n][n_save])."""
        x0a = np.ascontiguousarray(x0, dtype=np.int64)
        save = np.arange(self.n_species, dtype=np.int32) if save_idx is None else np.ascontiguousarray(save_idx, dtype=np.int32)
        out = np.zeros((cap, len(save)), dtype=np.int64)
        times = np.zeros(cap, dtype=np.float64)
        n = lib().ora_run_events(self._h, _p(x0a, C.c_int64), C.c_uint64(int(seed)), float(tmax),
                                 _p(save, C.c_int32), len(save), _p(out, C.c_int64), _p(times, C.c_double), cap)
        if n > cap:
            raise OverflowError("event log larger than cap")
        return times[:n].copy(), out[:n].copy()

    def run_batch(self, x0, seeds, tmax: float, nb_steps: int, save_idx=None, threads: int = 1,
                  want_out: bool = True, want_events: bool = True):
        """Ensemble -> (out[nb+1][n_save][N] int32 or None, events_per_traj or None, total_events)."""
        x0a = np.ascontiguousarray(x0, dtype=np.int64)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        n = len(seeds)
        stride = 0 if x0a.ndim == 1 else self.n_species
        save = np.arange(self.n_species, dtype=np.int32) if save_idx is None else np.ascontiguousarray(save_idx, dtype=np.int32)
        rows = (nb_steps + 1) if nb_steps > 0 else 1
        out = np.zeros((rows, len(save), n), dtype=np.int32) if want_out else None
        evs = np.zeros(n, dtype=np.uint64) if want_events else None
        tot = lib().ora_run_batch(self._h, _p(x0a, C.c_int64), stride, _p(seeds, C.c_uint64), n, float(tmax), int(nb_steps),
                                  _p(save, C.c_int32), len(save),
                                  _p(out, C.c_int32) if want_out else None,
                                  _p(evs, C.c_uint64) if want_events else None, int(threads))
        return out, evs, int(tot)


def run_batch_macro(name: str, params, x0, seeds, tmax: float, nb_steps: int, threads: int = 1,
                    want_out: bool = True, want_events: bool = True):
    """Hand-expanded define_system! form of a benchmark system ("vilar", "dimers", "sir")."""
    p = np.ascontiguousarray(params, dtype=np.float64)
    x0a = np.ascontiguousarray(x0, dtype=np.int64)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
    n = len(seeds)
    rows = (nb_steps + 1) if nb_steps > 0 else 1
    out = np.zeros((rows, len(x0a), n), dtype=np.int32) if want_out else None
    evs = np.zeros(n, dtype=np.uint64) if want_events else None
    tot = lib().ora_run_batch_macro(name.encode(), _p(p, C.c_double), _p(x0a, C.c_int64), _p(seeds, C.c_uint64), n,
                                    float(tmax), int(nb_steps),
                                    _p(out, C.c_int32) if want_out else None,
                                    _p(evs, C.c_uint64) if want_events else None, int(threads))
    if tot == 2**64 - 1:
        raise KeyError(name)
    return out, evs, int(tot)
