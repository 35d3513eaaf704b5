"""ctypes binding of the C ABI declared in include/rebop_b200.h.

This is the Python-side equivalent of the Rust `-sys` crate: thin, one function per
exported symbol, no logic.  The library is built in-tree (`rebop_b200/librebop_b200.so`);
if it is missing the import fails loudly -- there is no CPU or pure-Python fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np  # noqa: E402  (used by the constants below)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librebop_b200.so")

(OK, ERR_INVALID, ERR_OUT_OF_RANGE, ERR_PARSE, ERR_MISSING_PARAM, ERR_CUDA, ERR_NVRTC, ERR_LIMIT, ERR_ITER_CAP,
 ERR_NCCL) = range(10)
SAMPLES_I16, SAMPLES_I32, SAMPLES_I64 = 2, 4, 8
SAMPLE_DTYPES = {2: np.int16, 4: np.int32, 8: np.int64}
SCHEDULE_AUTO, SCHEDULE_STATIC, SCHEDULE_SPARSE, SCHEDULE_DENSE = 0, 1, 2, 3
ARITH_API, ARITH_MACRO = 0, 1
KERNEL_AUTO, KERNEL_TABLE, KERNEL_NVRTC, KERNEL_PREBUILT, KERNEL_PDM = 0, 1, 2, 3, 4
OPCODES = dict(const=0, species=1, neg=2, add=3, sub=4, mul=5, div=6, pow=7, max=8, min=9, exp=10)


class ExprOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("index", C.c_int32), ("value", C.c_double)]


class RebopError(RuntimeError):
    """A non-zero rebop_status; `.status` holds the code."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C rebop_b200/csrc` (rebop_b200 has no CPU fallback)")
    return C.CDLL(LIB_PATH)


lib = _load()

_vp = C.c_void_p
_pp = C.POINTER(C.c_void_p)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)
_szp = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); every symbol of include/rebop_b200.h
SIGNATURES = {
    "rebop_b200_last_error": (C.c_char_p, []),
    "rebop_b200_version": (C.c_char_p, []),
    "rebop_b200_device_count": (C.c_int, []),
    "rebop_network_create": (C.c_int, [C.c_uint32, C.c_int, _pp]),
    "rebop_network_destroy": (None, [_vp]),
    "rebop_network_add_reaction_lma": (C.c_int, [_vp, C.c_double, _u32p, _i64p]),
    "rebop_network_add_reaction_lma_sparse": (C.c_int, [_vp, C.c_double, _u32p, _u32p, C.c_size_t, _i64p]),
    "rebop_network_add_reaction_expr": (C.c_int, [_vp, C.POINTER(ExprOp), C.c_size_t, _i64p]),
    "rebop_network_nb_species": (C.c_int, [_vp, _u32p]),
    "rebop_network_nb_reactions": (C.c_int, [_vp, _u32p]),
    "rebop_network_codegen": (C.c_int, [_vp, C.c_char_p, C.c_size_t, _szp]),
    "rebop_network_jit_cubin": (C.c_int, [_vp, C.c_char_p, C.c_size_t, _szp]),
    "rebop_network_codegen_pdm": (C.c_int, [_vp, C.c_char_p, C.c_size_t, _szp]),
    "rebop_network_jit_cubin_pdm": (C.c_int, [_vp, C.c_char_p, C.c_size_t, _szp]),
    "rebop_system_parse": (C.c_int, [C.c_char_p, _pp]),
    "rebop_system_destroy": (None, [_vp]),
    "rebop_system_name": (C.c_int, [_vp, C.c_char_p, C.c_size_t, _szp]),
    "rebop_system_counts": (C.c_int, [_vp, _u32p, _u32p, _u32p]),
    "rebop_system_param_name": (C.c_int, [_vp, C.c_uint32, C.c_char_p, C.c_size_t, _szp]),
    "rebop_system_species_name": (C.c_int, [_vp, C.c_uint32, C.c_char_p, C.c_size_t, _szp]),
    "rebop_system_reaction_name": (C.c_int, [_vp, C.c_uint32, C.c_char_p, C.c_size_t, _szp]),
    "rebop_system_network": (C.c_int, [_vp, _f64p, C.c_size_t, _pp]),
    "rebop_system_rates": (C.c_int, [_vp, _f64p, C.c_size_t, _f64p]),
    "rebop_b200_prebuilt_count": (C.c_int, []),
    "rebop_b200_prebuilt_name": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t, _szp]),
    "rebop_network_has_prebuilt": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "rebop_pexpr_parse": (C.c_int, [C.c_char_p, _pp]),
    "rebop_pexpr_destroy": (None, [_vp]),
    "rebop_pexpr_format": (C.c_int, [_vp, C.c_char_p, C.c_size_t, _szp]),
    "rebop_pexpr_lower": (C.c_int, [_vp, C.POINTER(C.c_char_p), C.c_size_t, C.POINTER(C.c_char_p), _f64p, C.c_size_t,
                                    C.POINTER(ExprOp), C.c_size_t, _szp]),
    "rebop_batch_create": (C.c_int, [_vp, C.c_int, C.c_size_t, _i64p, C.c_int, _u64p, C.c_uint64, _pp]),
    "rebop_batch_destroy": (None, [_vp]),
    "rebop_batch_set_kernel": (C.c_int, [_vp, C.c_int]),
    "rebop_batch_get_kernel": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "rebop_batch_set_max_iters": (C.c_int, [_vp, C.c_uint32]),
    "rebop_batch_set_schedule": (C.c_int, [_vp, C.c_int]),
    "rebop_batch_get_schedule": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "rebop_batch_set_rates": (C.c_int, [_vp, _f64p, C.c_size_t]),
    "rebop_batch_seed": (C.c_int, [_vp, _u64p, C.c_uint64]),
    "rebop_batch_get_time": (C.c_int, [_vp, _f64p]),
    "rebop_batch_set_time": (C.c_int, [_vp, C.c_double]),
    "rebop_batch_get_species": (C.c_int, [_vp, _i64p]),
    "rebop_batch_set_species": (C.c_int, [_vp, _i64p, C.c_int]),
    "rebop_batch_advance_until": (C.c_int, [_vp, C.c_double]),
    "rebop_batch_advance_one_reaction": (C.c_int, [_vp]),
    "rebop_batch_set_sample_dtype": (C.c_int, [_vp, C.c_int]),
    "rebop_batch_get_sample_dtype": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "rebop_batch_run_grid_typed": (C.c_int, [_vp, C.c_double, C.c_uint32, _u32p, C.c_uint32, _vp]),
    "rebop_batch_run_grid_strided": (C.c_int, [_vp, C.c_double, C.c_uint32, _u32p, C.c_uint32, _vp, C.c_size_t]),
    "rebop_batch_samples_host": (C.c_int, [_vp, _vp]),
    "rebop_batch_samples_host_strided": (C.c_int, [_vp, _vp, C.c_size_t]),
    "rebop_batch_last_finish_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "rebop_batch_run_grid": (C.c_int, [_vp, C.c_double, C.c_uint32, _u32p, C.c_uint32, _i32p]),
    "rebop_batch_run_events": (C.c_int, [_vp, C.c_double, _u32p, C.c_uint32]),
    "rebop_batch_events_log_size": (C.c_int, [_vp, _u64p, _u32p]),
    "rebop_batch_events_log_host": (C.c_int, [_vp, _u64p, _f64p, _i32p]),
    "rebop_batch_samples_device": (C.c_int, [_vp, C.POINTER(C.c_void_p), _szp, _u32p]),
    "rebop_batch_samples_host_i32": (C.c_int, [_vp, _i32p]),
    "rebop_batch_samples_host_i64": (C.c_int, [_vp, _i64p]),
    "rebop_batch_samples_host_i32_strided": (C.c_int, [_vp, _i32p, C.c_size_t]),
    "rebop_batch_sample_sums": (C.c_int, [_vp, _i64p, _u64p]),
    "rebop_batch_sample_sums_device": (C.c_int, [_vp, C.POINTER(C.c_void_p), _u32p]),
    "rebop_batch_events": (C.c_int, [_vp, _u64p, _u64p]),
    "rebop_batch_lane_slots": (C.c_int, [_vp, _u64p]),
    "rebop_batch_last_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "rebop_batch_size": (C.c_int, [_vp, _szp]),
    "rebop_batch_synchronize": (C.c_int, [_vp]),
    "rebop_batch_get_stream": (C.c_int, [_vp, C.POINTER(C.c_void_p)]),
    "rebop_batch_set_stream": (C.c_int, [_vp, _vp]),
    "rebop_ensemble_create": (C.c_int, [_vp, C.POINTER(C.c_int), C.c_int, C.c_size_t, _i64p, C.c_int, _u64p, C.c_uint64, _pp]),
    "rebop_ensemble_destroy": (None, [_vp]),
    "rebop_ensemble_size": (C.c_int, [_vp, _szp]),
    "rebop_ensemble_shards": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "rebop_ensemble_shard": (C.c_int, [_vp, C.c_int, _pp, C.POINTER(C.c_int), _szp, _szp]),
    "rebop_ensemble_set_kernel": (C.c_int, [_vp, C.c_int]),
    "rebop_ensemble_set_schedule": (C.c_int, [_vp, C.c_int]),
    "rebop_ensemble_set_sample_dtype": (C.c_int, [_vp, C.c_int]),
    "rebop_ensemble_set_max_iters": (C.c_int, [_vp, C.c_uint32]),
    "rebop_ensemble_set_rates": (C.c_int, [_vp, _f64p, C.c_size_t]),
    "rebop_ensemble_set_time": (C.c_int, [_vp, C.c_double]),
    "rebop_ensemble_set_species": (C.c_int, [_vp, _i64p, C.c_int]),
    "rebop_ensemble_seed": (C.c_int, [_vp, _u64p, C.c_uint64]),
    "rebop_ensemble_advance_until": (C.c_int, [_vp, C.c_double]),
    "rebop_ensemble_advance_one_reaction": (C.c_int, [_vp]),
    "rebop_ensemble_run_grid": (C.c_int, [_vp, C.c_double, C.c_uint32, _u32p, C.c_uint32, _vp]),
    "rebop_ensemble_samples_host": (C.c_int, [_vp, _vp]),
    "rebop_ensemble_stats": (C.c_int, [_vp, _f64p, _f64p]),
    "rebop_ensemble_sums": (C.c_int, [_vp, _i64p, _u64p]),
    "rebop_ensemble_events": (C.c_int, [_vp, _u64p, _u64p]),
    "rebop_ensemble_last_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "rebop_b200_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "rebop_b200_host_free": (C.c_int, [_vp]),
    "rebop_b200_kernel_launches": (C.c_uint64, []),
    "rebop_b200_measure_fp64_rate": (C.c_int, [C.c_int, _f64p, _f64p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.rebop_b200_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    if status != OK:
        raise RebopError(status, last_error())


def ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def device_count() -> int:
    return int(lib.rebop_b200_device_count())


def make_program(ops) -> "C.Array[ExprOp]":
    """ops: iterable of (opcode or name, index, value) in post-order."""
    ops = list(ops)
    arr = (ExprOp * max(1, len(ops)))()
    for i, (op, idx, val) in enumerate(ops):
        arr[i].op = OPCODES[op] if isinstance(op, str) else int(op)
        arr[i].index = int(idx)
        arr[i].value = float(val)
    return arr


class PExpr:
    """A parsed rate expression (the reference's PExpr)."""

    def __init__(self, text: str):
        h = C.c_void_p()
        check(lib.rebop_pexpr_parse(text.encode("utf-8"), C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib.rebop_pexpr_destroy(self._h)
            self._h = None

    def __str__(self) -> str:
        need = C.c_size_t()
        check(lib.rebop_pexpr_format(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib.rebop_pexpr_format(self._h, buf, need.value, None))
        return buf.value.decode("utf-8")

    def lower(self, species_names, params) -> list:
        """-> [(op, index, value), ...] post-order program (PExpr::to_expr)."""
        sn = (C.c_char_p * max(1, len(species_names)))(*[s.encode("utf-8") for s in species_names])
        pn_list = list(params.keys())
        pn = (C.c_char_p * max(1, len(pn_list)))(*[s.encode("utf-8") for s in pn_list])
        pv = np.ascontiguousarray([float(params[k]) for k in pn_list] or [0.0], dtype=np.float64)
        n = C.c_size_t()
        check(lib.rebop_pexpr_lower(self._h, sn, len(species_names), pn, ptr(pv, C.c_double), len(pn_list), None, 0, C.byref(n)))
        prog = (ExprOp * max(1, n.value))()
        check(lib.rebop_pexpr_lower(self._h, sn, len(species_names), pn, ptr(pv, C.c_double), len(pn_list), prog, n.value, C.byref(n)))
        return [(prog[i].op, prog[i].index, prog[i].value) for i in range(n.value)]


class Network:
    """Owning wrapper of a rebop_network handle."""

    def __init__(self, n_species: int, arith: int = ARITH_API, _handle=None):
        if _handle is None:
            h = C.c_void_p()
            check(lib.rebop_network_create(int(n_species), int(arith), C.byref(h)))
        else:
            h = _handle
        self._h = h
        self.n_species = int(n_species)
        self.arith = int(arith)

    @property
    def has_prebuilt(self) -> bool:
        """True when a kernel compiled at build time (rebop_sysgen + nvcc) matches this network."""
        yes = C.c_int()
        check(lib.rebop_network_has_prebuilt(self._h, C.byref(yes)))
        return bool(yes.value)

    def __del__(self):
        if getattr(self, "_h", None):
            lib.rebop_network_destroy(self._h)
            self._h = None

    def _diff(self, differences) -> np.ndarray:
        d = np.ascontiguousarray(differences, dtype=np.int64)
        if d.shape != (self.n_species,):
            raise RebopError(ERR_INVALID, "assertion failed: differences.len() == nb_species")
        return d

    def add_reaction_lma(self, k: float, exponents, differences) -> None:
        e = np.ascontiguousarray(exponents, dtype=np.uint32)
        if e.shape != (self.n_species,):
            raise RebopError(ERR_INVALID, "assertion failed: stoechiometries.len() == nb_species")
        d = self._diff(differences)
        check(lib.rebop_network_add_reaction_lma(self._h, float(k), ptr(e, C.c_uint32), ptr(d, C.c_int64)))

    def add_reaction_lma_sparse(self, k: float, terms, differences) -> None:
        idx = np.ascontiguousarray([t[0] for t in terms] or [0], dtype=np.uint32)
        ex = np.ascontiguousarray([t[1] for t in terms] or [0], dtype=np.uint32)
        d = self._diff(differences)
        check(lib.rebop_network_add_reaction_lma_sparse(self._h, float(k), ptr(idx, C.c_uint32), ptr(ex, C.c_uint32),
                                                        len(terms), ptr(d, C.c_int64)))

    def add_reaction_expr(self, program, differences) -> None:
        program = list(program)
        d = self._diff(differences)
        check(lib.rebop_network_add_reaction_expr(self._h, make_program(program), len(program), ptr(d, C.c_int64)))

    @property
    def nb_reactions(self) -> int:
        n = C.c_uint32()
        check(lib.rebop_network_nb_reactions(self._h, C.byref(n)))
        return n.value

    def codegen(self) -> str:
        need = C.c_size_t()
        check(lib.rebop_network_codegen(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib.rebop_network_codegen(self._h, buf, need.value, None))
        return buf.value.decode("utf-8")

    def codegen_pdm(self) -> str:
        """Source of the partial-propensity kernel (KERNEL_PDM) for this network."""
        need = C.c_size_t()
        check(lib.rebop_network_codegen_pdm(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib.rebop_network_codegen_pdm(self._h, buf, need.value, None))
        return buf.value.decode()

    def jit_cubin_pdm(self) -> bytes:
        need = C.c_size_t()
        check(lib.rebop_network_jit_cubin_pdm(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib.rebop_network_jit_cubin_pdm(self._h, buf, need.value, None))
        return buf.raw[:need.value]

    def jit_cubin(self) -> bytes:
        need = C.c_size_t()
        check(lib.rebop_network_jit_cubin(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib.rebop_network_jit_cubin(self._h, buf, need.value, None))
        return buf.raw


def _string_out(fn, *args) -> str:
    need = C.c_size_t()
    check(fn(*args, None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    check(fn(*args, buf, need.value, None))
    return buf.value.decode("utf-8")


def prebuilt_systems() -> list:
    """Names of the systems whose kernels were compiled into the library at build time."""
    return [_string_out(lib.rebop_b200_prebuilt_name, i) for i in range(lib.rebop_b200_prebuilt_count())]


class System:
    """Owning wrapper of a rebop_system handle: parsed `define_system!` text."""

    def __init__(self, dsl_text: str):
        h = C.c_void_p()
        check(lib.rebop_system_parse(dsl_text.encode("utf-8"), C.byref(h)))
        self._h = h
        self.name = _string_out(lib.rebop_system_name, h)
        np_, ns, nr = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib.rebop_system_counts(h, C.byref(np_), C.byref(ns), C.byref(nr)))
        self.params = [_string_out(lib.rebop_system_param_name, h, i) for i in range(np_.value)]
        self.species = [_string_out(lib.rebop_system_species_name, h, i) for i in range(ns.value)]
        self.reactions = [_string_out(lib.rebop_system_reaction_name, h, i) for i in range(nr.value)]

    def __del__(self):
        if getattr(self, "_h", None):
            lib.rebop_system_destroy(self._h)
            self._h = None

    def rates(self, params) -> np.ndarray:
        """Value of every reaction's rate expression for these parameter values."""
        pv = np.ascontiguousarray(list(params) or [0.0], dtype=np.float64)
        out = np.empty(max(1, len(self.reactions)), dtype=np.float64)
        check(lib.rebop_system_rates(self._h, ptr(pv, C.c_double), len(list(params)), ptr(out, C.c_double)))
        return out[:len(self.reactions)]

    def network(self, params) -> Network:
        """Name::with_parameters(...) -> network in define_system! arithmetic."""
        pv = np.ascontiguousarray(list(params) or [0.0], dtype=np.float64)
        h = C.c_void_p()
        check(lib.rebop_system_network(self._h, ptr(pv, C.c_double), len(list(params)), C.byref(h)))
        return Network(len(self.species), ARITH_MACRO, _handle=h)


class Batch:
    """Owning wrapper of a rebop_batch handle: N trajectories resident on one GPU."""

    def __init__(self, net: Network, n_traj: int, x0, seeds=None, seed_base: int = 0, device: int = 0,
                 kernel: int = KERNEL_AUTO):
        x0a = np.ascontiguousarray(x0, dtype=np.int64)
        per_traj = 1 if x0a.ndim == 2 else 0
        if per_traj and x0a.shape != (n_traj, net.n_species):
            raise RebopError(ERR_INVALID, "x0 must be [n_species] or [n_traj][n_species]")
        if not per_traj and x0a.shape != (net.n_species,):
            raise RebopError(ERR_INVALID, "assertion failed: species.len() == nb_species")
        sp = None
        if seeds is not None:
            self._seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
            if self._seeds.shape != (n_traj,):
                raise RebopError(ERR_INVALID, "seeds must have one entry per trajectory")
            sp = ptr(self._seeds, C.c_uint64)
        h = C.c_void_p()
        check(lib.rebop_batch_create(net._h, int(device), int(n_traj), ptr(x0a, C.c_int64) if x0a.size else None, per_traj, sp,
                                     C.c_uint64(int(seed_base) & (2**64 - 1)), C.byref(h)))
        self._h = h
        self.net = net
        self.n_traj = int(n_traj)
        self.rows = 0
        self.n_save = 0
        if kernel != KERNEL_AUTO:
            self.set_kernel(kernel)

    def __del__(self):
        self.close()

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.rebop_batch_destroy(self._h)
            self._h = None

    def set_kernel(self, kind: int) -> None:
        check(lib.rebop_batch_set_kernel(self._h, int(kind)))

    @property
    def kernel_used(self) -> int:
        k = C.c_int()
        check(lib.rebop_batch_get_kernel(self._h, C.byref(k)))
        return k.value

    def set_rates(self, k) -> None:
        k = np.ascontiguousarray(k, dtype=np.float64)
        check(lib.rebop_batch_set_rates(self._h, ptr(k, C.c_double), k.size))

    def set_schedule(self, schedule: int) -> None:
        """0 auto, 1 static, 2 lanes claim trajectories (sparse samples), 3 the same for dense samples
        (see rebop_batch_set_schedule)."""
        check(lib.rebop_batch_set_schedule(self._h, int(schedule)))

    @property
    def schedule_used(self) -> int:
        v = C.c_int()
        check(lib.rebop_batch_get_schedule(self._h, C.byref(v)))
        return v.value

    def set_max_iters(self, n: int) -> None:
        check(lib.rebop_batch_set_max_iters(self._h, int(n)))

    def seed(self, seeds=None, seed_base: int = 0) -> None:
        if seeds is None:
            check(lib.rebop_batch_seed(self._h, None, C.c_uint64(int(seed_base))))
        else:
            s = np.ascontiguousarray(seeds, dtype=np.uint64)
            check(lib.rebop_batch_seed(self._h, ptr(s, C.c_uint64), 0))

    def advance_until(self, tmax: float) -> None:
        check(lib.rebop_batch_advance_until(self._h, float(tmax)))

    def advance_one_reaction(self) -> None:
        check(lib.rebop_batch_advance_one_reaction(self._h))

    def set_sample_dtype(self, dtype) -> None:
        """Sample type of run_grid results: np.int16, np.int32 (default) or np.int64."""
        check(lib.rebop_batch_set_sample_dtype(self._h, int(np.dtype(dtype).itemsize)))

    @property
    def sample_dtype(self):
        v = C.c_int()
        check(lib.rebop_batch_get_sample_dtype(self._h, C.byref(v)))
        return np.dtype(SAMPLE_DTYPES[v.value])

    def run_grid(self, tmax: float, nb_steps: int, save_idx=None, host_out: np.ndarray | None = None) -> None:
        n_save = self.net.n_species if save_idx is None else len(save_idx)
        sp = None
        if save_idx is not None:
            sv = np.ascontiguousarray(save_idx, dtype=np.uint32)
            sp = ptr(sv, C.c_uint32)
        hp = None
        if host_out is not None:
            assert host_out.dtype == self.sample_dtype and host_out.flags.c_contiguous
            assert host_out.size == (nb_steps + 1) * n_save * self.n_traj
            hp = C.c_void_p(host_out.ctypes.data)
        self.rows, self.n_save = (nb_steps + 1) * n_save, n_save
        check(lib.rebop_batch_run_grid_typed(self._h, float(tmax), int(nb_steps), sp, n_save, hp))

    def samples(self, dtype=None) -> np.ndarray:
        """Samples of the last run_grid as [nb_steps+1][n_save][n_traj]: in the batch's sample type by default, or
        as int32 / int64 (converted on the device)."""
        steps = self.rows // self.n_save if self.n_save else 0
        dtype = self.sample_dtype if dtype is None else np.dtype(dtype)
        out = np.empty((steps, self.n_save, self.n_traj), dtype=dtype)
        if out.size:
            if dtype == self.sample_dtype:
                check(lib.rebop_batch_samples_host(self._h, C.c_void_p(out.ctypes.data)))
            elif dtype == np.int32:
                check(lib.rebop_batch_samples_host_i32(self._h, ptr(out, C.c_int32)))
            elif dtype == np.int64:
                check(lib.rebop_batch_samples_host_i64(self._h, ptr(out, C.c_int64)))
            else:
                raise TypeError("dtype must be the batch's sample type, int32 or int64")
        return out

    def run_events(self, tmax: float, save_idx=None):
        """nb_steps = 0: one row per applied reaction.  Returns (offsets [n+1], times [rows], samples [n_save][rows]);
        trajectory n owns rows offsets[n]:offsets[n+1]."""
        n_save = self.net.n_species if save_idx is None else len(save_idx)
        sp = None
        if save_idx is not None:
            sv = np.ascontiguousarray(save_idx, dtype=np.uint32)
            sp = ptr(sv, C.c_uint32)
        check(lib.rebop_batch_run_events(self._h, float(tmax), sp, n_save))
        total, ns = C.c_uint64(), C.c_uint32()
        check(lib.rebop_batch_events_log_size(self._h, C.byref(total), C.byref(ns)))
        offsets = np.empty(self.n_traj + 1, dtype=np.uint64)
        times = np.empty(total.value, dtype=np.float64)
        samples = np.empty((ns.value, total.value), dtype=np.int32)
        check(lib.rebop_batch_events_log_host(self._h, ptr(offsets, C.c_uint64), ptr(times, C.c_double) if times.size else None,
                                              ptr(samples, C.c_int32) if samples.size else None))
        self.rows = 0
        return offsets, times, samples

    def samples_into(self, out: np.ndarray, first: int) -> None:
        """Write the samples into columns [first, first + n_traj) of a C-contiguous int32 array
        [nb_steps+1][n_save][n_total] (no intermediate copy)."""
        assert out.dtype == self.sample_dtype and out.flags.c_contiguous and out.ndim == 3
        assert out.shape[0] * out.shape[1] == self.rows and first + self.n_traj <= out.shape[2]
        if out.size:
            base = out.ctypes.data + out.dtype.itemsize * first
            check(lib.rebop_batch_samples_host_strided(self._h, C.c_void_p(base), out.shape[2]))

    def samples_device(self):
        """(device pointer, leading dimension in elements, rows) of the last run_grid's samples."""
        p, ld, rows = C.c_void_p(), C.c_size_t(), C.c_uint32()
        check(lib.rebop_batch_samples_device(self._h, C.byref(p), C.byref(ld), C.byref(rows)))
        return p.value, ld.value, rows.value

    def sample_sums(self):
        sums = np.zeros(self.rows, dtype=np.int64)
        sq = np.zeros(self.rows, dtype=np.uint64)
        check(lib.rebop_batch_sample_sums(self._h, ptr(sums, C.c_int64), ptr(sq, C.c_uint64)))
        return sums, sq

    def sample_sums_device(self):
        p, rows = C.c_void_p(), C.c_uint32()
        check(lib.rebop_batch_sample_sums_device(self._h, C.byref(p), C.byref(rows)))
        return p.value, rows.value

    def species(self) -> np.ndarray:
        out = np.empty((self.n_traj, self.net.n_species), dtype=np.int64)
        if out.size:
            check(lib.rebop_batch_get_species(self._h, ptr(out, C.c_int64)))
        return out

    def set_species(self, species) -> None:
        a = np.ascontiguousarray(species, dtype=np.int64)
        check(lib.rebop_batch_set_species(self._h, ptr(a, C.c_int64), 1 if a.ndim == 2 else 0))

    def times(self) -> np.ndarray:
        out = np.empty(self.n_traj, dtype=np.float64)
        check(lib.rebop_batch_get_time(self._h, ptr(out, C.c_double)))
        return out

    def set_time(self, t: float) -> None:
        check(lib.rebop_batch_set_time(self._h, float(t)))

    def events(self):
        tot, last = C.c_uint64(), C.c_uint64()
        check(lib.rebop_batch_events(self._h, C.byref(tot), C.byref(last)))
        return tot.value, last.value

    @property
    def lane_slots(self) -> int:
        """32 x loop iterations of every warp in the last launch (events / lane_slots = SIMT lane efficiency)."""
        v = C.c_uint64()
        check(lib.rebop_batch_lane_slots(self._h, C.byref(v)))
        return v.value

    @property
    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(lib.rebop_batch_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    @property
    def last_finish_ms(self) -> float:
        """Device time of the kernels that brought the last run_grid's samples into the result layout."""
        ms = C.c_float()
        check(lib.rebop_batch_last_finish_ms(self._h, C.byref(ms)))
        return ms.value

    def synchronize(self) -> None:
        check(lib.rebop_batch_synchronize(self._h))

    @property
    def stream(self) -> int:
        """cudaStream_t the batch issues its work on (as an integer handle)."""
        p = C.c_void_p()
        check(lib.rebop_batch_get_stream(self._h, C.byref(p)))
        return p.value or 0

    def set_stream(self, stream: int | None) -> None:
        check(lib.rebop_batch_set_stream(self._h, C.c_void_p(stream or None)))


class Ensemble:
    """Owning wrapper of a rebop_ensemble handle: N trajectories sharded over several GPUs of this node (one process,
    one batch and one host worker thread per device; NCCL only for the ensemble statistics)."""

    def __init__(self, net: Network, n_traj: int, x0, devices, seeds=None, seed_base: int = 0, kernel: int = KERNEL_AUTO,
                 dtype=np.int32):
        x0a = np.ascontiguousarray(x0, dtype=np.int64)
        per_traj = 1 if x0a.ndim == 2 else 0
        if per_traj and x0a.shape != (n_traj, net.n_species):
            raise RebopError(ERR_INVALID, "x0 must be [n_species] or [n_traj][n_species]")
        if not per_traj and x0a.shape != (net.n_species,):
            raise RebopError(ERR_INVALID, "assertion failed: species.len() == nb_species")
        sp = None
        if seeds is not None:
            self._seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
            if self._seeds.shape != (n_traj,):
                raise RebopError(ERR_INVALID, "seeds must have one entry per trajectory")
            sp = ptr(self._seeds, C.c_uint64)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        check(lib.rebop_ensemble_create(net._h, devs, len(devices), int(n_traj), ptr(x0a, C.c_int64) if x0a.size else None, per_traj,
                                        sp, C.c_uint64(int(seed_base) & (2**64 - 1)), C.byref(h)))
        self._h = h
        self.net = net
        self.n_traj = int(n_traj)
        self.devices = [int(d) for d in devices]
        self.rows = 0
        self.n_save = 0
        self.dtype = np.dtype(dtype)
        if kernel != KERNEL_AUTO:
            check(lib.rebop_ensemble_set_kernel(self._h, int(kernel)))
        if self.dtype != np.int32:
            check(lib.rebop_ensemble_set_sample_dtype(self._h, int(self.dtype.itemsize)))

    def __del__(self):
        self.close()

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.rebop_ensemble_destroy(self._h)
            self._h = None

    def shards(self):
        """[(device, first, count)] -- the contiguous trajectory range of every device."""
        n = C.c_int()
        check(lib.rebop_ensemble_shards(self._h, C.byref(n)))
        out = []
        for g in range(n.value):
            dev, first, count = C.c_int(), C.c_size_t(), C.c_size_t()
            check(lib.rebop_ensemble_shard(self._h, g, None, C.byref(dev), C.byref(first), C.byref(count)))
            out.append((dev.value, first.value, count.value))
        return out

    def set_schedule(self, schedule: int) -> None:
        check(lib.rebop_ensemble_set_schedule(self._h, int(schedule)))

    def set_max_iters(self, n: int) -> None:
        check(lib.rebop_ensemble_set_max_iters(self._h, int(n)))

    def set_time(self, t: float) -> None:
        check(lib.rebop_ensemble_set_time(self._h, float(t)))

    def set_species(self, species) -> None:
        a = np.ascontiguousarray(species, dtype=np.int64)
        check(lib.rebop_ensemble_set_species(self._h, ptr(a, C.c_int64), 1 if a.ndim == 2 else 0))

    def seed(self, seeds=None, seed_base: int = 0) -> None:
        if seeds is None:
            check(lib.rebop_ensemble_seed(self._h, None, C.c_uint64(int(seed_base))))
        else:
            s = np.ascontiguousarray(seeds, dtype=np.uint64)
            check(lib.rebop_ensemble_seed(self._h, ptr(s, C.c_uint64), 0))

    def advance_until(self, tmax: float) -> None:
        check(lib.rebop_ensemble_advance_until(self._h, float(tmax)))

    def run_grid(self, tmax: float, nb_steps: int, save_idx=None, host_out: np.ndarray | None = None) -> None:
        n_save = self.net.n_species if save_idx is None else len(save_idx)
        sp = None
        if save_idx is not None:
            sv = np.ascontiguousarray(save_idx, dtype=np.uint32)
            sp = ptr(sv, C.c_uint32)
        hp = None
        if host_out is not None:
            assert host_out.dtype == self.dtype and host_out.flags.c_contiguous
            assert host_out.size == (nb_steps + 1) * n_save * self.n_traj
            hp = C.c_void_p(host_out.ctypes.data)
        self.rows, self.n_save = (nb_steps + 1) * n_save, n_save
        check(lib.rebop_ensemble_run_grid(self._h, float(tmax), int(nb_steps), sp, n_save, hp))

    def samples(self) -> np.ndarray:
        steps = self.rows // self.n_save if self.n_save else 0
        out = np.empty((steps, self.n_save, self.n_traj), dtype=self.dtype)
        if out.size:
            check(lib.rebop_ensemble_samples_host(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def stats(self):
        """(mean, unbiased variance) per (step, saved species): K4 per device, NCCL all-reduce, device finalisation."""
        mean = np.zeros(self.rows, dtype=np.float64)
        var = np.zeros(self.rows, dtype=np.float64)
        check(lib.rebop_ensemble_stats(self._h, ptr(mean, C.c_double), ptr(var, C.c_double)))
        steps = self.rows // self.n_save if self.n_save else 0
        return mean.reshape(steps, self.n_save), var.reshape(steps, self.n_save)

    def sums(self):
        s1 = np.zeros(self.rows, dtype=np.int64)
        s2 = np.zeros(self.rows, dtype=np.uint64)
        check(lib.rebop_ensemble_sums(self._h, ptr(s1, C.c_int64), ptr(s2, C.c_uint64)))
        return s1, s2

    def events(self):
        tot, last = C.c_uint64(), C.c_uint64()
        check(lib.rebop_ensemble_events(self._h, C.byref(tot), C.byref(last)))
        return tot.value, last.value

    @property
    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(lib.rebop_ensemble_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value


def _find_nccl() -> None:
    """Point the library's dlopen at the NCCL that ships with the nvidia-nccl wheel when the system has none on the
    loader path (the C ABI honours REBOP_B200_NCCL_LIB)."""
    if os.environ.get("REBOP_B200_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["REBOP_B200_NCCL_LIB"] = cand
    except (ImportError, ValueError):
        pass


_find_nccl()


def kernel_launches() -> int:
    return int(lib.rebop_b200_kernel_launches())


class PinnedBuffer:
    """Page-locked host array (rebop_b200_host_alloc) exposed as a numpy view."""

    def __init__(self, shape, dtype=np.int32):
        self.shape = tuple(int(v) for v in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib.rebop_b200_host_alloc(nbytes, C.byref(p)))
        self._p = p
        buf = (C.c_char * max(1, nbytes)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=nbytes // self.dtype.itemsize).reshape(self.shape)

    def close(self) -> None:
        if getattr(self, "_p", None):
            self.array = None
            if lib is not None:  # interpreter shutdown may have dropped the module globals already
                lib.rebop_b200_host_free(self._p)
            self._p = None

    def __del__(self):
        self.close()


def measure_fp64_rate(device: int = 0):
    ops, mhz = C.c_double(), C.c_double()
    check(lib.rebop_b200_measure_fp64_rate(int(device), C.byref(ops), C.byref(mhz)))
    return ops.value, mhz.value
