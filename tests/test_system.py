"""define_system! on the GPU engine: DSL parsing, build-time kernels, the generated-struct API.

Models and assertions follow the reference's macro tests (src/gillespie_macro.rs:173-254) and its
macro users (examples/dimers.rs, benchmarks/benches/vilar/vilar.rs).
"""
import glob
import os

import numpy as np
import pytest

import rebop_b200
from rebop_b200 import models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SYSTEMS = {os.path.basename(p)[:-5]: open(p, encoding="utf-8").read()
           for p in glob.glob(os.path.join(ROOT, "rebop_b200", "systems", "*.rsys"))}

BIRTH_DEATH = """
    r_birth r_death;
    BirthDeath { A }
    birth:      => A    @ r_birth
    death:  A   =>      @ r_death
"""


# ---------------------------------------------------------------------------- host side (no GPU)
def test_parse_vilar_names():
    v = rebop_b200.define_system(SYSTEMS["vilar"])
    assert v.name == "Vilar"
    assert v.species == ["Da", "Dr", "Dpa", "Dpr", "Ma", "Mr", "A", "R", "C"]
    assert v.params[:3] == ["αA", "αpA", "αR"] and len(v.params) == 15
    assert len(v.reactions) == 16 and v.reactions[0] == "r_activation_a"


@pytest.mark.parametrize("text", [
    "a b Foo { X }",                                   # missing ';'
    "a; Foo { X } r: X => Y @ a",                      # unknown species
    "a; Foo { X } r: X => @ b",                        # unknown parameter
    "a; Foo { X } r: X => @ X",                        # species in a rate: stale-snapshot quirk not reproduced
    "a; Foo { X } r X => @ a",                         # missing ':'
    "a; Foo { X } r: 1.5 X => @ a",                    # non-integer coefficient
    "a; Foo { X } r: X -> @ a",                        # wrong arrow
    "a; Foo { X } r: X => @ (a",                       # unbalanced parenthesis
])
def test_parse_errors(ffi, text):
    with pytest.raises(ffi.RebopError) as e:
        rebop_b200.define_system(text)
    assert e.value.status == ffi.ERR_PARSE


def test_rate_expressions_and_lowering(ffi):
    s = ffi.System("k1 k2; Foo { A, B } fwd: 2 A + B => 3 B @ 2.0 * k1 / (k2 + 1.) - 1e-3\n back: B => A @ -k2")
    np.testing.assert_array_equal(s.rates([3.0, 0.5]), [2.0 * 3.0 / (0.5 + 1.0) - 1e-3, -0.5])
    src = s.network([3.0, 0.5]).codegen()
    # macro arithmetic: integer falling factorial of the `2 A` term, then B; jump (-2, +2) packed as 0x02fe
    assert "define_system! arithmetic" in src and "__dmul_rn(d0, __dsub_rn(d0, 1.0))" in src and "0x000102fe" in src  # -2, +2, event-count lane


def test_with_parameters_arity():
    d = rebop_b200.define_system(SYSTEMS["dimers"])
    with pytest.raises(TypeError):
        d.with_parameters(1.0, 2.0)


def test_build_time_kernels_are_registered(ffi):
    assert set(ffi.prebuilt_systems()) >= {"Vilar", "Dimers", "SIR", "BirthDeath"}
    for name, text in SYSTEMS.items():
        s = ffi.System(text)
        assert s.network([1.0] * len(s.params)).has_prebuilt, name
        assert s.network([0.5 + i for i in range(len(s.params))]).has_prebuilt  # whatever the parameter values


def test_models_match_their_systems(ffi):
    """The plain-data benchmark models lower to the same specialised source as the DSL texts."""
    for name in ("vilar", "dimers", "sir"):
        m = models.MODELS[name]()
        a = ffi.System(SYSTEMS[name]).network(m["params"]).codegen()
        assert a == models.build_network(m, ffi.ARITH_MACRO).codegen()
    # function-API arithmetic is a different kernel: not the build-time one
    assert not models.build_network(models.vilar(), ffi.ARITH_API).has_prebuilt
    # a system that is not under rebop_b200/systems/ goes through NVRTC
    assert not ffi.System("k; Other { X, Y } r: X + Y => @ k").network([1.0]).has_prebuilt


def test_struct_fields():
    d = rebop_b200.define_system(SYSTEMS["dimers"]).new()
    assert d.gene == 0 and d.t == 0.0 and np.isnan(d.rtx)
    d.gene = 1
    d.rtx = 25.0
    assert d.gene == 1 and d.rtx == 25.0
    with pytest.raises(AttributeError, match="no field `nope` on type `Dimers`"):
        d.nope = 1
    with pytest.raises(AttributeError):
        d.nope


# ---------------------------------------------------------------------------- on the GPU
@pytest.mark.gpu
def test_sir_conservation(gpu):
    """src/gillespie_macro.rs:175-191."""
    sir = rebop_b200.define_system(SYSTEMS["sir"]).new(n_trajectories=200)
    sir.r1 = 0.1 / 10000.0
    sir.r2 = 0.01
    sir.S = 9999
    sir.I = 1
    sir.seed(1)
    sir.advance_until(1000.0)
    assert np.all(sir.S + sir.I + sir.R == 10000)
    assert np.all(sir.t == 1000.0)


@pytest.mark.gpu
def test_dimers_bounds_and_kernel(gpu, ffi):
    """src/gillespie_macro.rs:192-208 / the doctest at :27-48."""
    dimers = rebop_b200.define_system(SYSTEMS["dimers"]).with_parameters(25.0, 1000.0, 0.001, 0.1, 1.0)
    dimers.gene = 1
    dimers.advance_until(1.0)  # unseeded: OS entropy, like Dimers::new()
    assert dimers.t == 1.0 and dimers.gene == 1
    assert 1000 < dimers.dimer < 10000
    assert dimers.kernel_used == ffi.KERNEL_PREBUILT


@pytest.mark.gpu
def test_birth_death(gpu, ffi):
    """src/gillespie_macro.rs:209-223; not a build-time system under another name => NVRTC."""
    bd = rebop_b200.define_system(BIRTH_DEATH.replace("BirthDeath", "BirthDeath2")).new(n_trajectories=64)
    bd.r_birth = 10.0
    bd.r_death = 0.1
    bd.seed(5)
    bd.advance_until(100.0)
    assert np.all((50 < bd.A) & (bd.A < 200))
    assert bd.kernel_used == ffi.KERNEL_PREBUILT  # same structure as systems/birth_death.rsys: the name is not part of the key


@pytest.mark.gpu
def test_forgotten_parameter_freezes(gpu):
    """src/gillespie_macro.rs:224-238: a NaN parameter => no reaction, t == tmax."""
    bd = rebop_b200.define_system(BIRTH_DEATH).new()
    bd.r_birth = 10.0
    bd.advance_until(100.0)
    assert bd.t == 100.0 and bd.A == 0


@pytest.mark.gpu
def test_no_reactions(gpu):
    """src/gillespie_macro.rs:239-253."""
    f = rebop_b200.define_system("; FooBarBuz { Foo, Bar, Buz }").new()
    f.Foo = 42
    f.Bar = 1337
    f.advance_until(1e20)
    assert f.t == 1e20 and (f.Foo, f.Bar, f.Buz) == (42, 1337, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name,tmax,steps", [("vilar", 10.0, 5), ("dimers", 1.0, 2), ("sir", 250.0, 5)])
def test_bit_exact_vs_oracle_macro_form(gpu, ffi, oracle, name, tmax, steps):
    """Repeated advance_until on the generated struct == the oracle's hand-expanded macro code."""
    m = models.MODELS[name]()
    n = 128
    seeds = models.seeds_sequence(n, 77)
    ref, _, ref_events = oracle.run_batch_macro(name, m["params"], m["x0"], seeds, tmax, steps)
    st = rebop_b200.define_system(SYSTEMS[name]).with_parameters(*m["params"], n_trajectories=n)
    for sp, v in zip(st._system.species, m["x0"]):
        setattr(st, sp, v)
    st.seed(77)
    for i in range(steps + 1):
        st.advance_until(tmax * i / steps)
        got = np.stack([getattr(st, sp) for sp in st._system.species])
        np.testing.assert_array_equal(got, ref[i])
    assert st.events == ref_events
    assert st.kernel_used == ffi.KERNEL_PREBUILT


@pytest.mark.gpu
def test_prebuilt_equals_nvrtc_equals_table(gpu, ffi):
    m = models.vilar()
    seeds = models.seeds_sequence(96, 3)
    outs = []
    for kernel in (ffi.KERNEL_PREBUILT, ffi.KERNEL_NVRTC, ffi.KERNEL_TABLE):
        b = ffi.Batch(models.build_network(m, ffi.ARITH_MACRO), len(seeds), m["x0"], seeds=seeds, kernel=kernel)
        b.run_grid(10.0, 10)
        assert b.kernel_used == kernel
        outs.append(b.samples())
        b.close()
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], outs[2])


@pytest.mark.gpu
def test_parameters_can_change_between_calls(gpu, ffi, oracle):
    """Parameters are plain struct fields; changing one keeps states, times and random streams."""
    bd = rebop_b200.define_system(BIRTH_DEATH).with_parameters(10.0, 0.1, n_trajectories=32)
    bd.seed(9)
    bd.advance_until(5.0)
    a_mid = bd.A.copy()
    bd.r_birth = 0.0          # only deaths from now on
    bd.advance_until(50.0)
    assert np.all(bd.A <= a_mid) and np.all(bd.t == 50.0)


def test_sysgen_tool_and_nvcc_cross_compile(tmp_path):
    """The build.rs analogue end to end, without a GPU: DSL text -> rebop_sysgen -> nvcc (sm_100a) -> an object
    with the five entry points (time grid in its three schedules, event log count/write) and the registration."""
    import shutil
    import subprocess
    csrc = os.path.join(ROOT, "rebop_b200", "csrc")
    tool = os.path.join(csrc, "build", "rebop_sysgen")
    if not os.path.exists(tool):
        subprocess.check_call(["make", "-C", csrc, "build/rebop_sysgen"])
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    rsys = tmp_path / "schloegl.rsys"
    rsys.write_text("k1 k2 k3 k4;\nSchloegl { A, B, X }\n r1: 2 X + A => 3 X @ k1 / 2.\n r2: 3 X => 2 X + A @ k2\n"
                    " r3: B => X @ k3\n r4: X => B @ k4\n", encoding="utf-8")
    cu = tmp_path / "schloegl.cu"
    subprocess.check_call([tool, str(rsys), "-o", str(cu)])
    text = cu.read_text(encoding="utf-8")
    for suffix in ("", "_dyn", "_dns", "_evc", "_evw"):
        assert f"rb_ssa_sys_Schloegl{suffix}(const __grid_constant__ SsaRunParams p)" in text
    assert "rb_register_prebuilt" in text
    obj = tmp_path / "schloegl.o"
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
                           "-Xcompiler", "-fPIC", "-I", csrc, "-c", str(cu), "-o", str(obj)])
    sass = subprocess.check_output(["cuobjdump", "-sass", str(obj)], text=True)
    assert sass.count("Function : rb_ssa_sys_Schloegl") == 5
    assert "DFMA" in sass and "IDP.4A" in sass  # IEEE divide and packed stoichiometry update are in there
    # a malformed system is refused with the parser's message
    bad = tmp_path / "bad.rsys"
    bad.write_text("k; Foo { X } r: X => Y @ k", encoding="utf-8")
    res = subprocess.run([tool, str(bad), "-o", str(tmp_path / "bad.cu")], capture_output=True, text=True)
    assert res.returncode != 0 and "no field `Y` on type `Foo`" in res.stderr
