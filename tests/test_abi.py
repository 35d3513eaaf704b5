"""Host side of the C ABI (no GPU needed): exports, validation, expression front end, code generation."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from rebop_b200 import models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rebop_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rebop_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ffi):
    declared = declared_symbols()
    assert len(declared) >= 40
    out = subprocess.check_output(["nm", "-D", "--defined-only", ffi.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    # the ctypes binding covers exactly the header
    assert sorted(ffi.SIGNATURES) == declared


def test_no_oracle_or_cpu_fallback_in_product():
    """The product never touches oracle/: only tests/, smoke() and bench.py's CPU legs may."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rebop_b200")):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "ssa_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_version_and_device_count(ffi):
    assert "rebop" in ffi.lib.rebop_b200_version().decode()
    assert ffi.device_count() >= 0


def test_compute_fails_loudly_without_gpu(ffi):
    if ffi.device_count() > 0:
        pytest.skip("a GPU is visible")
    net = models.build_network(models.sir())
    with pytest.raises(ffi.RebopError) as e:
        ffi.Batch(net, 8, [999, 1, 0])
    assert e.value.status == ffi.ERR_CUDA
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(ffi.RebopError):
        ffi.measure_fp64_rate(0)
    with pytest.raises(ffi.RebopError):
        ffi.PinnedBuffer((4,))


def test_add_reaction_validation(ffi):
    """src/gillespie.rs:225-244: shape and index asserts; test issue85_oob (:496-502)."""
    net = ffi.Network(2)
    with pytest.raises(ffi.RebopError) as e:
        net.add_reaction_lma_sparse(1.0, [(2, 1)], [0, 0])  # index out of range
    assert e.value.status == ffi.ERR_OUT_OF_RANGE
    with pytest.raises(ffi.RebopError) as e:
        net.add_reaction_lma(1.0, [1, 0, 0], [0, 0])
    assert e.value.status == ffi.ERR_INVALID
    with pytest.raises(ffi.RebopError) as e:
        net.add_reaction_lma(1.0, [1, 0], [0, 0, 1])
    assert e.value.status == ffi.ERR_INVALID
    net.add_reaction_lma(1.0, [1, 0], [-1, 1])
    net.add_reaction_lma_sparse(2.0, [(0, 2)], [-2, 1])
    assert net.nb_reactions == 2
    with pytest.raises(ffi.RebopError) as e:
        ffi.Network(2, arith=7)
    assert e.value.status == ffi.ERR_INVALID


def test_expression_program_validation(ffi):
    net = ffi.Network(2)
    net.add_reaction_expr([("const", 0, 1.5), ("species", 1, 0), ("mul", 0, 0)], [0, -1])
    with pytest.raises(ffi.RebopError) as e:
        net.add_reaction_expr([("species", 5, 0)], [0, 0])
    assert e.value.status == ffi.ERR_OUT_OF_RANGE
    with pytest.raises(ffi.RebopError) as e:
        net.add_reaction_expr([("const", 0, 1.0), ("add", 0, 0)], [0, 0])
    assert e.value.status == ffi.ERR_INVALID
    with pytest.raises(ffi.RebopError) as e:
        net.add_reaction_expr([("const", 0, 1.0), ("const", 0, 1.0)], [0, 0])
    assert e.value.status == ffi.ERR_INVALID


# ---- PExpr: the reference's parser tests (src/expr.rs:275-447) ---------------------------------
PARSE_CASES = [
    ("3", "3"), (".23", "0.23"), ("1.23", "1.23"), ("2e-3", "0.002"), ("2.04e3", "2040"), ("-2.04e3", "-2040"),
    ("+2.04e3", "2040"),
    ("a", "a"), ("_", "_"), ("exp", "exp"), ("max", "max"), ("min", "min"), ("PhoP3", "PhoP3"),
    ("Mono1_Mono2_", "Mono1_Mono2_"),
    ("max(x,y)", "max(x, y)"), ("min(x,y)", "min(x, y)"), ("exp(3)", "exp(3)"),
    ("3^4", "(3 ^ 4)"), ("3 ^4", "(3 ^ 4)"), ("3^ 4", "(3 ^ 4)"), ("3 ^ 4", "(3 ^ 4)"),
    ("3*4", "(3 * 4)"), ("3 *4", "(3 * 4)"), ("3* 4", "(3 * 4)"), ("3 * 4", "(3 * 4)"),
    ("3*4/2", "((3 * 4) / 2)"), ("3*4 /2", "((3 * 4) / 2)"), ("3*4/ 2", "((3 * 4) / 2)"), ("3*4 / 2", "((3 * 4) / 2)"),
    ("-A", "(-A)"), ("- A", "(-A)"),
    ("3+4", "(3 + 4)"), ("3-4+1", "((3 - 4) + 1)"), ("3-4-1", "((3 - 4) - 1)"),
    ("1.20 * A*B", "((1.2 * A) * B)"), ("1.20*Sugar / (3.5+Sugar)", "((1.2 * Sugar) / (3.5 + Sugar))"),
    ("-Sugar / (3.5+Sugar)", "(-(Sugar / (3.5 + Sugar)))"), ("A + -Sugar", "(A + (-Sugar))"),
    ("A * -Sugar", "(A * (-Sugar))"), ("A + -Sugar / (Kd + Sugar)", "(A + (-(Sugar / (Kd + Sugar))))"),
    ("A^-B", "(A ^ (-B))"),
    # names that look like floats or functions stay variables (src/expr.rs:409-447)
    ("inf", "inf"), ("infect", "infect"), ("nan", "nan"), ("nanny", "nanny"), ("E", "E"), ("e", "e"),
    ("explicit", "explicit"), ("maximum", "maximum"), ("minimum", "minimum"),
]


@pytest.mark.parametrize("text,shown", PARSE_CASES)
def test_pexpr_parse_and_format(ffi, text, shown):
    assert str(ffi.PExpr(text)) == shown


@pytest.mark.parametrize("text", ["+", "1+", "", "(", "a b", "max(1)", "2 ** 3", "1 + * 2"])
def test_pexpr_parse_failures(ffi, text):
    with pytest.raises(ffi.RebopError) as e:
        ffi.PExpr(text)
    assert e.value.status == ffi.ERR_PARSE


def test_pexpr_lowering_matches_reference_conversion(ffi):
    """src/expr.rs:462-497 test_conversion: the post-order program of the expected Expr tree."""
    e = ffi.PExpr("1.21 * C + B - A / D ^ E * (F + exp(D))")
    prog = e.lower(list("ABCDEF"), {})
    O = ffi.OPCODES
    want = [(O["const"], 1.21), (O["species"], 2), (O["mul"], None), (O["species"], 1), (O["add"], None),
            (O["species"], 0), (O["species"], 3), (O["species"], 4), (O["pow"], None), (O["div"], None),
            (O["species"], 5), (O["species"], 3), (O["exp"], None), (O["add"], None), (O["mul"], None), (O["sub"], None)]
    assert len(prog) == len(want)
    for (op, idx, val), (wop, warg) in zip(prog, want):
        assert op == wop
        if wop == O["const"]:
            assert val == warg
        if wop == O["species"]:
            assert idx == warg


def test_pexpr_lowering_names(ffi):
    """src/expr.rs:58-75: species shadow parameters; unknown names are an error with the reference's text."""
    e = ffi.PExpr("V * A / (Km + A)")
    prog = e.lower(["A", "P"], {"V": 1.0, "Km": 20.0})
    assert [p[0] for p in prog] == [ffi.OPCODES[k] for k in ("const", "species", "mul", "const", "species", "add", "div")]
    assert prog[0][2] == 1.0 and prog[3][2] == 20.0
    with pytest.raises(ffi.RebopError) as err:
        e.lower(["A", "P"], {"V": 1.0})
    assert err.value.status == ffi.ERR_MISSING_PARAM
    assert str(err.value) == "Parameter Km should have a value"
    # a species named like a parameter wins
    prog = ffi.PExpr("V").lower(["V"], {"V": 3.0})
    assert prog[0][0] == ffi.OPCODES["species"]


# ---- code generation (K2) ------------------------------------------------------------------
@pytest.mark.parametrize("name", ["sir", "dimers", "vilar", "mm_lma"])
@pytest.mark.parametrize("arith", [0, 1])
def test_codegen_source(ffi, name, arith):
    model = models.MODELS[name]()
    src = models.build_network(model, arith).codegen()
    assert '#include "ssa_kernel.cuh"' in src
    assert 'extern "C" __global__' in src and "rb_ssa_jit" in src
    # bit-exactness: every f64 operation of the propensities is an explicit round-to-nearest intrinsic,
    # never an infix a*b+c that nvcc could contract into an FMA
    body = src.split("double propensities(")[1].split("__device__ __forceinline__ bool fire")[0]
    assert "__dmul_rn" in body and ("__dadd_rn" in body or len(model["reactions"]) < 2)
    for line in body.splitlines():
        if "c[" in line and "=" in line and "rb_u64" not in line:
            assert " * " not in line and " + " not in line, line


def test_codegen_is_deterministic_and_keyed_by_network(ffi):
    a = models.build_network(models.vilar()).codegen()
    b = models.build_network(models.vilar()).codegen()
    c = models.build_network(models.vilar(), 1).codegen()
    assert a == b and a != c


def test_nvrtc_compiles_sm100a_cubin_without_gpu(ffi):
    cubin = models.build_network(models.vilar(), 1).jit_cubin()
    assert cubin[:4] == b"\x7fELF" and len(cubin) > 10000


def test_large_network_gets_the_shared_memory_form(ffi):
    """100 species / 500 reactions: f64 state columns in shared memory, checkpointed cumulative sums."""
    model = models.synthetic()
    net = models.build_network(model)
    assert net.nb_reactions == 500
    src = net.codegen()
    assert "large form" in src and "32 checkpoints of 16 reactions" in src and "rb_large_select<32, 16, 500, false, BLOCK>" in src
    assert "static constexpr int BLOCK = 128;" in src
    cubin = net.jit_cubin()  # NVRTC, sm_100a, no GPU needed
    assert cubin[:4] == b"\x7fELF"


def test_network_no_specialised_form_can_hold(ffi):
    """Three reactant terms in a network beyond the register-resident limits: only the table-driven kernel runs it."""
    S = 120
    net = ffi.Network(S)
    diff = [0] * S
    diff[0], diff[1], diff[2], diff[3] = -1, -1, -1, 1
    net.add_reaction_lma_sparse(1.0, [(0, 1), (1, 1), (2, 1)], diff)
    with pytest.raises(ffi.RebopError) as e:
        net.codegen()
    assert e.value.status == ffi.ERR_LIMIT


def test_synthetic_network_is_reproducible_and_conserves_mass():
    a, b = models.synthetic(), models.synthetic()
    assert a["reactions"] == b["reactions"] and a["x0"] == b["x0"]
    mass = np.array([1 + s % 3 for s in range(100)])
    for k, terms, diff in a["reactions"]:
        assert np.dot(mass, diff) == 0 and k > 0
        assert 1 <= sum(e for _, e in terms) <= 2
    assert any(e == 2 for _, terms, _ in a["reactions"] for _, e in terms)


def _build_c_demo(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "c_abi_demo")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "examples", "c_abi_demo.c"), "-L", os.path.join(root, "rebop_b200"),
                           "-lrebop_b200", "-Wl,-rpath," + os.path.join(root, "rebop_b200"), "-o", exe])
    return exe


def test_c99_program_links_against_the_abi(tmp_path, ffi):
    """The header is plain C99 and the library links from C: host-side entries run, compute fails loudly without a GPU."""
    import subprocess
    exe = _build_c_demo(tmp_path)
    if ffi.device_count() > 0:
        pytest.skip("covered by the gpu variant")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "specialised kernel source" in res.stdout and "no device: rebop_batch_create -> 5" in res.stdout


@pytest.mark.gpu
def test_c99_program_reproduces_the_golden_vector(tmp_path, ffi):
    import subprocess
    exe = _build_c_demo(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "S,I,R at t=250 = 0,227,773 after 1772 events" in res.stdout


def test_rust_sys_crate_is_current_and_complete(ffi):
    """bindings/rust/rebop-b200-sys/src/lib.rs is generated from the header (no Rust toolchain here): it must be
    up to date and declare every exported function; the safe wrapper may only call declared functions."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.call([sys.executable, os.path.join(root, "scripts", "gen_rust_sys.py"), "--check"]) == 0
    sys_rs = open(os.path.join(root, "bindings", "rust", "rebop-b200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (rebop_\w+)\(", sys_rs))
    assert declared == set(ffi.SIGNATURES)
    wrapper = open(os.path.join(root, "bindings", "rust", "rebop-b200", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(rebop_\w+)\(", wrapper))
    assert used and used <= declared
    consts = set(re.findall(r"sys::(REBOP_\w+)", wrapper))
    assert consts <= set(re.findall(r"pub const (REBOP_\w+):", sys_rs))


@pytest.mark.parametrize("name,kwargs,arith,max_regs", [
    ("vilar", {}, 1, 96),          # 5 CTAs of 128 threads per SM
    ("vilar", {}, 0, 96),
    ("dimers", {}, 1, 96),
    ("ring", {"n": 40}, 0, 255),   # register-resident at 2 CTAs per SM
    ("synthetic", {}, 0, 255),     # shared-memory form
])
def test_generated_kernels_fit_their_register_budget_without_spills(ffi, tmp_path, name, kwargs, arith, max_regs):
    """Resource usage of the NVRTC cubins (no GPU needed): the occupancy the launch logic assumes, and no
    local-memory spills in any of the specialised kernels."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    net = models.build_network(models.MODELS[name](**kwargs), arith)
    cubin = tmp_path / "k.cubin"
    cubin.write_bytes(net.jit_cubin())
    usage = subprocess.check_output(["cuobjdump", "-res-usage", str(cubin)], text=True)
    found = re.findall(r"Function (rb_ssa_jit\w*):\s*\n\s*REG:(\d+) STACK:(\d+)", usage)
    assert {f[0] for f in found} == {"rb_ssa_jit", "rb_ssa_jit_dyn", "rb_ssa_jit_dns"}
    for fn, regs, stack in found:
        assert int(regs) <= max_regs, (fn, regs)
        assert int(stack) == 0, (fn, "spills", stack)


@pytest.mark.parametrize("name,arith,kernel,budget", [
    ("vilar", 1, "rb_ssa_jit_dyn", 170),    # 164 at the end of round 2 (185 at its start): the headline kernel is issue-bound
    ("dimers", 1, "rb_ssa_jit_dyn", 112),   # 107
    ("sir", 0, "rb_ssa_jit_dns", 110),      # 103
])
def test_pass_stays_within_its_instruction_budget(ffi, tmp_path, name, arith, kernel, budget):
    """The ensemble kernels are bound by issue slots, so the number of SASS instructions on the
    hot path of the pass is the figure to watch: scripts/sass_path.py counts it from the NVRTC cubin, no GPU needed."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    net = models.build_network(models.MODELS[name](), arith)
    cubin = tmp_path / "k.cubin"
    cubin.write_bytes(net.jit_cubin())
    out = subprocess.check_output([sys.executable, os.path.join(root, "scripts", "sass_path.py"), str(cubin), kernel], text=True)
    m = re.search(r"hot path (\d+) instructions", out)
    assert m, out
    assert 60 <= int(m.group(1)) <= budget, out
