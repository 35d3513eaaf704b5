"""The build recipe of the Rust host crate (bindings/rust/rebop-b200-sys/build.rs), exercised without a Rust toolchain.

north_star: "define_system! emits a network-specialised kernel at build time via nvcc in build.rs"
(src/gillespie_macro.rs:49-129 is what it replaces).  cargo/rustc are not in this image, so the Rust code itself is
unverified; what CAN be verified is the recipe: the command templates are read out of build.rs and run as written --
generator tool, nvcc for sm_100a without FMA contraction, link -- on the crate's own systems/*.rsys, and the library
that comes out must register a build-time kernel for each of them.
"""
import ctypes
import os
import re
import shlex
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CRATE = os.path.join(ROOT, "bindings", "rust", "rebop-b200-sys")
CSRC = os.path.join(ROOT, "rebop_b200", "csrc")


def templates():
    text = open(os.path.join(CRATE, "build.rs")).read()
    found = dict(re.findall(r'const (\w+_CMD): &str = "([^"]+)";', text))
    assert set(found) == {"MAKE_CMD", "SYSGEN_CMD", "NVCC_CMD", "LINK_CMD"}, found
    return found


def test_recipe_uses_the_flags_the_engine_needs():
    t = templates()
    assert "arch=compute_100a,code=sm_100a" in t["NVCC_CMD"] and "-fmad=false" in t["NVCC_CMD"] and "-lineinfo" in t["NVCC_CMD"]
    assert "rebop_sysgen" in t["SYSGEN_CMD"] and "-shared" in t["LINK_CMD"]


def test_recipe_builds_a_library_with_the_crates_systems(tmp_path):
    t = templates()
    out = str(tmp_path)

    def run(template, **kw):
        line = template
        for k, v in kw.items():
            line = line.replace("{" + k + "}", v)
        assert "{" not in line, line
        subprocess.check_call(shlex.split(line), stdout=subprocess.DEVNULL)

    run(t["MAKE_CMD"], csrc=CSRC)
    objects = [os.path.join(CSRC, o) for o in open(os.path.join(CSRC, "build", "objects.txt")).read().split()]
    names = []
    for rsys in sorted(os.listdir(os.path.join(CRATE, "systems"))):
        if not rsys.endswith(".rsys"):
            continue
        stem = rsys[:-5]
        path = os.path.join(CRATE, "systems", rsys)
        run(t["SYSGEN_CMD"], csrc=CSRC, out=out, rsys=path, stem=stem)
        run(t["NVCC_CMD"], csrc=CSRC, out=out, stem=stem)
        objects.append(os.path.join(out, f"sys_{stem}.o"))
        names.append(re.search(r"^\s*(\w+)\s*\{", re.sub(r"//.*", "", open(path).read()).split(";", 1)[1], re.M).group(1))
    run(t["LINK_CMD"], out=out, objects=" ".join(objects))
    lib = ctypes.CDLL(os.path.join(out, "librebop_b200.so"))
    count = lib.rebop_b200_prebuilt_count()
    got = []
    for i in range(count):
        buf = ctypes.create_string_buffer(256)
        assert lib.rebop_b200_prebuilt_name(i, buf, 256, None) == 0
        got.append(buf.value.decode())
    assert sorted(got) == sorted(names) and "LotkaVolterra" in got, (got, names)


def test_stringified_macro_input_is_the_same_system(ffi):
    """`define_system!` hands the engine stringify!($($tt)*): one line, tokens separated by single spaces."""
    text = open(os.path.join(CRATE, "systems", "lotka_volterra.rsys")).read()
    one_line = " ".join(re.sub(r"//.*", "", text).replace(",", " , ").replace("=>", " => ").replace("@", " @ ").split())
    a, b = ffi.System(text), ffi.System(one_line)
    assert a.species == b.species == ["prey", "predator"] and a.params == b.params and a.reactions == b.reactions
    net_a, net_b = a.network([1.0, 0.01, 0.5]), b.network([1.0, 0.01, 0.5])
    assert net_a.codegen() == net_b.codegen()
