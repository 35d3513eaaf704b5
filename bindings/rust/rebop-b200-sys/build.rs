// build.rs of rebop-b200-sys: the build-time half of `define_system!` for the GPU engine.
//
// rebop's macro expands a system into straight-line Rust at compile time (src/gillespie_macro.rs:49-129).  The GPU
// equivalent is a network-specialised CUDA kernel per system, so this script
//   1. builds the engine's objects and its generator tool (make -C <csrc> objects),
//   2. runs the generator on every `systems/*.rsys` of this crate (the text of a define_system! invocation),
//   3. compiles each generated translation unit with nvcc for sm_100a, without FMA contraction,
//   4. links engine objects + system objects into $OUT_DIR/librebop_b200.so and tells cargo to link it.
// A `SystemBatch` created at run time from the same text then finds its kernel already compiled (the engine keys
// build-time kernels by the generated source, so parameter values do not matter) and never calls NVRTC.
//
// The command lines are kept as plain templates ({csrc}, {out}, {rsys}, {stem}, {objects}) so that the recipe can be
// exercised without a Rust toolchain: tests/test_rust_recipe.py reads them from this file and runs them.
// NOTE: no cargo/rustc in the image this repository is built in -- the Rust code is unverified, the recipe is tested.
use std::path::{Path, PathBuf};
use std::process::Command;

const MAKE_CMD: &str = "make -C {csrc} -j8 objects";
const SYSGEN_CMD: &str = "{csrc}/build/rebop_sysgen {rsys} -o {out}/sys_{stem}.cu";
const NVCC_CMD: &str = "nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -I{csrc} -c {out}/sys_{stem}.cu -o {out}/sys_{stem}.o";
const LINK_CMD: &str = "nvcc -gencode arch=compute_100a,code=sm_100a -cudart static -shared -o {out}/librebop_b200.so {objects} -ldl -lpthread";

fn run(template: &str, vars: &[(&str, &str)]) {
    let mut line = template.to_string();
    for (key, value) in vars {
        line = line.replace(&format!("{{{key}}}"), value);
    }
    let mut words = line.split_whitespace();
    let program = words.next().expect("empty command");
    let status = Command::new(program).args(words).status().unwrap_or_else(|e| panic!("cannot run `{line}`: {e}"));
    assert!(status.success(), "`{line}` failed");
}

fn main() {
    let manifest = PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
    let out = PathBuf::from(std::env::var("OUT_DIR").unwrap());
    // the engine's sources: REBOP_B200_CSRC, or the checkout this crate lives in
    let csrc = std::env::var("REBOP_B200_CSRC").map(PathBuf::from).unwrap_or_else(|_| manifest.join("../../../rebop_b200/csrc"));
    let csrc = csrc.canonicalize().expect("REBOP_B200_CSRC does not point at rebop_b200/csrc");
    let (csrc_s, out_s) = (csrc.to_str().unwrap(), out.to_str().unwrap());

    run(MAKE_CMD, &[("csrc", csrc_s)]);
    let mut objects: Vec<String> = std::fs::read_to_string(csrc.join("build/objects.txt"))
        .expect("make objects did not leave build/objects.txt")
        .split_whitespace()
        .map(|o| csrc.join(o).to_str().unwrap().to_string())
        .collect();

    let systems = std::env::var("REBOP_B200_SYSTEMS").map(PathBuf::from).unwrap_or_else(|_| manifest.join("systems"));
    println!("cargo:rerun-if-changed={}", systems.display());
    let mut entries: Vec<PathBuf> = std::fs::read_dir(&systems)
        .map(|d| d.filter_map(|e| e.ok().map(|e| e.path())).filter(|p| p.extension().map_or(false, |x| x == "rsys")).collect())
        .unwrap_or_default();
    entries.sort();
    for rsys in &entries {
        let stem = Path::new(rsys).file_stem().unwrap().to_str().unwrap();
        let vars = [("csrc", csrc_s), ("out", out_s), ("rsys", rsys.to_str().unwrap()), ("stem", stem)];
        run(SYSGEN_CMD, &vars);
        run(NVCC_CMD, &vars);
        objects.push(format!("{out_s}/sys_{stem}.o"));
        println!("cargo:rerun-if-changed={}", rsys.display());
    }
    run(LINK_CMD, &[("out", out_s), ("objects", &objects.join(" "))]);

    println!("cargo:rustc-link-search=native={out_s}");
    println!("cargo:rustc-link-lib=dylib=rebop_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{out_s}");
    println!("cargo:rerun-if-env-changed=REBOP_B200_CSRC");
    println!("cargo:rerun-if-env-changed=REBOP_B200_SYSTEMS");
}
