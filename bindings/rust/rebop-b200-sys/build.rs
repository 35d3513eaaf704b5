// Links the prebuilt engine.  Build it first with `make -C rebop_b200/csrc` (nvcc, sm_100a) and point
// REBOP_B200_LIB_DIR at the directory that holds librebop_b200.so.
// NOTE: written without a Rust toolchain at hand (none in the build image of this repository): untested.
fn main() {
    let dir = std::env::var("REBOP_B200_LIB_DIR")
        .expect("set REBOP_B200_LIB_DIR to the directory of librebop_b200.so (rebop_b200/ after `make -C rebop_b200/csrc`)");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=rebop_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=REBOP_B200_LIB_DIR");
}
