// Builds libgrail_cuda from the in-tree CUDA sources with nvcc (through the cc crate) for sm_100a and
// generates the bindings with bindgen.  NOTE: written for the reference's toolchain; this repository's
// build image has no rustc/cargo, so this file is exercised only on a maintainer's machine.
use std::{env, path::PathBuf};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../..");
    let csrc = root.join("grail-rs_b200/csrc");
    println!("cargo:rerun-if-changed={}", csrc.display());
    cc::Build::new()
        .cuda(true)
        .cudart("shared")
        .flag("-std=c++17")
        .flag("-O3")
        .flag("-lineinfo")
        .flag("--expt-relaxed-constexpr")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        // the host planner shares the strict-f32 clock code with the kernels: no contraction on the host side
        .flag("-Xcompiler")
        .flag("-ffp-contract=off")
        .file(csrc.join("grail_runtime.cu"))
        .file(csrc.join("grail_text.cpp"))          // grail_cuda_transcribe_batch
        .compile("grail_cuda");
    let bindings = bindgen::Builder::default()
        .header(root.join("include/grail_cuda.h").to_str().unwrap())
        .allowlist_function("grail_cuda_.*")
        .allowlist_type("grail_.*")
        .prepend_enum_name(false)                   // GRAIL_F32, not grail_sample_format_GRAIL_F32
        .generate()
        .expect("bindgen failed");
    bindings
        .write_to_file(PathBuf::from(env::var("OUT_DIR").unwrap()).join("bindings.rs"))
        .unwrap();
}
