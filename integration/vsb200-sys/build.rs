// build.rs — builds libvsb200.so with nvcc for sm_100a and tells cargo how to link it.
//
// VSB200_ROOT      checkout of the vsb200 repository (default: ../../third_party/vsb200)
// VSB200_LIB_DIR   skip the build and link a prebuilt libvsb200.so from this directory
// NVCC             compiler (default /usr/local/cuda/bin/nvcc); the Makefile passes
//                  -gencode arch=compute_100a,code=sm_100a — there is no other target and no CPU fallback.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    println!("cargo:rerun-if-env-changed=VSB200_ROOT");
    println!("cargo:rerun-if-env-changed=VSB200_LIB_DIR");
    println!("cargo:rerun-if-env-changed=NVCC");

    if let Ok(dir) = env::var("VSB200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=vsb200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
        return;
    }

    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").expect("CARGO_MANIFEST_DIR"));
    let root = env::var("VSB200_ROOT")
        .map(PathBuf::from)
        .unwrap_or_else(|_| manifest.join("../../third_party/vsb200"));
    let csrc = root.join("vector-store_b200").join("csrc");
    let header = root.join("include").join("vsb200.h");
    assert!(
        header.exists(),
        "vsb200 sources not found under {} (set VSB200_ROOT or VSB200_LIB_DIR)",
        root.display()
    );
    println!("cargo:rerun-if-changed={}", header.display());
    println!("cargo:rerun-if-changed={}", csrc.display());

    let jobs = env::var("NUM_JOBS").unwrap_or_else(|_| "8".to_string());
    let mut make = Command::new("make");
    make.arg("-C").arg(&csrc).arg("-j").arg(&jobs).arg("-s");
    if let Ok(nvcc) = env::var("NVCC") {
        make.arg(format!("NVCC={nvcc}"));
    }
    let status = make.status().expect("failed to run make for libvsb200.so (is nvcc installed?)");
    assert!(status.success(), "building libvsb200.so failed");

    let lib_dir = root.join("vector-store_b200");
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=vsb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
    println!("cargo:include={}", root.join("include").display());
}
