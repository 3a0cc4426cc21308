// Builds libqvnt_b200.so with nvcc for sm_100a (the Makefile passes
// `-gencode arch=compute_100a,code=sm_100a -lineinfo`) and links it dynamically.
// QVNT_B200_ROOT points at the qvnt-b200 checkout (default: two levels up from this crate).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = env::var("QVNT_B200_ROOT")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../.."));
    let csrc = root.join("qvnt_b200").join("csrc");
    let status = Command::new("make")
        .arg("-C")
        .arg(&csrc)
        .arg("-j")
        .status()
        .expect("failed to run make (nvcc required: there is no CPU fallback)");
    assert!(status.success(), "building libqvnt_b200.so failed");
    let libdir = root.join("qvnt_b200");
    println!("cargo:rustc-link-search=native={}", libdir.display());
    println!("cargo:rustc-link-lib=dylib=qvnt_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", libdir.display());
    println!("cargo:rerun-if-changed={}", root.join("include/qvnt_b200.h").display());
    println!("cargo:rerun-if-changed={}", csrc.display());
}
