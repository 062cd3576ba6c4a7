// build.rs — compiles the CUDA sources for sm_100a with nvcc and links them.
// Mirrors the-tessellator_b200/csrc/Makefile; -fmad=false because the reference never contracts a*b+c.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("the-tessellator_b200").join("csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let lib = out.join("libtess_b200.so");
    let status = Command::new(&nvcc)
        .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-prec-div=true", "-prec-sqrt=true"])
        .args(&["-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .args(["capi.cu", "grid.cu", "clip.cu", "clip_thread.cu", "outputs.cu", "query.cu"].iter().map(|f| csrc.join(f)))
        .arg("-lcudart")
        .status()
        .expect("nvcc not found: the-tessellator-b200 has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=tess_b200");
    for f in &["capi.cu", "grid.cu", "clip.cu", "clip_thread.cu", "outputs.cu", "query.cu", "common.cuh", "cube_tables.cuh", "tess_math.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include").join("tess.h").display());
}
