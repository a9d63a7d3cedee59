// core/build.rs — replaces the WGSL preprocessing step of the reference (core/build.rs:5-17):
// instead of expanding shaders into OUT_DIR, compile the CUDA library for sm_100a with nvcc and
// tell cargo to link it.  The CUDA sources (kmeans-gpu_b200/csrc and include/kmeans_gpu.h of the
// B200 repository) are expected under core/csrc and core/include; KMG_CSRC / KMG_INCLUDE override.
use std::{
    env,
    path::{Path, PathBuf},
    process::Command,
};

const SOURCES: &[&str] = &["kmg_api.cu", "kmg_host.cpp"];
const HEADERS: &[&str] = &[
    "kmg_kernels.cuh",
    "kmg_lloyd_ring.cuh",
    "kmg_audit.cuh",
    "kmg_init_lazy.cuh",
    "kmg_small.cuh",
    "kmg_math.cuh",
];

fn main() {
    let out_dir = PathBuf::from(env::var_os("OUT_DIR").expect("OUT_DIR"));
    let csrc = env::var_os("KMG_CSRC")
        .map(PathBuf::from)
        .unwrap_or_else(|| Path::new("csrc").to_path_buf());
    let include = env::var_os("KMG_INCLUDE")
        .map(PathBuf::from)
        .unwrap_or_else(|| Path::new("include").to_path_buf());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| {
        let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
        format!("{cuda}/bin/nvcc")
    });

    let lib = out_dir.join("libkmeans_gpu.so");
    let mut cmd = Command::new(&nvcc);
    cmd.args([
        "-gencode",
        "arch=compute_100a,code=sm_100a", // B200 only: no PTX for other architectures, no fallback
        "-O3",
        "-std=c++17",
        "-lineinfo",
        "--compiler-options",
        "-fPIC,-fvisibility=default,-ffp-contract=off",
        "-shared",
        "-I",
    ]);
    cmd.arg(&include).arg("-o").arg(&lib);
    for s in SOURCES {
        cmd.arg(csrc.join(s));
    }
    cmd.arg("-ldl");
    let status = cmd
        .status()
        .unwrap_or_else(|e| panic!("could not run {nvcc}: {e} (set NVCC or CUDA_HOME)"));
    assert!(status.success(), "nvcc failed building libkmeans_gpu.so");

    println!("cargo:rustc-link-search=native={}", out_dir.display());
    println!("cargo:rustc-link-lib=dylib=kmeans_gpu");
    // the test and example binaries find the library next to the build output
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out_dir.display());
    for f in SOURCES.iter().chain(HEADERS.iter()) {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!(
        "cargo:rerun-if-changed={}",
        include.join("kmeans_gpu.h").display()
    );
    println!("cargo:rerun-if-env-changed=NVCC");
    println!("cargo:rerun-if-env-changed=CUDA_HOME");
}
