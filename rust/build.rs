// build.rs for annembed with `--features cuda` (SOURCE ONLY; replaces the 7-line build.rs of the reference,
// /root/reference/build.rs:1-7, keeping its macOS branch).  Compiles the CUDA library for sm_100a with nvcc and
// links it; there is no CPU fallback and no multi-backend dispatch.
use std::{env, path::PathBuf, process::Command};

fn main() {
    if cfg!(target_os = "macos") {
        println!("cargo:rustc-link-lib=framework=Accelerate");
    }
    if env::var("CARGO_FEATURE_CUDA").is_err() {
        return;
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src = PathBuf::from(env::var("ANNEMBED_CUDA_SRC").unwrap_or_else(|_| "annembed_b200/csrc".into()));
    let lib = out.join("libannembed_cuda.a");
    let obj = out.join("annembed_cuda.o");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let st = Command::new(&nvcc)
        .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-c"])
        .arg(src.join("annembed_cuda.cu"))
        .arg("-o")
        .arg(&obj)
        .status()
        .expect("nvcc not found");
    assert!(st.success(), "nvcc failed");
    let st = Command::new("ar").arg("rcs").arg(&lib).arg(&obj).status().expect("ar not found");
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=annembed_cuda");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl"); // NCCL is dlopen'ed at comm_init, never linked
    println!("cargo:rerun-if-changed={}", src.display());
}
