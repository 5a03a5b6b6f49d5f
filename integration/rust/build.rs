// NOT COMPILED IN THIS REPO: rustc/cargo are absent from the build image.
// Reviewed source of the Rust side of the drop-in; the same call sequence is exercised by host/rtw.hpp (C++) and api.py.
// build.rs — compiles the CUDA side for sm_100a and links it.
use std::{env, path::PathBuf, process::Command};
fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from("rtw/csrc");                      // rust-ray-tracing-in-a-weekend_b200/csrc vendored as rtw/
    let lib = out.join("librtw.so");
    let st = Command::new("nvcc")
        .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("rtw_api.cu")).arg(csrc.join("scene_host.cpp")).arg(csrc.join("image_out.cpp"))
        .arg("-lcudart")
        .status().expect("nvcc not found");
    assert!(st.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rtw");
    println!("cargo:rerun-if-changed=rtw/csrc");
}
