// UNBUILT IN THIS ENVIRONMENT (no Rust toolchain). Links libincerto_b200.so built by
// `python -m incerto_b200._build`, or rebuilds it from csrc/ with nvcc when INCERTO_B200_BUILD=1.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let lib_dir = env::var("INCERTO_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| root.join("incerto_b200/lib"));
    if env::var("INCERTO_B200_BUILD").is_ok() {
        let status = Command::new("python").args(["-m", "incerto_b200._build"]).current_dir(&root).status().expect("run _build.py");
        assert!(status.success(), "nvcc build of the sm_100a engine failed");
    }
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=incerto_b200");
    println!("cargo:rerun-if-changed=../../include/incerto_b200.h");
}
