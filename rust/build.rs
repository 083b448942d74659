// build.rs addition for tphakala/birda (the reference's build.rs:1-26 only embeds version strings).
// Builds the sm_100a library with nvcc and links it.  Untested here: no Rust toolchain in the image.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("BIRDA_B200_DIR").unwrap_or_else(|_| "../birda_b200".into()));
    let status = Command::new("make")
        .arg("-C").arg(root.join("csrc"))
        .arg(format!("-j{}", env::var("NUM_JOBS").unwrap_or_else(|_| "8".into())))
        .status()
        .expect("failed to run make for birda_b200 (is nvcc on PATH?)");
    assert!(status.success(), "birda_b200 build failed");
    println!("cargo:rustc-link-search=native={}", root.display());
    println!("cargo:rustc-link-lib=dylib=birda_b200");
    println!("cargo:rerun-if-changed={}", root.join("csrc").display());
    println!("cargo:rerun-if-changed=../include/birda_b200.h");
}
