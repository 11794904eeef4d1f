// Links the prebuilt CUDA library (built by `python -m rusty_compression_b200.build`, i.e. nvcc for sm_100a).
fn main() {
    let dir = std::env::var("RC_B200_LIB_DIR").expect("set RC_B200_LIB_DIR to the directory holding librc_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rc_b200");
    println!("cargo:rerun-if-env-changed=RC_B200_LIB_DIR");
}
