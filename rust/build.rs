fn main() {
    // libwhisper_b200.so is built by `make -C whisper-burn_b200/csrc` (nvcc, sm_100a)
    let dir = std::env::var("WB200_LIB_DIR").unwrap_or_else(|_| "../whisper-burn_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=whisper_b200");
}
