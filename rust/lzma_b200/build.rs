// Links against liblzma_b200.so built by `make -C lzma_rs_b200/csrc` (path via LZMA_B200_LIB_DIR).
fn main() {
    let dir = std::env::var("LZMA_B200_LIB_DIR").unwrap_or_else(|_| "../../lzma_rs_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=lzma_b200");
    println!("cargo:rerun-if-env-changed=LZMA_B200_LIB_DIR");
}
