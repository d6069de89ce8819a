// build.rs -- link the prebuilt CUDA library (include/fse_b200.h).  FSE_B200_LIB_DIR points at the
// directory holding libfse_b200.so (in this repository: entropy_coders_b200/).
fn main() {
    let dir = std::env::var("FSE_B200_LIB_DIR").unwrap_or_else(|_| "../entropy_coders_b200".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=fse_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rerun-if-env-changed=FSE_B200_LIB_DIR");
}
