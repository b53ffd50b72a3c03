// Points the linker at <repo>/takzero_b200/libtakzero_b200.so (built by `python -m takzero_b200.build`).
// Override with TAKZERO_B200_LIB_DIR.
fn main() {
    let dir = std::env::var("TAKZERO_B200_LIB_DIR")
        .unwrap_or_else(|_| format!("{}/../../takzero_b200", env!("CARGO_MANIFEST_DIR")));
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=takzero_b200");
    println!("cargo:rerun-if-env-changed=TAKZERO_B200_LIB_DIR");
}
