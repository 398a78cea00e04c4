// Links libb200mm.so (built in-tree by wgpu_mm_b200/csrc/Makefile).  B200MM_LIB_DIR overrides the location.
fn main() {
    let dir = std::env::var("B200MM_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../wgpu_mm_b200/lib", here)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=b200mm");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=B200MM_LIB_DIR");
}
