// Links libmagnetite_b200.so (built by `make -C magnetite_b200/csrc`).
fn main() {
    let dir = std::env::var("MAGNETITE_B200_LIB_DIR").unwrap_or_else(|_| "../../magnetite_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=magnetite_b200");
    println!("cargo:rerun-if-env-changed=MAGNETITE_B200_LIB_DIR");
}
