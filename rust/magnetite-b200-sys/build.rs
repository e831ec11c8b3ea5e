// Links libmagnetite_b200.so (built by `make -C magnetite_b200/csrc`).  MAGNETITE_B200_LIB_DIR overrides the
// directory; the default is this repository's magnetite_b200/ next to rust/.  At run time the loader must find it
// too: LD_LIBRARY_PATH=<that directory> (a library crate's build script cannot hand an rpath to the final binary).
fn main() {
    let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap_or_else(|_| ".".to_string());
    let dir = std::env::var("MAGNETITE_B200_LIB_DIR").unwrap_or_else(|_| format!("{manifest}/../../magnetite_b200"));
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=magnetite_b200");
    println!("cargo:rerun-if-env-changed=MAGNETITE_B200_LIB_DIR");
}
