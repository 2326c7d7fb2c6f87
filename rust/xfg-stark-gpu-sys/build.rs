// Links against libxfgstark.so.  XFGSTARK_LIB_DIR = directory holding the library built by `make -C xfg-stark_b200/csrc`.
fn main() {
    let dir = std::env::var("XFGSTARK_LIB_DIR").unwrap_or_else(|_| "../../xfg-stark_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=xfgstark");
    println!("cargo:rerun-if-env-changed=XFGSTARK_LIB_DIR");
}
