// Links libellcuda.so.  ELLCUDA_LIB_DIR = directory holding the library (…/ellalgo_rs_b200/csrc).
fn main() {
    if let Ok(dir) = std::env::var("ELLCUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=ellcuda");
    println!("cargo:rerun-if-env-changed=ELLCUDA_LIB_DIR");
}
