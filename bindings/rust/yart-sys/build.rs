// Points the linker at libyart_b200.so.  YART_LIB_DIR = the directory that holds it (the package directory
// yet-another-raytracer_b200/ of the repository after `python yet-another-raytracer_b200/build.py`).
fn main() {
    let dir = std::env::var("YART_LIB_DIR").unwrap_or_else(|_| "../../../yet-another-raytracer_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=yart_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=YART_LIB_DIR");
}
