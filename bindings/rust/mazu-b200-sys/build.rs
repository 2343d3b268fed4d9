// Links libmazu_b200.so.  MAZU_B200_LIB_DIR = directory holding the library (the repo's mazu_b200/ directory).
fn main() {
    if let Ok(dir) = std::env::var("MAZU_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=mazu_b200");
    println!("cargo:rerun-if-env-changed=MAZU_B200_LIB_DIR");
}
