// Link against libsimb200.so built by `make -C sim_b200/csrc` (SIMB200_LIB_DIR points at sim_b200/).
fn main() {
    if let Ok(dir) = std::env::var("SIMB200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=simb200");
}
