// Link libtaper_b200.so (built by `make` at the repository root) from TAPER_B200_LIB_DIR, next to the reference's
// macOS Accelerate hook (build.rs:3-7 of the reference).
fn main() {
    let dir = std::env::var("TAPER_B200_LIB_DIR").unwrap_or_else(|_| "../../taper_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=taper_b200");
    println!("cargo:rerun-if-env-changed=TAPER_B200_LIB_DIR");
}
