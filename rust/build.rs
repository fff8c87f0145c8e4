// Link against the prebuilt C-ABI library (built by `python gym_rs_b200/build.py`).
fn main() {
    let dir = std::env::var("GYMRS_B200_LIB_DIR").unwrap_or_else(|_| "../gym_rs_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=gymrs_b200");
    println!("cargo:rerun-if-env-changed=GYMRS_B200_LIB_DIR");
}
