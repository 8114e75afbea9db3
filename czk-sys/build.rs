// Link against libczk_b200.so built by `python collaborative-zksnark_b200/build.py` (nvcc, sm_100a).
// CZK_B200_LIB_DIR points at the directory holding the library (default: ../collaborative-zksnark_b200).
fn main() {
    let dir = std::env::var("CZK_B200_LIB_DIR").unwrap_or_else(|_| "../collaborative-zksnark_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=czk_b200");
    println!("cargo:rerun-if-env-changed=CZK_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/czk.h");
    println!("cargo:rerun-if-changed=../include/czk_groth16.h");
    println!("cargo:rerun-if-changed=../include/czk_plonk.h");
}
