// Links libzkb.so (built by `python eigen-zeth_b200/build.py`: nvcc -gencode arch=compute_100a,code=sm_100a).
// eigen-zeth's own build.rs (tonic_build of proto/prover/v1/prover.proto) is untouched.
fn main() {
    let dir = std::env::var("ZKB_LIB_DIR").expect("set ZKB_LIB_DIR to the directory holding libzkb.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zkb");
    println!("cargo:rerun-if-changed=../../include/zkb.h");
    println!("cargo:rerun-if-env-changed=ZKB_LIB_DIR");
}
