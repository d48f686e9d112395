// Links libboomphf_cuda.so.  BOOMPHF_CUDA_LIB_DIR = directory holding the library (rust-boomphf_b200/ of this
// repository after `python -c "import __graft_entry__ as g; g.build()"`); BOOMPHF_CUDA_INCLUDE_DIR = include/.
fn main() {
    let lib_dir = std::env::var("BOOMPHF_CUDA_LIB_DIR").expect("set BOOMPHF_CUDA_LIB_DIR to the directory of libboomphf_cuda.so");
    println!("cargo:rustc-link-search=native={}", lib_dir);
    println!("cargo:rustc-link-lib=dylib=boomphf_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir);
    println!("cargo:rerun-if-env-changed=BOOMPHF_CUDA_LIB_DIR");
    #[cfg(feature = "bindgen")]
    {
        let inc = std::env::var("BOOMPHF_CUDA_INCLUDE_DIR").expect("set BOOMPHF_CUDA_INCLUDE_DIR to include/");
        let out = std::path::PathBuf::from(std::env::var("OUT_DIR").unwrap()).join("bindings.rs");
        bindgen::Builder::default()
            .header(format!("{}/boomphf_cuda.h", inc))
            .allowlist_function("bphf_.*")
            .allowlist_var("BPHF_.*")
            .allowlist_type("bphf_.*")
            .generate()
            .expect("bindgen failed")
            .write_to_file(out)
            .unwrap();
    }
}
