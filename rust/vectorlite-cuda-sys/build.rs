// VECTORLITE_CUDA_LIB_DIR = the directory that holds libvectorlite_cuda.so (vectorlite_b200/ in the vectorlite-b200 repo)
fn main() {
    let dir = std::env::var("VECTORLITE_CUDA_LIB_DIR").expect("set VECTORLITE_CUDA_LIB_DIR to the directory of libvectorlite_cuda.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=vectorlite_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=VECTORLITE_CUDA_LIB_DIR");
}
