// qrusty/build.rs -- links the prebuilt CUDA library.  Build it first with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
//        -shared -o libqrusty_cuda.so qrusty_b200/csrc/qrusty_cuda.cu -ldl
// and point QRUSTY_CUDA_LIB_DIR at the directory that holds it.
fn main() {
    let dir = std::env::var("QRUSTY_CUDA_LIB_DIR").unwrap_or_else(|_| "/usr/local/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=qrusty_cuda");
    println!("cargo:rerun-if-env-changed=QRUSTY_CUDA_LIB_DIR");
}
