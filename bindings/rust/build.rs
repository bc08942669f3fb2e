// lib/libmemex/build.rs -- the library is prebuilt with nvcc for sm_100a (python -m memex_b200.build);
// MEMEX_B200_LIB_DIR points at memex_b200/_lib/.
fn main() {
    let dir = std::env::var("MEMEX_B200_LIB_DIR").expect("MEMEX_B200_LIB_DIR = directory of libmemex_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=memex_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=MEMEX_B200_LIB_DIR");
}
