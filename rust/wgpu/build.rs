fn main() {
    // libwgebra_b200.so is built by `python -m wgmath_b200.build` (nvcc, sm_100a).
    if let Ok(dir) = std::env::var("WGEBRA_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=wgebra_b200");
    println!("cargo:rerun-if-env-changed=WGEBRA_B200_LIB_DIR");
}
