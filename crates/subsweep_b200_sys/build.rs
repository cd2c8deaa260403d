// Links libsubsweep_b200.so.  SUBSWEEP_B200_LIB_DIR = the directory that holds it (subsweep_b200/lib after
// `python -m subsweep_b200.build`, or wherever the .so was installed).
fn main() {
    println!("cargo:rerun-if-env-changed=SUBSWEEP_B200_LIB_DIR");
    let dir = std::env::var("SUBSWEEP_B200_LIB_DIR")
        .expect("set SUBSWEEP_B200_LIB_DIR to the directory holding libsubsweep_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=subsweep_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
