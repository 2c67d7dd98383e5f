// Links libptb200.so.  PT_B200_LIB_DIR = <repo>/thu-acg-f2024-path-tracer_b200/lib (built by `make` there, sm_100a only).
fn main() {
    let dir = std::env::var("PT_B200_LIB_DIR").expect("set PT_B200_LIB_DIR to the directory that holds libptb200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ptb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=PT_B200_LIB_DIR");
}
