// Authored, NOT compiled here (no cargo in the build image).
// FHESTR_ENGINE_DIR = directory holding libfhestr_engine.so (fhestring_b200/ in this repository).
fn main() {
    let dir = std::env::var("FHESTR_ENGINE_DIR").expect("set FHESTR_ENGINE_DIR to the directory of libfhestr_engine.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=fhestr_engine");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=FHESTR_ENGINE_DIR");
}
