// Tells cargo where libfinch_b200.so lives and to link it.  `links = "finch_b200"` in Cargo.toml requires this
// build script; FINCH_B200_LIB_DIR defaults to the in-tree build output (finch_rs_b200/ next to rust/).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("FINCH_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../finch_rs_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=finch_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=FINCH_B200_LIB_DIR");
}
