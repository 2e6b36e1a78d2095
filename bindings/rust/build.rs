// Links libmgym.so (which carries libcudart statically and resolves NCCL with dlsym at run time).
fn main() {
    let dir = std::env::var("MGYM_LIB_DIR").unwrap_or_else(|_| "../../modurl_gym_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=mgym");
    println!("cargo:rerun-if-env-changed=MGYM_LIB_DIR");
}
