// Link libq3tts_b200.so when the `b200` feature is on.  Q3TTS_B200_LIB_DIR points at the directory holding the library
// (default: ../qwen3_tts_rs_b200 relative to the crate, where `python -m qwen3_tts_rs_b200.build` leaves it).
fn main() {
    if std::env::var_os("CARGO_FEATURE_B200").is_some() {
        let dir = std::env::var("Q3TTS_B200_LIB_DIR").unwrap_or_else(|_| "../qwen3_tts_rs_b200".to_string());
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=q3tts_b200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
        println!("cargo:rerun-if-env-changed=Q3TTS_B200_LIB_DIR");
    }
}
