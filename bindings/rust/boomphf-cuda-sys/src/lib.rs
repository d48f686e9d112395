//! Declarations of the C ABI in `include/boomphf_cuda.h` (hand-written equivalent of the bindgen output).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[cfg(feature = "bindgen")]
include!(concat!(env!("OUT_DIR"), "/bindings.rs"));

#[cfg(not(feature = "bindgen"))]
mod hand {
    use super::*;

    #[repr(C)]
    pub struct bphf_mphf {
        _private: [u8; 0],
    }

    pub const BPHF_OK: c_int = 0;
    pub const BPHF_ERR_ARG: c_int = -1;
    pub const BPHF_ERR_GAMMA: c_int = -2; // assert!(gamma > 1.01)
    pub const BPHF_ERR_MAX_ITERS: c_int = -3; // panic!("counldn't find unique hashes")
    pub const BPHF_ERR_CUDA: c_int = -4;
    pub const BPHF_ERR_NOMEM: c_int = -5;
    pub const BPHF_ERR_INTERNAL: c_int = -6;
    pub const BPHF_NONE: u64 = u64::MAX; // try_hash == None
    pub const BPHF_MEM_HOST: c_int = 0;
    pub const BPHF_MEM_DEVICE: c_int = 1;
    pub const BPHF_KEY_U64: c_int = 0;
    pub const BPHF_KEY_UINT: c_int = 1;
    pub const BPHF_KEY_BYTES: c_int = 2;
    pub const BPHF_KEY_STREAM: c_int = 3;

    extern "C" {
        pub fn bphf_abi_version() -> c_int;
        pub fn bphf_last_error() -> *const c_char;
        pub fn bphf_device_count(count: *mut c_int) -> c_int;

        pub fn bphf_build_u64(keys: *const u64, n: u64, gamma: f64, has_seed: c_int, seed: u64, device: c_int,
                              keys_mem: c_int, stream: *mut c_void, out: *mut *mut bphf_mphf) -> c_int;
        pub fn bphf_build_keys(keys: *const c_void, n: u64, key_kind: c_int, key_width: c_int, gamma: f64,
                               has_seed: c_int, seed: u64, device: c_int, keys_mem: c_int,
                               out: *mut *mut bphf_mphf) -> c_int;
        pub fn bphf_build_stream(bytes: *const u8, n_bytes: u64, seg_ends: *const u64, n_segs: u64,
                                 key_segs: *const u64, n_keys: u64, gamma: f64, has_seed: c_int, seed: u64,
                                 device: c_int, mem: c_int, out: *mut *mut bphf_mphf) -> c_int;
        pub fn bphf_build_chunked_u64(chunks: *const *const u64, chunk_lens: *const u64, n_chunks: u64, n: u64,
                                      gamma: f64, parallel_variant: c_int, has_max_iters: c_int, max_iters: u64,
                                      device: c_int, keys_mem: c_int, out: *mut *mut bphf_mphf) -> c_int;

        pub fn bphf_num_levels(m: *const bphf_mphf, n_levels: *mut u32) -> c_int;
        pub fn bphf_level_dims(m: *const bphf_mphf, level: u32, bits: *mut u64, n_words: *mut u64,
                               n_ranks: *mut u64) -> c_int;
        pub fn bphf_level_export(m: *const bphf_mphf, level: u32, words_out: *mut u64, ranks_out: *mut u64,
                                 out_mem: c_int) -> c_int;
        pub fn bphf_from_levels_keys(n_levels: u32, bits: *const u64, words: *const *const u64,
                                     ranks: *const *const u64, key_kind: c_int, key_width: c_int, device: c_int,
                                     out: *mut *mut bphf_mphf) -> c_int;
        pub fn bphf_free(m: *mut bphf_mphf);

        pub fn bphf_hash_batch_u64(m: *const bphf_mphf, keys: *const u64, n: u64, out: *mut u64, mem: c_int,
                                   stream: *mut c_void) -> c_int;
        pub fn bphf_hash_batch_keys(m: *const bphf_mphf, keys: *const c_void, n: u64, out: *mut u64, mem: c_int,
                                    stream: *mut c_void) -> c_int;
        pub fn bphf_hash_batch_stream(m: *const bphf_mphf, bytes: *const u8, n_bytes: u64, seg_ends: *const u64,
                                      n_segs: u64, key_segs: *const u64, n_keys: u64, out: *mut u64, mem: c_int,
                                      stream: *mut c_void) -> c_int;

        pub fn bphf_map_permute_u64(m: *const bphf_mphf, keys: *const u64, values: *const u64, n: u64,
                                    keys_out: *mut u64, values_out: *mut u64, mem: c_int, stream: *mut c_void) -> c_int;
        pub fn bphf_map_get_batch_u64(m: *const bphf_mphf, map_keys: *const u64, map_values: *const u64, n_map: u64,
                                      queries: *const u64, n_queries: u64, values_out: *mut u64, found_out: *mut u8,
                                      mem: c_int, stream: *mut c_void) -> c_int;

        pub fn bphf_set_ingest_mode(mode: c_int, ring_keys: u64) -> c_int;
    }
}
#[cfg(not(feature = "bindgen"))]
pub use hand::*;
