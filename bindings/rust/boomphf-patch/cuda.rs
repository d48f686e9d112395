//! `#[cfg(feature = "cuda")] mod cuda;` inside the `boomphf` crate (next to `src/lib.rs`).
//!
//! `Mphf::new` / `Mphf::new_parallel` start with
//!     #[cfg(feature = "cuda")]
//!     if let Some(m) = cuda::try_new_cuda(gamma, objects, None /* or starting_seed */) { return m; }
//! and keep their bodies as the path for builds without the feature.  The struct, its serde derive and
//! `hash` / `try_hash` are untouched: a GPU-built Mphf is an ordinary Mphf.
use std::ffi::CStr;
use std::hash::{Hash, Hasher};
use std::marker::PhantomData;

use boomphf_cuda_sys as sys;

use crate::bitvector::BitVector;
use crate::Mphf;

/// Everything a wyhash `Hasher` would have seen from `T::hash`, for every key
#[derive(Default)]
struct Recorder {
    bytes: Vec<u8>,
    seg_ends: Vec<u64>,
    key_segs: Vec<u64>,
}
struct OneKey<'a>(&'a mut Recorder);
impl<'a> Hasher for OneKey<'a> {
    fn write(&mut self, b: &[u8]) {
        self.0.bytes.extend_from_slice(b);
        self.0.seg_ends.push(self.0.bytes.len() as u64);
    }
    fn finish(&self) -> u64 {
        unreachable!("the recorder never finishes a hash")
    }
}
impl Recorder {
    fn record<'a, T: Hash + 'a, I: IntoIterator<Item = &'a T>>(objects: I) -> Recorder {
        let mut r = Recorder::default();
        r.key_segs.push(0);
        for o in objects {
            o.hash(&mut OneKey(&mut r));
            let n = r.seg_ends.len() as u64;
            r.key_segs.push(n);
        }
        r
    }
}

/// (kind, width) of the key types that need no recording pass: their `Hash` impl is one `write` of the value's bytes.
///
/// The reference's impl block is `impl<'a, T: 'a + Hash + Debug> Mphf<T>` (src/lib.rs:100, :229): no `'static`, so
/// `TypeId` is not available without changing the public bounds (`Mphf<&str>` must keep compiling).  The primitive is
/// recognised by `type_name` (exactly "u64", "i32", ... for primitives; any user type carries its module path) and
/// then CONFIRMED on the first key: what `T::hash` writes must be one segment equal to the key's own bytes.  Anything
/// else takes the recorded-stream path, which is correct for every `T: Hash`.
fn word_kind<T: Hash>(first: Option<&T>) -> Option<(i32, i32)> {
    let (kind, width) = match std::any::type_name::<T>() {
        "u64" | "i64" | "usize" | "isize" => (sys::BPHF_KEY_U64, 8usize),
        "u32" | "i32" => (sys::BPHF_KEY_UINT, 4),
        "u16" | "i16" => (sys::BPHF_KEY_UINT, 2),
        "u8" | "i8" => (sys::BPHF_KEY_UINT, 1),
        _ => return None,
    };
    if std::mem::size_of::<T>() != width {
        return None;
    }
    if let Some(k) = first {
        let r = Recorder::record(std::iter::once(k));
        // SAFETY: T is one of the primitive integers above (size checked): no padding, any bit pattern is a value
        let raw = unsafe { std::slice::from_raw_parts(k as *const T as *const u8, width) };
        if r.seg_ends.len() != 1 || r.bytes.as_slice() != raw {
            return None;
        }
    }
    Some((kind as i32, width as i32))
}

fn check(rc: i32) {
    match rc {
        0 => {}
        sys::BPHF_ERR_GAMMA => panic!("assertion failed: gamma > 1.01"), // src/lib.rs:236,364
        sys::BPHF_ERR_MAX_ITERS => panic!("counldn't find unique hashes"), // src/lib.rs:267,400
        _ => panic!("{}", unsafe { CStr::from_ptr(sys::bphf_last_error()) }.to_string_lossy()),
    }
}

/// Build on the GPU and copy the levels into the reference's own struct.
pub(crate) fn try_new_cuda<T: Hash>(gamma: f64, objects: &[T], starting_seed: Option<u64>) -> Option<Mphf<T>> {
    assert!(gamma > 1.01);
    let mut h: *mut sys::bphf_mphf = std::ptr::null_mut();
    let (has_seed, seed) = (starting_seed.is_some() as i32, starting_seed.unwrap_or(0));
    let rc = if let Some((kind, width)) = word_kind::<T>(objects.first()) {
        unsafe {
            sys::bphf_build_keys(objects.as_ptr() as *const core::ffi::c_void, objects.len() as u64, kind, width, gamma,
                                 has_seed, seed, 0, sys::BPHF_MEM_HOST, &mut h)
        }
    } else {
        let r = Recorder::record(objects.iter());
        unsafe {
            sys::bphf_build_stream(r.bytes.as_ptr(), r.bytes.len() as u64, r.seg_ends.as_ptr(), r.seg_ends.len() as u64,
                                   r.key_segs.as_ptr(), objects.len() as u64, gamma, has_seed, seed, 0,
                                   sys::BPHF_MEM_HOST, &mut h)
        }
    };
    check(rc);
    let mut n_levels = 0u32;
    check(unsafe { sys::bphf_num_levels(h, &mut n_levels) });
    let mut bitvecs = Vec::with_capacity(n_levels as usize);
    for l in 0..n_levels {
        let (mut bits, mut nw, mut nr) = (0u64, 0u64, 0u64);
        check(unsafe { sys::bphf_level_dims(h, l, &mut bits, &mut nw, &mut nr) });
        let mut words = vec![0u64; nw as usize];
        let mut ranks = vec![0u64; nr as usize];
        check(unsafe { sys::bphf_level_export(h, l, words.as_mut_ptr(), ranks.as_mut_ptr(), sys::BPHF_MEM_HOST) });
        bitvecs.push((BitVector::from_words(bits, words), ranks.into_boxed_slice()));
    }
    unsafe { sys::bphf_free(h) };
    Some(Mphf { bitvecs: bitvecs.into_boxed_slice(), phantom: PhantomData })
}

impl<T: Hash> Mphf<T> {
    /// `try_hash` over a slice on the GPU; `u64::MAX` where `try_hash` returns `None`
    pub fn try_hash_many(&self, items: &[T]) -> Vec<u64> {
        // upload the levels once per call here; a production shim caches the handle in a OnceCell next to `bitvecs`
        let bits: Vec<u64> = self.bitvecs.iter().map(|(bv, _)| bv.capacity()).collect();
        let words: Vec<Vec<u64>> = self.bitvecs.iter().map(|(bv, _)| (0..bv.num_words()).map(|i| bv.get_word(i)).collect()).collect();
        let wp: Vec<*const u64> = words.iter().map(|w| w.as_ptr()).collect();
        let rp: Vec<*const u64> = self.bitvecs.iter().map(|(_, r)| r.as_ptr()).collect();
        let wk = word_kind::<T>(items.first());
        let (kind, width) = wk.unwrap_or((sys::BPHF_KEY_STREAM, 0));
        let mut h: *mut sys::bphf_mphf = std::ptr::null_mut();
        check(unsafe { sys::bphf_from_levels_keys(bits.len() as u32, bits.as_ptr(), wp.as_ptr(), rp.as_ptr(), kind, width, 0, &mut h) });
        let mut out = vec![0u64; items.len()];
        let rc = if wk.is_some() {
            unsafe { sys::bphf_hash_batch_keys(h, items.as_ptr() as *const core::ffi::c_void, items.len() as u64, out.as_mut_ptr(),
                                               sys::BPHF_MEM_HOST, std::ptr::null_mut()) }
        } else {
            let r = Recorder::record(items.iter());
            unsafe { sys::bphf_hash_batch_stream(h, r.bytes.as_ptr(), r.bytes.len() as u64, r.seg_ends.as_ptr(), r.seg_ends.len() as u64,
                                                 r.key_segs.as_ptr(), items.len() as u64, out.as_mut_ptr(), sys::BPHF_MEM_HOST,
                                                 std::ptr::null_mut()) }
        };
        unsafe { sys::bphf_free(h) };
        check(rc);
        out
    }
}

// src/bitvector.rs gains (10 lines): wrap exported words as the crate's own storage
// impl BitVector {
//     pub(crate) fn from_words(bits: u64, words: Vec<u64>) -> BitVector {
//         #[cfg(feature = "parallel")]
//         let vector = words.into_iter().map(AtomicU64::new).collect::<Vec<_>>().into_boxed_slice();
//         #[cfg(not(feature = "parallel"))]
//         let vector = words.into_boxed_slice();
//         BitVector { bits, vector }
//     }
// }
