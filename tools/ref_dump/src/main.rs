//! ref_dump -- prints, as JSON, what the real `boomphf` 0.6.0 (+ the `wyhash` crate it resolves to) computes for the
//! inputs of tests/golden/vectors.json:
//!
//!   cd tools/ref_dump && cargo run --release > ../../tests/golden/ref_dump.json
//!
//! * `hash_with_seed`: the reference's private function (src/lib.rs:60-65) is three lines around the wyhash crate;
//!   they are repeated here verbatim in meaning (`WyHash::with_seed(1 << (iter + iter))`, `v.hash()`, `finish()`), so
//!   the values pin the crate's 8-byte tail read -- the one thing the oracle could not check offline.
//! * `mphf`: `Mphf::new` / `Mphf::new_parallel` through the public API, read back through serde: per level `bits`,
//!   the words and ranks (small sets in full; all sets as FNV-1a over the little-endian bytes of bits | words | ranks),
//!   plus `hash()` of the first keys.
//! The key generators are the ones of oracle/boomphf_oracle.c (`bo_keygen`).
use std::hash::{Hash, Hasher};

use boomphf::Mphf;
use serde_json::{json, Value};

fn hash_with_seed<T: Hash + ?Sized>(iter: u64, v: &T) -> u64 {
    // release-mode semantics of `1 << (iter + iter)` (the shift amount wraps at 64; debug builds would panic at iter >= 32)
    let mut state = wyhash::WyHash::with_seed(1u64.wrapping_shl((iter + iter) as u32));
    v.hash(&mut state);
    state.finish()
}

fn splitmix64_fin(mut z: u64) -> u64 {
    z = (z ^ (z >> 30)).wrapping_mul(0xbf58476d1ce4e5b9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94d049bb133111eb);
    z ^ (z >> 31)
}

/// bo_keygen(out, 0, n, seed, kind)
fn keygen(n: u64, seed: u64, kind: u32) -> Vec<u64> {
    (0..n)
        .map(|i| match kind {
            0 => splitmix64_fin(seed.wrapping_add(i.wrapping_mul(0x9E3779B97F4A7C15))),
            1 => {
                let m: u64 = (1u64 << 62) - 1;
                let mut z = (i.wrapping_add(seed)) & m;
                z = z.wrapping_mul(0x9E3779B97F4A7C15) & m;
                z ^= z >> 31;
                z = z.wrapping_mul(0xbf58476d1ce4e5b9) & m;
                z ^= z >> 29;
                z
            }
            _ => 2 * i,
        })
        .collect()
}

fn fnv1a(h: &mut u64, x: u64) {
    for b in x.to_le_bytes() {
        *h ^= b as u64;
        *h = h.wrapping_mul(0x100000001b3);
    }
}

/// levels of an Mphf through its serde data model: {"bitvecs": [[{"bits": .., "vector": [..]}, [ranks..]], ..]}
fn levels(m: &Mphf<u64>) -> Vec<(u64, Vec<u64>, Vec<u64>)> {
    let v: Value = serde_json::to_value(m).expect("serde");
    v["bitvecs"]
        .as_array()
        .expect("bitvecs")
        .iter()
        .map(|lv| {
            let bits = lv[0]["bits"].as_u64().unwrap();
            let words = lv[0]["vector"].as_array().unwrap().iter().map(|x| x.as_u64().unwrap()).collect();
            let ranks = lv[1].as_array().unwrap().iter().map(|x| x.as_u64().unwrap()).collect();
            (bits, words, ranks)
        })
        .collect()
}

fn mphf_entry(name: &str, gamma: f64, keys: &[u64], seed: Option<u64>, parallel: bool) -> Value {
    let m: Mphf<u64> = if parallel { Mphf::new_parallel(gamma, keys, seed) } else { Mphf::new(gamma, keys) };
    let lv = levels(&m);
    let mut h: u64 = 0xcbf29ce484222325;
    for (bits, words, ranks) in &lv {
        fnv1a(&mut h, *bits);
        words.iter().for_each(|w| fnv1a(&mut h, *w));
        ranks.iter().for_each(|r| fnv1a(&mut h, *r));
    }
    let mut e = json!({
        "name": name, "gamma": gamma, "starting_seed": seed, "n": keys.len(), "constructor": if parallel { "new_parallel" } else { "new" },
        "bits": lv.iter().map(|l| l.0).collect::<Vec<_>>(),
        "popcounts": lv.iter().map(|l| l.1.iter().map(|w| w.count_ones() as u64).sum::<u64>()).collect::<Vec<_>>(),
        "fnv1a64": format!("{:016x}", h),
    });
    if seed.is_none() {
        // Mphf::hash uses the plain level index as the seed (src/lib.rs:327): only meaningful without a starting seed
        e["hash_first"] = json!(keys.iter().take(6).map(|k| m.hash(k)).collect::<Vec<_>>());
    }
    if keys.len() <= 5000 {
        e["levels"] = json!(lv
            .iter()
            .map(|(b, w, r)| json!({"bits": b,
                "words": w.iter().map(|x| format!("{:016x}", x)).collect::<Vec<_>>(),
                "ranks": r}))
            .collect::<Vec<_>>());
    }
    e
}

fn main() {
    let hkeys: [u64; 6] = [0, u64::MAX, 1, 2, 0x0123456789ABCDEF, 1999998];
    let iters: [u64; 9] = [0, 1, 5, 31, 32, 33, 63, 64, 100];
    let mut hs = Vec::new();
    for k in hkeys {
        for it in iters {
            hs.push(json!({"key": format!("{:016x}", k), "iter": it, "h": format!("{:016x}", hash_with_seed(it, &k))}));
        }
    }
    // other key types of SURVEY 8f row 4: what their Hash impls feed the hasher
    let typed = json!({
        "u32_0x89abcdef_iter0": format!("{:016x}", hash_with_seed(0, &0x89abcdefu32)),
        "u8_7_iter3": format!("{:016x}", hash_with_seed(3, &7u8)),
        "bytes5_iter1": format!("{:016x}", hash_with_seed(1, &[1u8, 2, 3, 4, 5])),
        "vec_u8_abc_iter0": format!("{:016x}", hash_with_seed(0, &b"abc".to_vec())),
        "string_hello_iter2": format!("{:016x}", hash_with_seed(2, &String::from("hello"))),
        "u128_iter0": format!("{:016x}", hash_with_seed(0, &0x0123456789abcdef_fedcba9876543210u128)),
    });
    let even: Vec<u64> = (0..1000u64).map(|i| 2 * i).collect();
    let ms = vec![
        mphf_entry("even_1000_g1.7", 1.7, &even, None, false),
        mphf_entry("even_1000_g2.0", 2.0, &even, None, false),
        mphf_entry("doctest_9", 1.7, &[1, 10, 1000, 23, 457, 856, 845, 124, 912], None, false),
        mphf_entry("one", 1.7, &[42], None, false),
        mphf_entry("two", 1.7, &[42, 43], None, false),
        mphf_entry("n149", 1.7, &keygen(149, 3, 0), None, false),
        mphf_entry("n150", 1.7, &keygen(150, 3, 0), None, false),
        mphf_entry("rand_5000_g1.7", 1.7, &keygen(5000, 1, 0), None, false),
        mphf_entry("rand_100k_g1.7", 1.7, &keygen(100000, 1, 0), None, false),
        mphf_entry("rand_100k_g1.7_parallel", 1.7, &keygen(100000, 1, 0), None, true),
        mphf_entry("rand_100k_g2.0", 2.0, &keygen(100000, 2, 0), None, false),
        mphf_entry("rand_100k_g5.0", 5.0, &keygen(100000, 3, 0), None, false),
        mphf_entry("rand_20k_g1.02", 1.02, &keygen(20000, 4, 0), None, false),
        mphf_entry("kmer31_100k_g1.7", 1.7, &keygen(100000, 5, 1), None, false),
        mphf_entry("bench_even_1M_g2.0", 2.0, &keygen(1000000, 0, 2), None, false),
        mphf_entry("rand_1M_g1.7", 1.7, &keygen(1000000, 1, 0), None, true),
        mphf_entry("rand_100k_g1.7_seed5", 1.7, &keygen(100000, 1, 0), Some(5), true),
        mphf_entry("rand_100k_g1.7_seed30", 1.7, &keygen(100000, 1, 0), Some(30), true),
    ];
    let out = json!({
        "source": "boomphf 0.6.0 through cargo; wyhash as resolved by the reference's own range (>=0.3, <=0.5)",
        "hash_with_seed": hs,
        "hash_with_seed_typed": typed,
        "mphf": ms,
    });
    println!("{}", serde_json::to_string_pretty(&out).unwrap());
}
