"""Regenerates tests/golden/vectors.json from the CPU oracle.

Provenance of each block:
  "readme_kat"  (R) known answers recalled from the `wyhash` crate README -- the only vectors that come
                from outside this repo; they pin constants, wymum, seed^=P0, the <8-byte tail, finish.
  everything else (S) self-generated from the oracle = the restated reference algorithm; the
                reference's own tests hold no golden vector for this path ("parity unpinned").
                The Appendix-A numbers of SURVEY.md were produced by an independent throw-away
                restatement during the survey and are reproduced here unchanged.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import boomphf_oracle as o  # noqa: E402


def popcounts(levels):
    return [int(sum(bin(int(w)).count("1") for w in l[1])) for l in levels]


def mphf_entry(name, keys, gamma, seed=None, probe=6):
    m = o.OracleMphf.new_parallel(gamma, keys, seed) if seed is not None else o.OracleMphf.new(gamma, keys)
    lv = m.levels()
    e = {"name": name, "gamma": gamma, "starting_seed": seed, "n": int(len(keys)),
         "bits": [int(l[0]) for l in lv], "popcounts": popcounts(lv), "sha256": o.fingerprint(lv)}
    if seed is None:
        e["hash_first"] = [int(m.hash(int(k))) for k in keys[:probe]]
    return e


def main():
    g = {"wy_swap_8byte_tail": 1}
    g["readme_kat"] = {"wyhash_bytes_012_seed3": "b0f941520b1ad95d", "wyrng_seed3": "03e99a772750dcbe"}
    keys = [0, 0xFFFFFFFFFFFFFFFF, 1, 2, 0x0123456789ABCDEF, 1999998]
    iters = [0, 1, 5, 31, 32, 33, 63, 64, 100]
    g["hash_with_seed"] = [{"key": "%016x" % k, "iter": it, "h": "%016x" % o.hash_with_seed(it, k)}
                           for k in keys for it in iters]
    triples = [(0x0123456789ABCDEF, 0, 255), (0x0123456789ABCDEF, 0, 1700000000), (0x0123456789ABCDEF, 0, 5000000000),
               (42, 3, 170000000), (42, 3, 2**32 - 1), (42, 3, 2**32), (42, 40, 2**32 + 12345), (7, 99, 255)]
    g["hashmod"] = [{"key": "%016x" % k, "iter": it, "n": n, "idx": o.hashmod(it, k, n)} for k, it, n in triples]
    g["level_size"] = [{"gamma": ga, "k": k, "size": o.level_size(ga, k)}
                       for ga, k in [(1.7, 0), (1.7, 149), (1.7, 150), (1.7, 151), (1.7, 10**9), (2.0, 10**6),
                                     (5.0, 5 * 10**8), (1.02, 12345678901), (1.7, 3), (1000.0, 4300000)]]
    ms = []
    ms.append(mphf_entry("even_1000_g1.7", np.arange(1000, dtype=np.uint64) * 2, 1.7))
    ms.append(mphf_entry("even_1000_g2.0", np.arange(1000, dtype=np.uint64) * 2, 2.0))
    ms.append(mphf_entry("doctest_9", np.array([1, 10, 1000, 23, 457, 856, 845, 124, 912], dtype=np.uint64), 1.7))
    ms.append(mphf_entry("empty", np.zeros(0, dtype=np.uint64), 1.7))
    ms.append(mphf_entry("one", np.array([42], dtype=np.uint64), 1.7))
    ms.append(mphf_entry("two", np.array([42, 43], dtype=np.uint64), 1.7))
    ms.append(mphf_entry("n149", o.keygen(149, 3, 0), 1.7))
    ms.append(mphf_entry("n150", o.keygen(150, 3, 0), 1.7))
    ms.append(mphf_entry("rand_5000_g1.7", o.keygen(5000, 1, 0), 1.7))
    ms.append(mphf_entry("rand_100k_g1.7", o.keygen(100000, 1, 0), 1.7))
    ms.append(mphf_entry("rand_100k_g2.0", o.keygen(100000, 2, 0), 2.0))
    ms.append(mphf_entry("rand_100k_g5.0", o.keygen(100000, 3, 0), 5.0))
    ms.append(mphf_entry("rand_20k_g1.02", o.keygen(20000, 4, 0), 1.02))
    ms.append(mphf_entry("kmer31_100k_g1.7", o.keygen(100000, 5, 1), 1.7))
    ms.append(mphf_entry("bench_even_1M_g2.0", o.keygen(1000000, 0, 2), 2.0))
    ms.append(mphf_entry("rand_1M_g1.7", o.keygen(1000000, 1, 0), 1.7))
    ms.append(mphf_entry("rand_100k_g1.7_seed5", o.keygen(100000, 1, 0), 1.7, seed=5))
    ms.append(mphf_entry("rand_100k_g1.7_seed30", o.keygen(100000, 1, 0), 1.7, seed=30))
    g["mphf"] = ms
    # key kinds and the chunked-parallel constructor (8f row 3): fingerprints of whole structures
    def kind_keys(name):
        rng = np.random.default_rng(2026)
        if name == "u32":
            return np.unique(rng.integers(0, 2**32, size=60000, dtype=np.uint64)).astype(np.uint32)[:50000]
        if name == "i32":
            return (np.unique(rng.integers(0, 2**32, size=60000, dtype=np.uint64)).astype(np.uint32)[:50000]).view(np.int32)
        if name == "u16":
            return rng.permutation(65536)[:40000].astype(np.uint16)
        if name == "u8":
            return rng.permutation(256)[:200].astype(np.uint8)
        w = int(name[5:])
        b = np.unique(rng.integers(0, 256, size=(30000, w), dtype=np.uint8), axis=0)
        return b[:20000] if w > 1 else b
    ks = []
    for name in ["u32", "i32", "u16", "u8", "bytes1", "bytes3", "bytes5", "bytes8"]:
        k = kind_keys(name)
        m = o.OracleMphf.new(1.7, k)
        ks.append({"name": name, "n": int(len(k)), "gamma": 1.7, "bits": [int(l[0]) for l in m.levels()],
                   "sha256": o.fingerprint(m.levels()), "hash_first": [int(v) for v in m.hash_batch(k[:4])]})
    g["key_kinds"] = ks
    ck = o.keygen(200000, 6, 0)
    cm = o.OracleMphf.from_chunked_iterator_parallel(2.0, np.array_split(ck, 5), None, len(ck))
    g["chunked_parallel"] = {"n": 200000, "keygen_seed": 6, "gamma": 2.0, "bits": [int(l[0]) for l in cm.levels()],
                             "sha256": o.fingerprint(cm.levels())}
    # stream keys (any T: Hash as its recorded writes): raw hashes of hand-written write sequences, then whole
    # structures for the seeded key sets of tests/stream_cases.py
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    import importlib
    import __graft_entry__ as ge
    from stream_cases import STREAM_TYPES, make_keys
    K = importlib.import_module(ge.load_package().__name__ + ".keys")
    raw = []
    seqs = [[b"ab"], [b"", b"x"], [b"x", b""], [bytes(range(40))], [bytes(range(32))], [bytes(range(8)), b"\xff"],
            [(2).to_bytes(8, "little"), b"ab", b"\xff", b"c", b"\xff"], [bytes(range(16))], [bytes(range(100)), bytes(range(7))]]
    for w in seqs:
        ends, pos = [], 0
        for x in w:
            pos += len(x)
            ends.append(pos)
        st = o.Stream(np.frombuffer(b"".join(w), dtype=np.uint8), ends, [0, len(w)])
        raw.append({"writes": [x.hex() for x in w], "h": {str(it): "%016x" % o.hash_stream_with_seed(it, st) for it in (0, 1, 7, 33)}})
    g["stream_hashes"] = raw
    sk_entries = []
    for t in STREAM_TYPES:
        sk = K.encode_keys(make_keys(t, 3000, seed=1), t)
        st = o.Stream(sk.data, sk.seg_ends, sk.key_segs)
        m = o.OracleMphf.new_stream(1.7, st)
        sk_entries.append({"rust_type": t, "n": 3000, "keys_seed": 1, "gamma": 1.7, "bits": [int(l[0]) for l in m.levels()],
                           "sha256": o.fingerprint(m.levels()), "hash_first": [int(v) for v in m.hash_batch_stream(st)[:4]]})
    g["stream_keys"] = sk_entries
    g["keygen"] = {"kind0_seed1_first4": ["%016x" % int(v) for v in o.keygen(4, 1, 0)],
                   "kind1_seed5_first4": ["%016x" % int(v) for v in o.keygen(4, 5, 1)]}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json")
    with open(path, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
