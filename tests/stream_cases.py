"""Seeded key sets of the types that hash as STREAM keys (shared by the CPU and the GPU tests)."""
import random
import string


def _rand_word(rng, lo=0, hi=12):
    return "".join(rng.choice(string.ascii_letters + string.digits + " äöü€") for _ in range(rng.randint(lo, hi)))


def distinct(gen, n, rng):
    seen, out = set(), []
    while len(out) < n:
        v = gen(rng)
        k = repr(v)
        if k not in seen:
            seen.add(k)
            out.append(v)
    return out


def make_keys(rust_type, n, seed=1):
    """n distinct Python values of the Rust type `rust_type`"""
    rng = random.Random(seed * 7919 + n)
    if rust_type == "Vec<u8>":
        # every tail length of the wyhash core (0..31 mod 32), several 32-byte blocks, and the empty vector
        return distinct(lambda r: bytes(r.getrandbits(8) for _ in range(r.choice([0, 1, 2, 3, 7, 8, 9, 15, 16, 17, 23, 24, 25,
                                                                                31, 32, 33, 40, 63, 64, 65, 100, 200]))), n, rng)
    if rust_type == "String":
        return distinct(lambda r: _rand_word(r, 0, 40), n, rng)
    if rust_type == "Vec<String>":
        return distinct(lambda r: [_rand_word(r) for _ in range(r.randint(0, 5))], n, rng)
    if rust_type == "u128":
        return distinct(lambda r: r.getrandbits(128), n, rng)
    if rust_type == "i128":
        return distinct(lambda r: r.getrandbits(128) - (1 << 127), n, rng)
    if rust_type == "[u8; 40]":
        return distinct(lambda r: bytes(r.getrandbits(8) for _ in range(40)), n, rng)
    if rust_type == "(u32, String)":
        return distinct(lambda r: (r.getrandbits(32), _rand_word(r, 0, 9)), n, rng)
    if rust_type == "Vec<u32>":
        return distinct(lambda r: [r.getrandbits(32) for _ in range(r.randint(0, 12))], n, rng)
    if rust_type == "(char, bool, i16)":
        return distinct(lambda r: (chr(r.randint(32, 0x2fff)), bool(r.getrandbits(1)), r.randint(-32768, 32767)), n, rng)
    if rust_type == "u64":
        return distinct(lambda r: r.getrandbits(64), n, rng)
    if rust_type == "u32":
        return distinct(lambda r: r.getrandbits(32), n, rng)
    if rust_type == "[u8; 5]":
        return distinct(lambda r: bytes(r.getrandbits(8) for _ in range(5)), n, rng)
    raise KeyError(rust_type)


STREAM_TYPES = ["Vec<u8>", "String", "Vec<String>", "u128", "i128", "[u8; 40]", "(u32, String)", "Vec<u32>",
                "(char, bool, i16)"]
