"""Stream keys: `Mphf<T>` for any `T: Hash` (include/boomphf_cuda.h, BPHF_KEY_STREAM).

`hash_with_seed` (/root/reference/src/lib.rs:60-65) hands `T::hash` a wyhash `Hasher`; all the hasher ever sees is a
sequence of `write(&[u8])` calls.  The Rust shim records those calls with a recording `Hasher` and
passes them through the C ABI; this module is the Python-side stand-in for that recorder: it spells out what the
standard library's `Hash` impls write for the common key types, so tests and Python users can build the same tables.

    u8..u128, i8..i128, usize, isize   one write of the value's little-endian bytes          (Hasher::write_uN)
    bool                               write_u8(0 | 1)
    char                               write_u32(code point)
    String, &str, str                  write(utf-8 bytes), write_u8(0xff)                     (Hasher::write_str)
    Vec<T>, &[T], [T], [T; N]          write_usize(len), then hash_slice:
                                         integer T: ONE write of all elements' bytes (impl_write! hash_slice)
                                         any other T: every element in turn
    (A, B, ...)                        every field in turn, no prefix
    &T, Box<T>                         as T

Python values: int for integers / bool / char (or a 1-character str for char), bytes or str for strings, bytes /
sequences for slices of integers, sequences for other slices and tuples.
"""
import numpy as np

_INTS = {"u8": (1, False), "u16": (2, False), "u32": (4, False), "u64": (8, False), "u128": (16, False),
         "usize": (8, False), "i8": (1, True), "i16": (2, True), "i32": (4, True), "i64": (8, True),
         "i128": (16, True), "isize": (8, True)}


class StreamKeys:
    """The recorded writes of n keys: `data` (uint8, all writes back to back), `seg_ends` (u64, end offset of every
    write), `key_segs` (u64[n+1], writes of key i = [key_segs[i], key_segs[i+1]))."""

    def __init__(self, data, seg_ends, key_segs, rust_type=None):
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        self.seg_ends = np.ascontiguousarray(seg_ends, dtype=np.uint64)
        self.key_segs = np.ascontiguousarray(key_segs, dtype=np.uint64)
        self.rust_type = rust_type
        if self.key_segs.ndim != 1 or self.key_segs.size < 1:
            raise ValueError("key_segs needs n + 1 entries")

    def __len__(self):
        return int(self.key_segs.size) - 1

    def writes_of(self, i):
        """the list of byte strings key i wrote (for tests and debugging)"""
        s0, s1 = int(self.key_segs[i]), int(self.key_segs[i + 1])
        out, pos = [], int(self.seg_ends[s0 - 1]) if s0 else 0
        for j in range(s0, s1):
            end = int(self.seg_ends[j])
            out.append(self.data[pos:end].tobytes())
            pos = end
        return out


# ---- a small parser for Rust type names ---------------------------------------------------------------------
def _split_top(s, sep):
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "<([":
            depth += 1
        elif ch in ">)]":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return [p.strip() for p in parts]


def parse_type(t):
    """'Vec<String>' -> ('slice', ('str',)); 'u32' -> ('int', 4, False); '(u8, String)' -> ('tuple', [...]) ..."""
    t = t.strip()
    while t.startswith("&"):
        t = t[1:].strip()
        if t.startswith("'"):  # lifetime
            t = t.split(None, 1)[1].strip()
        if t.startswith("mut "):
            t = t[4:].strip()
    if t in _INTS:
        return ("int",) + _INTS[t]
    if t == "bool":
        return ("bool",)
    if t == "char":
        return ("char",)
    if t in ("String", "str"):
        return ("str",)
    if t.startswith("Box<") and t.endswith(">"):
        return parse_type(t[4:-1])
    if t.startswith("Vec<") and t.endswith(">"):
        return ("slice", parse_type(t[4:-1]))
    if t.startswith("[") and t.endswith("]"):
        inner = _split_top(t[1:-1], ";")
        return ("slice", parse_type(inner[0]))  # [T; N] hashes as its slice, length prefix included
    if t.startswith("(") and t.endswith(")"):
        body = t[1:-1].strip()
        fields = [p for p in _split_top(body, ",") if p] if body else []
        return ("tuple", [parse_type(p) for p in fields])
    raise TypeError("unsupported Rust key type %r" % t)


def _int_bytes(v, width, signed):
    return int(v).to_bytes(width, "little", signed=signed)


def _writes(ty, v, out):
    """append the byte strings `v: ty` writes into a Hasher"""
    tag = ty[0]
    if tag == "int":
        out.append(_int_bytes(v, ty[1], ty[2]))
    elif tag == "bool":
        out.append(b"\x01" if v else b"\x00")
    elif tag == "char":
        out.append(_int_bytes(ord(v) if isinstance(v, str) else v, 4, False))
    elif tag == "str":
        out.append(v.encode("utf-8") if isinstance(v, str) else bytes(v))
        out.append(b"\xff")
    elif tag == "slice":
        elem = ty[1]
        out.append(_int_bytes(len(v), 8, False))  # write_length_prefix = write_usize
        if elem[0] == "int":  # integer slices hash as one write of their memory (impl_write! hash_slice)
            if isinstance(v, (bytes, bytearray)) and elem[1] == 1:
                out.append(bytes(v))
            else:
                out.append(b"".join(_int_bytes(x, elem[1], elem[2]) for x in v))
        else:
            for x in v:
                _writes(elem, x, out)
    elif tag == "tuple":
        if len(v) != len(ty[1]):
            raise ValueError("tuple arity does not match the type")
        for fty, x in zip(ty[1], v):
            _writes(fty, x, out)
    else:
        raise TypeError(ty)


def infer_type(value):
    """a Rust type for plain Python values: bytes -> Vec<u8>, str -> String, sequence of str -> Vec<String>"""
    if isinstance(value, (bytes, bytearray)):
        return "Vec<u8>"
    if isinstance(value, str):
        return "String"
    if isinstance(value, (list, tuple)) and all(isinstance(x, str) for x in value):
        return "Vec<String>"
    raise TypeError("cannot infer a Rust key type for %r; pass rust_type" % type(value).__name__)


def encode_keys(values, rust_type=None):
    """Record what `T::hash` writes for every key of `values` (T = rust_type) -> StreamKeys"""
    values = list(values)
    if rust_type is None:
        rust_type = infer_type(values[0]) if values else "Vec<u8>"
    ty = parse_type(rust_type)
    chunks, seg_lens, key_segs = [], [], [0]
    for v in values:
        w = []
        _writes(ty, v, w)
        chunks.extend(w)
        seg_lens.extend(len(x) for x in w)
        key_segs.append(len(seg_lens))
    data = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    seg_ends = np.cumsum(np.asarray(seg_lens, dtype=np.uint64), dtype=np.uint64) if seg_lens else np.zeros(0, np.uint64)
    return StreamKeys(data, seg_ends, np.asarray(key_segs, dtype=np.uint64), rust_type)


def is_stream_like(objects):
    """plain Python key collections that only the stream path can take (bytes / str / sequences of str)"""
    if isinstance(objects, StreamKeys):
        return True
    if isinstance(objects, (list, tuple)) and objects:
        x = objects[0]
        return isinstance(x, (bytes, bytearray, str)) or (
            isinstance(x, (list, tuple)) and len(x) > 0 and all(isinstance(e, str) for e in x))
    return False
