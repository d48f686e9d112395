// boomphf.hpp -- header-only C++17 mirror of the reference's Rust interface over the C ABI (boomphf_cuda.h).
//
// Same names, argument meaning and failure behaviour as /root/reference/src/lib.rs:
//   Mphf<T>::new_(gamma, objects)                          Mphf::new            src/lib.rs:235  ("new" is a C++ keyword)
//   Mphf<T>::new_parallel(gamma, objects, starting_seed)   Mphf::new_parallel   src/lib.rs:363
//   Mphf<T>::from_chunked_iterator[_parallel](...)         src/lib.rs:112, :539   (T = uint64_t)
//   hash(item)       Mphf::hash      src/lib.rs:324   -- throws Panic("...must find a hash value") on a miss
//   try_hash(item)   Mphf::try_hash  src/lib.rs:340   -- std::optional
//   try_hash_many(items)                               -- the batched form the GPU is for (BPHF_NONE = None)
//   levels() / from_levels()                           -- the serde data model: (bits, words, ranks) per level
// A Rust panic becomes a boomphf::Panic exception carrying the C status code and the reference's message.
//
// T: the integer types (u64-like: one 8-byte write; 1/2/4-byte integers: one write of their width), std::array<uint8_t, N>
// with N <= 8 ([u8; N]), and -- through the recorded-writes path (BPHF_KEY_STREAM) -- std::string (Rust String / &str:
// bytes, then 0xff), std::vector<uint8_t> (Vec<u8>: usize length, then the bytes) and std::vector<std::string>
// (Vec<String>).  Other types: specialise boomphf::StreamWriter<T> with the writes their Rust `Hash` impl makes.
//
// This header is host code only; everything it calls lives in libboomphf_cuda.so.  No CPU fallback exists: without an
// sm_100 device every constructor throws Panic(BPHF_ERR_CUDA).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "boomphf_cuda.h"

namespace boomphf {

struct Panic : std::runtime_error {
    int code;
    Panic(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

namespace detail {
inline void check(int rc) {
    if (rc == BPHF_OK) return;
    const char* e = bphf_last_error();
    throw Panic(rc, (e && *e) ? std::string(e) : ("boomphf_cuda error " + std::to_string(rc)));
}

// the recorded Hasher::write calls of a key set
struct Recorder {
    std::vector<uint8_t> bytes;
    std::vector<uint64_t> seg_ends;
    std::vector<uint64_t> key_segs{0};
    void write(const void* p, size_t n) {
        const uint8_t* b = static_cast<const uint8_t*>(p);
        bytes.insert(bytes.end(), b, b + n);
        seg_ends.push_back(bytes.size());
    }
    void write_usize(uint64_t v) { write(&v, 8); }  // little-endian hosts only (x86-64 / aarch64), like the GPU side
    void write_u8(uint8_t v) { write(&v, 1); }
    void end_key() { key_segs.push_back(seg_ends.size()); }
};
}  // namespace detail

// ---- how a T reaches the hasher ------------------------------------------------------------------------------
// word keys: kind / width for bphf_build_keys (no recording pass)
template <class T, class = void>
struct WordKey : std::false_type {};
template <class T>
struct WordKey<T, std::enable_if_t<std::is_integral_v<T> && !std::is_same_v<T, bool>>> : std::true_type {
    static constexpr int kind = sizeof(T) == 8 ? BPHF_KEY_U64 : BPHF_KEY_UINT;
    static constexpr int width = (int)sizeof(T);
    static_assert(sizeof(T) == 1 || sizeof(T) == 2 || sizeof(T) == 4 || sizeof(T) == 8, "128-bit integers: StreamWriter");
};
template <size_t N>
struct WordKey<std::array<uint8_t, N>, std::enable_if_t<(N >= 1 && N <= 8)>> : std::true_type {
    static constexpr int kind = BPHF_KEY_BYTES;
    static constexpr int width = (int)N;
};

// stream keys: the writes T's Rust `Hash` impl makes
template <class T, class = void>
struct StreamWriter;  // specialise: static void write(detail::Recorder&, const T&)
template <>
struct StreamWriter<std::string> {  // impl Hash for str: write(bytes); write_u8(0xff)
    static void write(detail::Recorder& r, const std::string& s) {
        r.write(s.data(), s.size());
        r.write_u8(0xff);
    }
};
template <class E>
struct StreamWriter<std::vector<E>, std::enable_if_t<std::is_integral_v<E> && !std::is_same_v<E, bool>>> {
    // impl Hash for [u8] / [u32] / ...: write_usize(len); ONE write of the elements' memory (hash_slice of the integers)
    static void write(detail::Recorder& r, const std::vector<E>& v) {
        r.write_usize(v.size());
        r.write(v.data(), v.size() * sizeof(E));
    }
};
template <size_t N>
struct StreamWriter<std::array<uint8_t, N>, std::enable_if_t<(N > 8)>> {  // [u8; N] hashes as its slice
    static void write(detail::Recorder& r, const std::array<uint8_t, N>& v) {
        r.write_usize(N);
        r.write(v.data(), N);
    }
};
template <class E>
struct StreamWriter<std::vector<E>, std::enable_if_t<!std::is_integral_v<E>>> {
    static void write(detail::Recorder& r, const std::vector<E>& v) {  // impl Hash for [T]: length, then every element
        r.write_usize(v.size());
        for (const E& e : v) StreamWriter<E>::write(r, e);
    }
};
template <class A, class B>
struct StreamWriter<std::pair<A, B>> {  // tuples: the fields in order
    template <class X>
    static void one(detail::Recorder& r, const X& x) {
        if constexpr (std::is_integral_v<X>) r.write(&x, sizeof(X));
        else StreamWriter<X>::write(r, x);
    }
    static void write(detail::Recorder& r, const std::pair<A, B>& p) {
        one(r, p.first);
        one(r, p.second);
    }
};

template <class T>
class Mphf {
   public:
    struct Level {  // (BitVector { bits, vector }, Box<[u64]>) of the reference, src/lib.rs:90-96
        uint64_t bits = 0;
        std::vector<uint64_t> words, ranks;
    };

    Mphf() = default;
    Mphf(const Mphf&) = delete;
    Mphf& operator=(const Mphf&) = delete;
    Mphf(Mphf&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Mphf& operator=(Mphf&& o) noexcept {
        if (this != &o) {
            bphf_free(h_);
            h_ = o.h_;
            o.h_ = nullptr;
        }
        return *this;
    }
    ~Mphf() { bphf_free(h_); }

    /// Mphf::new(gamma, objects)  (src/lib.rs:235)
    static Mphf new_(double gamma, const std::vector<T>& objects, int device = 0) {
        return build(gamma, objects.data(), objects.size(), std::nullopt, device);
    }
    /// Mphf::new_parallel(gamma, objects, starting_seed)  (src/lib.rs:363)
    static Mphf new_parallel(double gamma, const std::vector<T>& objects, std::optional<uint64_t> starting_seed = std::nullopt,
                             int device = 0) {
        return build(gamma, objects.data(), objects.size(), starting_seed, device);
    }
    /// Mphf::from_chunked_iterator(gamma, &objects, n)  (src/lib.rs:112), T = u64-like
    static Mphf from_chunked_iterator(double gamma, const std::vector<std::vector<T>>& objects, uint64_t n, int device = 0) {
        return build_chunked(gamma, objects, n, 0, std::nullopt, device);
    }
    /// Mphf::from_chunked_iterator_parallel(gamma, &objects, max_iters, n, num_threads)  (src/lib.rs:539)
    static Mphf from_chunked_iterator_parallel(double gamma, const std::vector<std::vector<T>>& objects,
                                               std::optional<uint64_t> max_iters, uint64_t n, size_t /*num_threads*/ = 0,
                                               int device = 0) {
        return build_chunked(gamma, objects, n, 1, max_iters, device);
    }

    /// try_hash over a slice on the GPU; BPHF_NONE where the reference returns None
    std::vector<uint64_t> try_hash_many(const T* items, size_t n) const {
        std::vector<uint64_t> out(n);
        if (n == 0) return out;
        if constexpr (WordKey<T>::value) {
            detail::check(bphf_hash_batch_keys(h_, items, n, out.data(), BPHF_MEM_HOST, nullptr));
        } else {
            detail::Recorder r;
            for (size_t i = 0; i < n; i++) {
                StreamWriter<T>::write(r, items[i]);
                r.end_key();
            }
            detail::check(bphf_hash_batch_stream(h_, r.bytes.data(), r.bytes.size(), r.seg_ends.data(), r.seg_ends.size(),
                                                 r.key_segs.data(), n, out.data(), BPHF_MEM_HOST, nullptr));
        }
        return out;
    }
    std::vector<uint64_t> try_hash_many(const std::vector<T>& items) const { return try_hash_many(items.data(), items.size()); }
    /// Mphf::try_hash (src/lib.rs:340)
    std::optional<uint64_t> try_hash(const T& item) const {
        const uint64_t v = try_hash_many(&item, 1)[0];
        return v == BPHF_NONE ? std::nullopt : std::optional<uint64_t>(v);
    }
    /// Mphf::hash (src/lib.rs:324): an item outside the construction set may panic
    uint64_t hash(const T& item) const {
        const auto v = try_hash(item);
        if (!v) throw Panic(BPHF_ERR_ARG, "internal error: entered unreachable code: must find a hash value");
        return *v;
    }

    /// the serde data model
    std::vector<Level> levels() const {
        uint32_t n = 0;
        detail::check(bphf_num_levels(h_, &n));
        std::vector<Level> out(n);
        for (uint32_t l = 0; l < n; l++) {
            uint64_t nw = 0, nr = 0;
            detail::check(bphf_level_dims(h_, l, &out[l].bits, &nw, &nr));
            out[l].words.resize(nw);
            out[l].ranks.resize(nr);
            detail::check(bphf_level_export(h_, l, out[l].words.data(), out[l].ranks.data(), BPHF_MEM_HOST));
        }
        return out;
    }
    static Mphf from_levels(const std::vector<Level>& lv, int device = 0) {
        std::vector<uint64_t> bits;
        std::vector<const uint64_t*> w, r;
        for (const Level& l : lv) {
            bits.push_back(l.bits);
            w.push_back(l.words.data());
            r.push_back(l.ranks.data());
        }
        Mphf m;
        int kind = BPHF_KEY_STREAM, width = 0;
        if constexpr (WordKey<T>::value) {
            kind = WordKey<T>::kind;
            width = WordKey<T>::width;
        }
        detail::check(bphf_from_levels_keys((uint32_t)lv.size(), bits.data(), w.data(), r.data(), kind, width, device, &m.h_));
        return m;
    }
    const bphf_mphf* handle() const { return h_; }

   private:
    bphf_mphf* h_ = nullptr;

    static Mphf build(double gamma, const T* objects, size_t n, std::optional<uint64_t> seed, int device) {
        Mphf m;
        if constexpr (WordKey<T>::value) {
            detail::check(bphf_build_keys(objects, n, WordKey<T>::kind, WordKey<T>::width, gamma, seed.has_value() ? 1 : 0,
                                          seed.value_or(0), device, BPHF_MEM_HOST, &m.h_));
        } else {
            detail::Recorder r;
            for (size_t i = 0; i < n; i++) {
                StreamWriter<T>::write(r, objects[i]);
                r.end_key();
            }
            detail::check(bphf_build_stream(r.bytes.data(), r.bytes.size(), r.seg_ends.data(), r.seg_ends.size(),
                                            r.key_segs.data(), n, gamma, seed.has_value() ? 1 : 0, seed.value_or(0), device,
                                            BPHF_MEM_HOST, &m.h_));
        }
        return m;
    }
    static Mphf build_chunked(double gamma, const std::vector<std::vector<T>>& objects, uint64_t n, int parallel,
                              std::optional<uint64_t> max_iters, int device) {
        static_assert(WordKey<T>::value && sizeof(T) == 8, "the chunked constructors take u64-like keys");
        std::vector<const uint64_t*> ptr;
        std::vector<uint64_t> len;
        for (const auto& c : objects) {
            ptr.push_back(reinterpret_cast<const uint64_t*>(c.data()));
            len.push_back(c.size());
        }
        Mphf m;
        detail::check(bphf_build_chunked_u64(ptr.data(), len.data(), ptr.size(), n, gamma, parallel, max_iters.has_value() ? 1 : 0,
                                             max_iters.value_or(0), device, BPHF_MEM_HOST, &m.h_));
        return m;
    }
};

}  // namespace boomphf
