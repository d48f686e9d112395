// Exercises include/boomphf.hpp (the C++ mirror of the reference's Rust interface) end to end: the reference's own test
// property -- hashes of the construction set are a permutation of 0..n-1 (src/lib.rs:721-727) -- for u64, u32, [u8; 5],
// String, Vec<u8> and Vec<String> keys, the chunked constructors, the data-model round trip and the panics.
// Exit code 0: all good; 3: no usable GPU (the library has no CPU fallback); 1: a check failed.
#include <algorithm>
#include <cstdio>
#include <numeric>
#include <string>

#include "boomphf.hpp"

using boomphf::Mphf;

template <class T>
static bool is_permutation_of_range(const Mphf<T>& m, const std::vector<T>& keys) {
    std::vector<uint64_t> h = m.try_hash_many(keys);
    std::sort(h.begin(), h.end());
    for (size_t i = 0; i < h.size(); i++)
        if (h[i] != i) return false;
    return true;
}

#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) {                                                   \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                    \
        }                                                                \
    } while (0)

int main() {
    try {
        // the literal bench input of the reference: keys 2i (benches/build.rs:11-12)
        std::vector<uint64_t> k64(200000);
        for (size_t i = 0; i < k64.size(); i++) k64[i] = 2 * i;
        auto m64 = Mphf<uint64_t>::new_(1.7, k64);
        CHECK(is_permutation_of_range(m64, k64));
        auto m64p = Mphf<uint64_t>::new_parallel(1.7, k64, std::nullopt);
        CHECK(m64p.hash(k64[5]) == m64.hash(k64[5]));           // new == new_parallel(None), bit for bit
        CHECK(!m64.try_hash(1).has_value() || *m64.try_hash(1) < k64.size());  // absent key: None or an arbitrary index
        // serde data model round trip
        auto lv = m64.levels();
        CHECK(!lv.empty() && lv[0].bits == 340000 && lv[0].words.size() == (lv[0].bits + 63) / 64);
        auto m64b = Mphf<uint64_t>::from_levels(lv);
        CHECK(m64b.try_hash_many(k64) == m64.try_hash_many(k64));
        // chunked constructors
        std::vector<std::vector<uint64_t>> chunks{{k64.begin(), k64.begin() + 7}, {}, {k64.begin() + 7, k64.end()}};
        auto mc = Mphf<uint64_t>::from_chunked_iterator(1.7, chunks, k64.size());
        CHECK(mc.try_hash_many(k64) == m64.try_hash_many(k64));
        auto mcp = Mphf<uint64_t>::from_chunked_iterator_parallel(1.7, chunks, std::nullopt, k64.size(), 2);
        CHECK(is_permutation_of_range(mcp, k64));
        // other integer widths and short byte arrays
        std::vector<uint32_t> k32(50000);
        for (size_t i = 0; i < k32.size(); i++) k32[i] = (uint32_t)(i * 2654435761u);
        CHECK(is_permutation_of_range(Mphf<uint32_t>::new_(1.7, k32), k32));
        std::vector<std::array<uint8_t, 5>> kb(30000);
        for (size_t i = 0; i < kb.size(); i++) kb[i] = {(uint8_t)i, (uint8_t)(i >> 8), (uint8_t)(i >> 16), 7, (uint8_t)(i * 31)};
        CHECK(is_permutation_of_range(Mphf<std::array<uint8_t, 5>>::new_parallel(2.0, kb), kb));
        // stream keys: String, Vec<u8>, Vec<String>
        std::vector<std::string> ks;
        for (int i = 0; i < 20000; i++) ks.push_back("key-" + std::to_string(i * 7) + std::string(i % 40, 'x'));
        auto ms = Mphf<std::string>::new_(1.7, ks);
        CHECK(is_permutation_of_range(ms, ks));
        CHECK(ms.hash(ks[123]) < ks.size());
        std::vector<std::vector<uint8_t>> kv;
        for (int i = 0; i < 20000; i++) {
            std::vector<uint8_t> v(i % 50);
            for (size_t j = 0; j < v.size(); j++) v[j] = (uint8_t)(i + j);
            v.push_back((uint8_t)i);
            v.push_back((uint8_t)(i >> 8));
            kv.push_back(v);
        }
        CHECK(is_permutation_of_range(Mphf<std::vector<uint8_t>>::new_(1.7, kv), kv));
        std::vector<std::vector<std::string>> kvs;
        for (int i = 0; i < 5000; i++) kvs.push_back({std::to_string(i), i % 3 ? "a" : "", std::string(i % 5, 'b')});
        CHECK(is_permutation_of_range(Mphf<std::vector<std::string>>::new_parallel(1.7, kvs), kvs));
        // the reference's panics
        bool threw = false;
        try {
            Mphf<uint64_t>::new_(1.0, k64);  // assert!(gamma > 1.01)
        } catch (const boomphf::Panic& p) {
            threw = p.code == BPHF_ERR_GAMMA;
        }
        CHECK(threw);
        threw = false;
        try {
            Mphf<std::string>::new_(1.7, {"dup", "dup", "x"});  // panic!("counldn't find unique hashes")
        } catch (const boomphf::Panic& p) {
            threw = p.code == BPHF_ERR_MAX_ITERS && std::string(p.what()).find("counldn't find unique hashes") != std::string::npos;
        }
        CHECK(threw);
    } catch (const boomphf::Panic& p) {
        std::fprintf(stderr, "panic (%d): %s\n", p.code, p.what());
        return p.code == BPHF_ERR_CUDA ? 3 : 1;
    }
    std::puts("mirror_example: all checks passed");
    return 0;
}
