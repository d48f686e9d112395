/*
 * boomphf_oracle.c -- CPU restatement of the rust-boomphf MPHF hot path for u64 keys.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (rust-boomphf_b200/csrc, libboomphf_cuda.so) never links, loads or calls it.
 *
 * PARITY STATUS: "parity unpinned" at the wyhash boundary.  The reference
 * (boomphf 0.6.0) takes its hash from the third-party crate `wyhash`
 * (/root/reference/Cargo.toml:19, requirement ">=0.3, <=0.5", no Cargo.lock, resolves to
 * 0.5.0) which is not on disk, and no test of the reference pins a hash value, a
 * bitvector word or a rank.  The hash below restates the published wyhash "v1"
 * algorithm that crate's top-level `WyHash` implements; it is pinned against the two
 * known answers from that crate's README (wyhash(&[0,1,2],3), wyrng(3)) and against the
 * self-generated vectors of SURVEY.md Appendix A.  The one line those answers do not
 * exercise is the 8-byte tail read (swapped 32-bit halves), see BO_WY_SWAP_8BYTE_TAIL.
 * Everything else follows /root/reference/src/lib.rs and src/bitvector.rs line by line
 * (cited at each function).
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BO_API __attribute__((visibility("default")))
#define BO_ERR_ARG_EARLY (-1)

/* 1: wyhash-v1 `read64_swapped` for an exactly-8-byte tail (primary candidate, matches the
 * C wyhash v1 `__wyr64` and the crate's `read_rest` case 8).  0: plain little-endian read.
 * The CUDA side has the same single switch (BPHF_WY_SWAP_8BYTE_TAIL in hash.cuh). */
#ifndef BO_WY_SWAP_8BYTE_TAIL
#define BO_WY_SWAP_8BYTE_TAIL 1
#endif

/* ------------------------------------------------------------------------------------------
 * wyhash v1 (crate `wyhash` 0.3-0.5, module v1; constants P0..P5)
 * ---------------------------------------------------------------------------------------- */
static const uint64_t P0 = 0xa0761d6478bd642fULL;
static const uint64_t P1 = 0xe7037ed1a0b428dbULL;
static const uint64_t P2 = 0x8ebc6af09c88c6e3ULL;
static const uint64_t P3 = 0x589965cc75374cc3ULL;
static const uint64_t P4 = 0x1d8e4e27c47d124fULL;
static const uint64_t P5 = 0xeb44accab455d165ULL;

static inline uint64_t wymum(uint64_t a, uint64_t b) {
    __uint128_t r = (__uint128_t)a * (__uint128_t)b;
    return (uint64_t)(r >> 64) ^ (uint64_t)r;
}
static inline uint64_t rd08(const uint8_t *p) { return p[0]; }
static inline uint64_t rd16(const uint8_t *p) { return (uint64_t)p[0] | ((uint64_t)p[1] << 8); }
static inline uint64_t rd32(const uint8_t *p) {
    return (uint64_t)p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 24);
}
static inline uint64_t rd64(const uint8_t *p) { return rd32(p) | (rd32(p + 4) << 32); }
static inline uint64_t rd64_swapped(const uint8_t *p) { return (rd32(p) << 32) | rd32(p + 4); }
/* 1..8 trailing bytes */
static inline uint64_t rd_rest(const uint8_t *p, size_t len) {
    switch (len) {
    case 1: return rd08(p);
    case 2: return rd16(p);
    case 3: return (rd16(p) << 8) | rd08(p + 2);
    case 4: return rd32(p);
    case 5: return (rd32(p) << 8) | rd08(p + 4);
    case 6: return (rd32(p) << 16) | rd16(p + 4);
    case 7: return (rd32(p) << 24) | (rd16(p + 4) << 8) | rd08(p + 6);
    default:
#if BO_WY_SWAP_8BYTE_TAIL
        return rd64_swapped(p);
#else
        return rd64(p);
#endif
    }
}

/* one `Hasher::write` call: the v1 core over `len` bytes, carrying the state `seed` */
static uint64_t wyhash_core(const uint8_t *p, size_t len, uint64_t seed) {
    size_t i;
    for (i = 0; i + 32 <= len; i += 32, p += 32) {
        seed = wymum(seed ^ P0, wymum(rd64(p) ^ P1, rd64(p + 8) ^ P2) ^
                                    wymum(rd64(p + 16) ^ P3, rd64(p + 24) ^ P4));
    }
    seed ^= P0;
    size_t rest = len & 31;
    if (rest != 0) {
        switch (((len - 1) & 31) / 8) {
        case 0: seed = wymum(seed, rd_rest(p, rest) ^ P1); break;
        case 1: seed = wymum(rd64_swapped(p) ^ seed, rd_rest(p + 8, rest - 8) ^ P2); break;
        case 2:
            seed = wymum(rd64_swapped(p) ^ seed, rd64_swapped(p + 8) ^ P2) ^
                   wymum(seed, rd_rest(p + 16, rest - 16) ^ P3);
            break;
        default:
            seed = wymum(rd64_swapped(p) ^ seed, rd64_swapped(p + 8) ^ P2) ^
                   wymum(rd64_swapped(p + 16) ^ seed, rd_rest(p + 24, rest - 24) ^ P4);
            break;
        }
    }
    return seed;
}
static inline uint64_t wyhash_finish(uint64_t total_len, uint64_t seed) { return wymum(seed, total_len ^ P5); }

/* wyhash::wyhash(bytes, seed) -- exposed for the README known-answer test */
BO_API uint64_t bo_wyhash_bytes(const uint8_t *p, uint64_t len, uint64_t seed) {
    return wyhash_finish(len, wyhash_core(p, (size_t)len, seed));
}
/* one Hasher::write and Hasher::finish, exposed so tests can rebuild any key's hash from its byte stream */
BO_API uint64_t bo_wyhash_core(const uint8_t *p, uint64_t len, uint64_t state) { return wyhash_core(p, (size_t)len, state); }
BO_API uint64_t bo_wyhash_finish(uint64_t total_len, uint64_t state) { return wyhash_finish(total_len, state); }
/* wyhash::wyrng(&mut seed) -- second README known answer (validates constants + wymum) */
BO_API uint64_t bo_wyrng(uint64_t *seed) {
    *seed += P0;
    return wymum(*seed ^ P1, *seed);
}

/* ------------------------------------------------------------------------------------------
 * src/lib.rs:55-88  fold / hash_with_seed / hash_with_seed32 / fastmod / hashmod,  T = u64
 * ---------------------------------------------------------------------------------------- */
/* src/lib.rs:60-65.  `1 << (iter + iter)` on u64: release-mode shift wraps modulo 64
 * (debug builds panic for iter >= 32); u64::hash = one 8-byte native-endian write, then finish. */
/* Key kind: which byte stream `T::hash` feeds the hasher.  Keys are handed to every function of
 * this file as u64 values holding the key's bytes little-endian in their low `width` bytes.
 *   kind 0: u64 / i64 / usize / isize  -> Hasher::write_u64          = one 8-byte write
 *   kind 1: u8 / u16 / u32 (and signed) -> write_u8 / _u16 / _u32     = one write of `width` bytes
 *   kind 2: [u8; N], &[u8], Vec<u8>     -> impl Hash for [T]: write_usize(N) (8 bytes), then one write of N bytes
 * The wyhash Hasher runs the core once per write, carrying h, and finish() mixes in the total byte count.
 * Process-wide setting (test infrastructure; set before building / hashing).
 *   kind 3: any other T: Hash ("stream keys") -- the recorded sequence of Hasher::write calls of every key
 *           (bytes, end offset of every write, writes per key; bo_set_key_stream); the u64 "key" handed to the
 *           functions of this file is then the key's INDEX into those tables.
 * A zero-length write leaves the hasher alone (the crate's write() loops over bytes.chunks(..), which yields nothing
 * for an empty slice): BO_WY_EMPTY_WRITE_IS_NOOP, same switch as BPHF_WY_EMPTY_WRITE_IS_NOOP on the CUDA side. */
#ifndef BO_WY_EMPTY_WRITE_IS_NOOP
#define BO_WY_EMPTY_WRITE_IS_NOOP 1
#endif
static int g_key_kind = 0, g_key_width = 8;
static const uint8_t *g_s_bytes = NULL;
static const uint64_t *g_s_seg_ends = NULL, *g_s_key_segs = NULL;
BO_API int bo_set_key_kind(int kind, int width) {
    if (kind == 0) width = 8;
    if (kind < 0 || kind > 2 || width < 1 || width > 8) return BO_ERR_ARG_EARLY;
    g_key_kind = kind;
    g_key_width = width;
    return 0;
}
BO_API int bo_set_key_stream(const uint8_t *bytes, const uint64_t *seg_ends, const uint64_t *key_segs) {
    if (!key_segs) return BO_ERR_ARG_EARLY;
    g_key_kind = 3;
    g_key_width = 0;
    g_s_bytes = bytes;
    g_s_seg_ends = seg_ends;
    g_s_key_segs = key_segs;
    return 0;
}
BO_API uint64_t bo_hash_with_seed(uint64_t iter, uint64_t key) {
    uint64_t h = 1ULL << ((iter + iter) & 63);
    uint64_t size = 0;
    uint8_t b[8];
    if (g_key_kind == 3) { /* one core per recorded write, state carried; finish(total bytes) */
        const uint64_t s0 = g_s_key_segs[key], s1 = g_s_key_segs[key + 1];
        uint64_t pos = s0 ? g_s_seg_ends[s0 - 1] : 0;
        for (uint64_t sg = s0; sg < s1; sg++) {
            const uint64_t end = g_s_seg_ends[sg];
#if BO_WY_EMPTY_WRITE_IS_NOOP
            if (end != pos)
#endif
                h = wyhash_core(g_s_bytes + pos, (size_t)(end - pos), h);
            size += end - pos;
            pos = end;
        }
        return wyhash_finish(size, h);
    }
    if (g_key_kind == 2) { /* length prefix: usize N, native endian */
        const uint64_t len = (uint64_t)g_key_width;
        for (int i = 0; i < 8; i++) b[i] = (uint8_t)(len >> (8 * i));
        h = wyhash_core(b, 8, h);
        size += 8;
    }
    const size_t w = (size_t)g_key_width;
    for (size_t i = 0; i < w; i++) b[i] = (uint8_t)(key >> (8 * i));
    h = wyhash_core(b, w, h);
    size += w;
    return wyhash_finish(size, h);
}
/* src/lib.rs:55-58 */
static inline uint32_t fold(uint64_t v) { return (uint32_t)(v & 0xFFFFFFFFULL) ^ (uint32_t)(v >> 32); }
/* src/lib.rs:72-75 */
static inline uint64_t fastmod(uint32_t hash, uint32_t n) { return ((uint64_t)hash * (uint64_t)n) >> 32; }
/* src/lib.rs:77-88 */
BO_API uint64_t bo_hashmod(uint64_t iter, uint64_t key, uint64_t n) {
    if (n < (1ULL << 32)) {
        uint32_t h = fold(bo_hash_with_seed(iter, key));
        return fastmod(h, (uint32_t)n);
    } else {
        return bo_hash_with_seed(iter, key) % n;
    }
}

/* ------------------------------------------------------------------------------------------
 * src/bitvector.rs  BitVector (the subset the hot path uses)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t bits;    /* capacity(), src/bitvector.rs:359 */
    uint64_t nwords;  /* u64s(bits) = (bits+63)/64, src/bitvector.rs:426-428 */
    uint64_t *v;
} bo_bitvec;

/* src/bitvector.rs:139-150 -- words pushed one at a time (serial zero fill) */
static int bv_new(bo_bitvec *bv, uint64_t bits) {
    bv->bits = bits;
    bv->nwords = (bits + 63) / 64;
    bv->v = (uint64_t *)malloc((size_t)(bv->nwords ? bv->nwords : 1) * sizeof(uint64_t));
    if (!bv->v) return -1;
    for (uint64_t i = 0; i < bv->nwords; i++) bv->v[i] = 0;
    return 0;
}
static void bv_free(bo_bitvec *bv) { free(bv->v); bv->v = NULL; }
/* src/bitvector.rs:211-214 */
static inline int bv_contains(const bo_bitvec *bv, uint64_t bit) {
    return (__atomic_load_n(&bv->v[bit / 64], __ATOMIC_RELAXED) & (1ULL << (bit % 64))) != 0;
}
/* src/bitvector.rs:253-259  insert = fetch_or(mask, Relaxed); returns "was clear" */
static inline int bv_insert(bo_bitvec *bv, uint64_t bit) {
    uint64_t mask = 1ULL << (bit % 64);
    uint64_t prev = __atomic_fetch_or(&bv->v[bit / 64], mask, __ATOMIC_RELAXED);
    return (prev & mask) == 0;
}
/* src/bitvector.rs:282-292  insert_sync (non-atomic) */
static inline int bv_insert_sync(bo_bitvec *bv, uint64_t bit) {
    uint64_t mask = 1ULL << (bit % 64);
    uint64_t prev = bv->v[bit / 64];
    bv->v[bit / 64] = prev | mask;
    return (prev & mask) == 0;
}
/* src/bitvector.rs:301-307  remove = fetch_and(!mask, Relaxed) */
static inline void bv_remove(bo_bitvec *bv, uint64_t bit) {
    uint64_t mask = 1ULL << (bit % 64);
    __atomic_fetch_and(&bv->v[bit / 64], ~mask, __ATOMIC_RELAXED);
}

/* ------------------------------------------------------------------------------------------
 * src/lib.rs:90-96  Mphf { bitvecs: Box<[(BitVector, Box<[u64]>)]> }
 * ---------------------------------------------------------------------------------------- */
#define BO_MAX_ITERS 100 /* src/lib.rs:98 */
#define BO_MAX_LEVELS (BO_MAX_ITERS + 2)

typedef struct bo_mphf {
    uint32_t nlevels;
    bo_bitvec bv[BO_MAX_LEVELS];
    uint64_t *ranks[BO_MAX_LEVELS];
    uint64_t nranks[BO_MAX_LEVELS];
} bo_mphf;

enum { BO_OK = 0, BO_ERR_ARG = -1, BO_ERR_GAMMA = -2, BO_ERR_MAX_ITERS = -3, BO_ERR_NOMEM = -5 };

BO_API void bo_free(bo_mphf *m) {
    if (!m) return;
    for (uint32_t i = 0; i < m->nlevels; i++) {
        bv_free(&m->bv[i]);
        free(m->ranks[i]);
    }
    free(m);
}

/* src/lib.rs:241,256,369,384   max(255, (gamma * k as f64) as u64)  -- `as u64` truncates toward
 * zero and saturates; gamma*k is far below 2^64 for any k that fits in memory. */
BO_API uint64_t bo_level_size(double gamma, uint64_t k) {
    double s = gamma * (double)k;
    uint64_t u;
    if (!(s > 0.0)) u = 0;
    else if (s >= 18446744073709551616.0) u = UINT64_MAX;
    else u = (uint64_t)s;
    return u < 255 ? 255 : u;
}

/* src/lib.rs:277-297  compute_ranks: ONE running popcount across all levels, one entry per 8 words */
static int compute_ranks(bo_mphf *m) {
    uint64_t pop = 0;
    for (uint32_t l = 0; l < m->nlevels; l++) {
        uint64_t nw = m->bv[l].nwords;
        uint64_t nr = (nw + 7) / 8;
        m->ranks[l] = (uint64_t *)malloc((size_t)(nr ? nr : 1) * sizeof(uint64_t));
        if (!m->ranks[l]) return BO_ERR_NOMEM;
        m->nranks[l] = nr;
        uint64_t r = 0;
        for (uint64_t i = 0; i < nw; i++) {
            if (i % 8 == 0) m->ranks[l][r++] = pop;
            pop += (uint64_t)__builtin_popcountll(m->bv[l].v[i]);
        }
    }
    return BO_OK;
}

/* src/lib.rs:411-464 Context */
typedef struct {
    uint64_t size, seed;
    bo_bitvec a, collide;
} bo_ctx;
static int ctx_new(bo_ctx *cx, uint64_t size, uint64_t seed) {
    cx->size = size;
    cx->seed = seed;
    if (bv_new(&cx->a, size)) return -1;
    if (bv_new(&cx->collide, size)) return -1;
    return 0;
}
/* src/lib.rs:429-434 */
static inline void find_collisions(bo_ctx *cx, uint64_t key) {
    uint64_t idx = bo_hashmod(cx->seed, key, cx->size);
    if (!bv_contains(&cx->collide, idx) && !bv_insert(&cx->a, idx)) bv_insert(&cx->collide, idx);
}
/* src/lib.rs:436-441 */
static inline void find_collisions_sync(bo_ctx *cx, uint64_t key) {
    uint64_t idx = bo_hashmod(cx->seed, key, cx->size);
    if (!bv_contains(&cx->collide, idx) && !bv_insert_sync(&cx->a, idx)) bv_insert_sync(&cx->collide, idx);
}
/* src/lib.rs:443-452  returns 1 when the key must be redone */
static inline int filter(bo_ctx *cx, uint64_t key) {
    uint64_t idx = bo_hashmod(cx->seed, key, cx->size);
    if (bv_contains(&cx->collide, idx)) {
        bv_remove(&cx->a, idx);
        return 1;
    }
    return 0;
}

static bo_mphf *mphf_alloc(void) { return (bo_mphf *)calloc(1, sizeof(bo_mphf)); }

/* src/lib.rs:235-275  Mphf::new -- serial; redo list is a Vec<&T> (pointer vector) */
BO_API int bo_build_serial(const uint64_t *keys, uint64_t n, double gamma, bo_mphf **out) {
    if (!out || (n && !keys)) return BO_ERR_ARG;
    *out = NULL;
    if (!(gamma > 1.01)) return BO_ERR_GAMMA; /* assert!(gamma > 1.01), src/lib.rs:236 */
    bo_mphf *m = mphf_alloc();
    if (!m) return BO_ERR_NOMEM;
    uint64_t iter = 0;
    bo_ctx cx;
    if (ctx_new(&cx, bo_level_size(gamma, n), iter)) { bo_free(m); return BO_ERR_NOMEM; }
    for (uint64_t i = 0; i < n; i++) find_collisions_sync(&cx, keys[i]);
    const uint64_t **redo = (const uint64_t **)malloc((size_t)(n ? n : 1) * sizeof(*redo));
    if (!redo) { bo_free(m); return BO_ERR_NOMEM; }
    uint64_t nredo = 0;
    for (uint64_t i = 0; i < n; i++)
        if (filter(&cx, keys[i])) redo[nredo++] = &keys[i];
    m->bv[m->nlevels++] = cx.a;
    bv_free(&cx.collide);
    iter += 1;
    int rc = BO_OK;
    while (nredo != 0) {
        if (ctx_new(&cx, bo_level_size(gamma, nredo), iter)) { rc = BO_ERR_NOMEM; break; }
        for (uint64_t i = 0; i < nredo; i++) find_collisions_sync(&cx, *redo[i]);
        uint64_t w = 0;
        for (uint64_t i = 0; i < nredo; i++)
            if (filter(&cx, *redo[i])) redo[w++] = redo[i];
        nredo = w;
        m->bv[m->nlevels++] = cx.a;
        bv_free(&cx.collide);
        iter += 1;
        if (iter > BO_MAX_ITERS) { rc = BO_ERR_MAX_ITERS; break; } /* src/lib.rs:265-268 */
    }
    free(redo);
    if (rc == BO_OK) rc = compute_ranks(m);
    if (rc != BO_OK) { bo_free(m); return rc; }
    *out = m;
    return BO_OK;
}

/* parallel filter_map().collect(): per-thread chunks, concatenated in order (rayon's collect keeps
 * order; the result does not depend on it -- SURVEY.md section 3.1) */
static uint64_t par_filter_keys(bo_ctx *cx, const uint64_t *keys, uint64_t n, const uint64_t **dst, int nthreads) {
    uint64_t *counts = (uint64_t *)calloc((size_t)nthreads + 1, sizeof(uint64_t));
    uint8_t *flag = (uint8_t *)malloc((size_t)(n ? n : 1));
    uint64_t total = 0;
#pragma omp parallel num_threads(nthreads)
    {
        int t = 0, nt = 1;
#ifdef _OPENMP
        t = omp_get_thread_num();
        nt = omp_get_num_threads();
#endif
        uint64_t lo = n * (uint64_t)t / (uint64_t)nt, hi = n * (uint64_t)(t + 1) / (uint64_t)nt, c = 0;
        for (uint64_t i = lo; i < hi; i++) { flag[i] = (uint8_t)filter(cx, keys[i]); c += flag[i]; }
        counts[t + 1] = c;
#pragma omp barrier
#pragma omp single
        { for (int j = 0; j < nt; j++) counts[j + 1] += counts[j]; total = counts[nt]; }
        uint64_t w = counts[t];
        for (uint64_t i = lo; i < hi; i++) if (flag[i]) dst[w++] = &keys[i];
    }
    free(counts);
    free(flag);
    return total;
}
static uint64_t par_filter_ptrs(bo_ctx *cx, const uint64_t **src, uint64_t n, const uint64_t **dst, int nthreads) {
    uint64_t *counts = (uint64_t *)calloc((size_t)nthreads + 1, sizeof(uint64_t));
    uint8_t *flag = (uint8_t *)malloc((size_t)(n ? n : 1));
    uint64_t total = 0;
#pragma omp parallel num_threads(nthreads)
    {
        int t = 0, nt = 1;
#ifdef _OPENMP
        t = omp_get_thread_num();
        nt = omp_get_num_threads();
#endif
        uint64_t lo = n * (uint64_t)t / (uint64_t)nt, hi = n * (uint64_t)(t + 1) / (uint64_t)nt, c = 0;
        for (uint64_t i = lo; i < hi; i++) { flag[i] = (uint8_t)filter(cx, *src[i]); c += flag[i]; }
        counts[t + 1] = c;
#pragma omp barrier
#pragma omp single
        { for (int j = 0; j < nt; j++) counts[j + 1] += counts[j]; total = counts[nt]; }
        uint64_t w = counts[t];
        for (uint64_t i = lo; i < hi; i++) if (flag[i]) dst[w++] = src[i];
    }
    free(counts);
    free(flag);
    return total;
}

/* src/lib.rs:363-408  Mphf::new_parallel -- two data-parallel passes per level with relaxed
 * fetch_or / fetch_and on shared u64 words; serial zero-fill and serial compute_ranks as in the
 * reference.  nthreads <= 0 -> all cores (rayon's global pool default). */
BO_API int bo_build_parallel(const uint64_t *keys, uint64_t n, double gamma, int has_seed, uint64_t starting_seed,
                             int nthreads, bo_mphf **out) {
    if (!out || (n && !keys)) return BO_ERR_ARG;
    *out = NULL;
    if (!(gamma > 1.01)) return BO_ERR_GAMMA; /* src/lib.rs:364 */
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    bo_mphf *m = mphf_alloc();
    if (!m) return BO_ERR_NOMEM;
    uint64_t seed0 = has_seed ? starting_seed : 0; /* starting_seed.unwrap_or(0) */
    uint64_t iter = 0;
    bo_ctx cx;
    if (ctx_new(&cx, bo_level_size(gamma, n), seed0 + iter)) { bo_free(m); return BO_ERR_NOMEM; }
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (uint64_t i = 0; i < n; i++) find_collisions(&cx, keys[i]);
    const uint64_t **redo = (const uint64_t **)malloc((size_t)(n ? n : 1) * sizeof(*redo));
    const uint64_t **redo2 = NULL;
    if (!redo) { bo_free(m); return BO_ERR_NOMEM; }
    uint64_t nredo = par_filter_keys(&cx, keys, n, redo, nthreads);
    m->bv[m->nlevels++] = cx.a;
    bv_free(&cx.collide);
    iter += 1;
    int rc = BO_OK;
    if (nredo) {
        redo2 = (const uint64_t **)malloc((size_t)nredo * sizeof(*redo2));
        if (!redo2) rc = BO_ERR_NOMEM;
    }
    while (rc == BO_OK && nredo != 0) {
        if (ctx_new(&cx, bo_level_size(gamma, nredo), seed0 + iter)) { rc = BO_ERR_NOMEM; break; }
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (uint64_t i = 0; i < nredo; i++) find_collisions(&cx, *redo[i]);
        nredo = par_filter_ptrs(&cx, redo, nredo, redo2, nthreads);
        const uint64_t **t = redo; redo = redo2; redo2 = t;
        m->bv[m->nlevels++] = cx.a;
        bv_free(&cx.collide);
        iter += 1;
        if (iter > BO_MAX_ITERS) { rc = BO_ERR_MAX_ITERS; break; } /* src/lib.rs:398-401 */
    }
    free(redo);
    free(redo2);
    if (rc == BO_OK) rc = compute_ranks(m);
    if (rc != BO_OK) { bo_free(m); return rc; }
    *out = m;
    return BO_OK;
}

/* ------------------------------------------------------------------------------------------
 * src/lib.rs:112-226  Mphf::from_chunked_iterator -- serial, keys arrive as a sequence of chunks that is walked
 * twice per level; a `done_keys` bit vector over key positions replaces the redo list.
 * chunks are given as one concatenated array + chunk lengths (n must equal their sum, else the reference
 * panics "max number of items overflowed" or never finishes).
 * ---------------------------------------------------------------------------------------- */
BO_API int bo_build_chunked_serial(const uint64_t *keys, const uint64_t *chunk_lens, uint64_t n_chunks, uint64_t n,
                                   double gamma, bo_mphf **out) {
    if (!out || (n && !keys)) return BO_ERR_ARG;
    *out = NULL;
    uint64_t total = 0;
    for (uint64_t c = 0; c < n_chunks; c++) total += chunk_lens[c];
    if (total != n) return BO_ERR_ARG;
    if (!(gamma > 1.01)) return BO_ERR_GAMMA; /* src/lib.rs:125 */
    bo_mphf *m = mphf_alloc();
    if (!m) return BO_ERR_NOMEM;
    bo_bitvec done;
    if (bv_new(&done, n > 255 ? n : 255)) { bo_free(m); return BO_ERR_NOMEM; }
    uint64_t iter = 0, n_done = 0;
    int rc = BO_OK;
    for (;;) {
        if (iter > BO_MAX_ITERS) { rc = BO_ERR_MAX_ITERS; break; } /* src/lib.rs:128-131 */
        const uint64_t remaining = iter == 0 ? n : n - n_done;
        bo_ctx cx;
        if (ctx_new(&cx, bo_level_size(gamma, remaining), iter)) { rc = BO_ERR_NOMEM; break; }
        uint64_t offset = 0;
        for (uint64_t c = 0; c < n_chunks; c++) { /* first walk: find_collisions of the keys not done yet */
            for (uint64_t i = 0; i < chunk_lens[c]; i++)
                if (!bv_contains(&done, offset + i)) find_collisions_sync(&cx, keys[offset + i]);
            offset += chunk_lens[c];
        }
        offset = 0;
        for (uint64_t c = 0; c < n_chunks; c++) { /* second walk: a.remove on collisions, else the key is done */
            for (uint64_t i = 0; i < chunk_lens[c]; i++) {
                if (bv_contains(&done, offset + i)) continue;
                if (!filter(&cx, keys[offset + i])) { bv_insert_sync(&done, offset + i); n_done++; }
            }
            offset += chunk_lens[c];
        }
        if (m->nlevels >= BO_MAX_LEVELS) { bv_free(&cx.a); bv_free(&cx.collide); rc = BO_ERR_MAX_ITERS; break; }
        m->bv[m->nlevels++] = cx.a;
        bv_free(&cx.collide);
        if (n_done == n) break; /* src/lib.rs:215-217 */
        iter += 1;
    }
    bv_free(&done);
    if (rc == BO_OK) rc = compute_ranks(m);
    if (rc != BO_OK) { bo_free(m); return rc; }
    *out = m;
    return BO_OK;
}

/* ------------------------------------------------------------------------------------------
 * src/lib.rs:539-667  Mphf::from_chunked_iterator_parallel.  Levels are built over the chunks with the caller's
 * gamma until fewer than min(50 M, 1 % of n) keys are left; the colliding keys of that level are buffered and the
 * rest is built by Mphf::new_parallel(1.7, &buffered, Some(iter)) -- gamma 1.7 whatever the caller passed
 * (src/lib.rs:654), seeds continuing at `iter`.  The 1 % threshold is computed in f32 (src/lib.rs:557).
 * (The work-queue / thread plumbing does not influence the result; the level passes run in OpenMP here.)
 * ---------------------------------------------------------------------------------------- */
BO_API int bo_build_chunked_parallel(const uint64_t *keys, const uint64_t *chunk_lens, uint64_t n_chunks, uint64_t n,
                                     double gamma, int has_max_iters, uint64_t max_iters, int nthreads, bo_mphf **out) {
    if (!out || (n && !keys)) return BO_ERR_ARG;
    *out = NULL;
    uint64_t total = 0;
    for (uint64_t c = 0; c < n_chunks; c++) total += chunk_lens[c];
    if (total != n) return BO_ERR_ARG;
    if (!(gamma > 1.01)) return BO_ERR_GAMMA; /* src/lib.rs:562 */
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    const uint64_t MAX_BUFFER_SIZE = 50000000ULL;
    const uint64_t min_buffer_keys_threshold = (uint64_t)(0.01f * (float)n); /* f32 arithmetic, src/lib.rs:556-557 */
    bo_mphf *m = mphf_alloc();
    if (!m) return BO_ERR_NOMEM;
    uint8_t *done = (uint8_t *)calloc((size_t)(n ? n : 1), 1); /* done_keys, one byte per key position */
    uint64_t *buffered = NULL;
    uint64_t n_buffered = 0, n_done = 0, iter = 0;
    int buffer_keys = 0, rc = BO_OK;
    if (!done) { bo_free(m); return BO_ERR_NOMEM; }
    for (;;) {
        if (has_max_iters && iter > max_iters) { rc = BO_ERR_MAX_ITERS; break; } /* src/lib.rs:570-573 */
        const uint64_t remaining = iter == 0 ? n : n - n_done;
        if (remaining == 0) break;
        if (remaining < MAX_BUFFER_SIZE && remaining < min_buffer_keys_threshold) buffer_keys = 1;
        if (m->nlevels >= BO_MAX_LEVELS) { rc = BO_ERR_MAX_ITERS; break; }
        bo_ctx cx;
        if (ctx_new(&cx, bo_level_size(gamma, remaining), iter)) { rc = BO_ERR_NOMEM; break; }
        /* job 0 of every chunk, then job 1 of every chunk (Queue::next waits for all of job 0, src/lib.rs:504-532) */
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (uint64_t i = 0; i < n; i++)
            if (!done[i]) find_collisions(&cx, keys[i]);
        if (buffer_keys) {
            buffered = (uint64_t *)malloc((size_t)(remaining ? remaining : 1) * 8);
            if (!buffered) { bv_free(&cx.a); bv_free(&cx.collide); rc = BO_ERR_NOMEM; break; }
        }
        uint64_t newly_done = 0;
#pragma omp parallel for num_threads(nthreads) schedule(static) reduction(+ : newly_done)
        for (uint64_t i = 0; i < n; i++) {
            if (done[i]) continue;
            if (filter(&cx, keys[i])) {
                if (buffer_keys) {
                    uint64_t pos;
#pragma omp atomic capture
                    pos = n_buffered++;
                    buffered[pos] = keys[i];
                }
            } else {
                done[i] = 1;
                newly_done++;
            }
        }
        n_done += newly_done;
        m->bv[m->nlevels++] = cx.a;
        bv_free(&cx.collide);
        iter += 1;
        if (buffer_keys) break;
    }
    free(done);
    if (rc == BO_OK && n_buffered > 1) { /* src/lib.rs:652-661 */
        bo_mphf *tail = NULL;
        rc = bo_build_parallel(buffered, n_buffered, 1.7, 1, iter, nthreads, &tail);
        if (rc == BO_OK) {
            for (uint32_t l = 0; l < tail->nlevels && rc == BO_OK; l++) {
                if (m->nlevels >= BO_MAX_LEVELS) { rc = BO_ERR_MAX_ITERS; break; }
                m->bv[m->nlevels++] = tail->bv[l];
                tail->bv[l].v = NULL;
            }
            bo_free(tail);
        }
    }
    free(buffered);
    if (rc == BO_OK) rc = compute_ranks(m);
    if (rc != BO_OK) { bo_free(m); return rc; }
    *out = m;
    return BO_OK;
}

/* build an Mphf from already-serialised levels (the serde data model: per level bits, words, ranks) */
BO_API int bo_from_levels(uint32_t nlevels, const uint64_t *bits, const uint64_t *const *words,
                          const uint64_t *const *ranks, bo_mphf **out) {
    if (!out || nlevels > BO_MAX_LEVELS) return BO_ERR_ARG;
    bo_mphf *m = mphf_alloc();
    if (!m) return BO_ERR_NOMEM;
    for (uint32_t l = 0; l < nlevels; l++) {
        if (bv_new(&m->bv[l], bits[l])) { m->nlevels = l; bo_free(m); return BO_ERR_NOMEM; }
        memcpy(m->bv[l].v, words[l], (size_t)m->bv[l].nwords * 8);
        m->nranks[l] = (m->bv[l].nwords + 7) / 8;
        m->ranks[l] = (uint64_t *)malloc((size_t)(m->nranks[l] ? m->nranks[l] : 1) * 8);
        memcpy(m->ranks[l], ranks[l], (size_t)m->nranks[l] * 8);
        m->nlevels = l + 1;
    }
    *out = m;
    return BO_OK;
}

BO_API uint32_t bo_num_levels(const bo_mphf *m) { return m->nlevels; }
BO_API int bo_level_dims(const bo_mphf *m, uint32_t l, uint64_t *bits, uint64_t *nwords, uint64_t *nranks) {
    if (!m || l >= m->nlevels) return BO_ERR_ARG;
    if (bits) *bits = m->bv[l].bits;
    if (nwords) *nwords = m->bv[l].nwords;
    if (nranks) *nranks = m->nranks[l];
    return BO_OK;
}
BO_API const uint64_t *bo_level_words(const bo_mphf *m, uint32_t l) { return l < m->nlevels ? m->bv[l].v : NULL; }
BO_API const uint64_t *bo_level_ranks(const bo_mphf *m, uint32_t l) { return l < m->nlevels ? m->ranks[l] : NULL; }

/* src/lib.rs:299-318 get_rank */
static inline uint64_t get_rank(const bo_mphf *m, uint64_t hash, uint32_t i) {
    uint64_t idx = hash;
    const bo_bitvec *bv = &m->bv[i];
    uint64_t rank = m->ranks[i][idx / 512];
    for (uint64_t j = (idx / 64) & ~7ULL; j < idx / 64; j++) rank += (uint64_t)__builtin_popcountll(bv->v[j]);
    uint64_t final_word = bv->v[idx / 64];
    if (idx % 64 > 0) rank += (uint64_t)__builtin_popcountll(final_word << (64 - (idx % 64)));
    return rank;
}
/* src/lib.rs:324-355  hash / try_hash.  Returns UINT64_MAX where try_hash returns None
 * (and where `hash` would hit unreachable!).  Seed index = level index (src/lib.rs:327,347). */
BO_API uint64_t bo_try_hash(const bo_mphf *m, uint64_t key) {
    for (uint32_t i = 0; i < m->nlevels; i++) {
        const bo_bitvec *bv = &m->bv[i];
        uint64_t h = bo_hashmod((uint64_t)i, key, bv->bits);
        if (bv_contains(bv, h)) return get_rank(m, h, i);
    }
    return UINT64_MAX;
}
BO_API void bo_hash_batch(const bo_mphf *m, const uint64_t *keys, uint64_t n, uint64_t *out, int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    if (nthreads == 1) {
        for (uint64_t i = 0; i < n; i++) out[i] = bo_try_hash(m, keys[i]);
        return;
    }
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (uint64_t i = 0; i < n; i++) out[i] = bo_try_hash(m, keys[i]);
}

/* ------------------------------------------------------------------------------------------
 * src/hashmap.rs:27-40  BoomHashMap::create_map -- in-place cycle-following permutation of
 * parallel key / value arrays (values are u64 here) into MPHF order.
 * ---------------------------------------------------------------------------------------- */
BO_API int bo_create_map(const bo_mphf *m, uint64_t *keys, uint64_t *values, uint64_t n) {
    for (uint64_t i = 0; i < n; i++) {
        for (;;) {
            uint64_t slot = bo_try_hash(m, keys[i]);
            if (slot >= n) return BO_ERR_ARG;
            if (slot == i) break;
            uint64_t t = keys[i]; keys[i] = keys[slot]; keys[slot] = t;
            if (values) { t = values[i]; values[i] = values[slot]; values[slot] = t; }
        }
    }
    return BO_OK;
}
/* src/hashmap.rs:49-66  BoomHashMap::get: try_hash + key equality; writes found flag */
BO_API void bo_map_get_batch(const bo_mphf *m, const uint64_t *map_keys, const uint64_t *map_values, uint64_t n,
                             const uint64_t *queries, uint64_t nq, uint64_t *values_out, uint8_t *found_out) {
    for (uint64_t i = 0; i < nq; i++) {
        uint64_t pos = bo_try_hash(m, queries[i]);
        if (pos != UINT64_MAX && pos < n && map_keys[pos] == queries[i]) {
            values_out[i] = map_values ? map_values[pos] : pos;
            found_out[i] = 1;
        } else {
            values_out[i] = 0;
            found_out[i] = 0;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Level-granular entry points on plain word arrays, used by the tests that model the sharded
 * (multi-GPU) schedule of SURVEY.md 8e on the CPU: every shard marks into its own (a, collide),
 * the pairs are folded with the OR-collide monoid, each shard filters its own keys.
 * ---------------------------------------------------------------------------------------- */
/* Context::find_collisions_sync (src/lib.rs:436-441) for n keys into caller-owned word arrays */
BO_API void bo_mark_level(const uint64_t *keys, uint64_t n, uint64_t seed_index, uint64_t bits, uint64_t *a,
                          uint64_t *collide) {
    bo_ctx cx;
    cx.size = bits;
    cx.seed = seed_index;
    cx.a.bits = cx.collide.bits = bits;
    cx.a.nwords = cx.collide.nwords = (bits + 63) / 64;
    cx.a.v = a;
    cx.collide.v = collide;
    for (uint64_t i = 0; i < n; i++) find_collisions_sync(&cx, keys[i]);
}
/* the redo half of Context::filter (src/lib.rs:443-452): keys whose slot is set in `collide`, in order */
BO_API uint64_t bo_filter_level(const uint64_t *keys, uint64_t n, uint64_t seed_index, uint64_t bits,
                                const uint64_t *collide, uint64_t *redo_out) {
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t idx = bo_hashmod(seed_index, keys[i], bits);
        if ((collide[idx >> 6] >> (idx & 63)) & 1ULL) redo_out[m++] = keys[i];
    }
    return m;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic key generators: bijections of the counter, so keys are
 * distinct by construction.  The CUDA library has the same generators (bphf_keygen_u64).
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64_fin(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
/* kind 0: splitmix64 finaliser of seed + i*golden (bijective on u64)
 * kind 1: 62-bit "2-bit packed 31-mer": a bijection on [0,2^62) (odd multiply + xorshift, masked)
 * kind 2: the literal benches/build.rs input, key = 2*i */
BO_API void bo_keygen(uint64_t *out, uint64_t start, uint64_t n, uint64_t seed, int kind) {
#pragma omp parallel for schedule(static)
    for (uint64_t j = 0; j < n; j++) {
        uint64_t i = start + j, k;
        if (kind == 0) {
            k = splitmix64_fin(seed + i * 0x9E3779B97F4A7C15ULL);
        } else if (kind == 1) {
            const uint64_t M = (1ULL << 62) - 1;
            uint64_t z = (i + seed) & M;
            z = (z * 0x9E3779B97F4A7C15ULL) & M;
            z ^= z >> 31;
            z = (z * 0xbf58476d1ce4e5b9ULL) & M;
            z ^= z >> 29;
            k = z;
        } else {
            k = 2 * i;
        }
        out[j] = k;
    }
}

BO_API int bo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
BO_API int bo_wy_swap_8byte_tail(void) { return BO_WY_SWAP_8BYTE_TAIL; }
