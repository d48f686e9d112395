/*
 * boomphf_cuda.h -- C ABI of the B200-native (sm_100a) MPHF build / batched lookup path.
 *
 * This is the drop-in boundary for the hot path of 10XGenomics/rust-boomphf (crate boomphf
 * 0.6.0).  The reference has no FFI of its own; the boundary is its Rust public API
 * (citations are file:line into the reference tree):
 *
 *   Mphf::new(gamma, &[T])                       src/lib.rs:235   -> bphf_build_u64
 *   Mphf::new_parallel(gamma, &[T], Option<u64>) src/lib.rs:363   -> bphf_build_u64
 *   Mphf::hash / Mphf::try_hash                  src/lib.rs:324,340 -> bphf_hash_batch_u64
 *   Mphf { bitvecs: Box<[(BitVector, Box<[u64]>)]> } (serde data model)
 *                                                src/lib.rs:90-96, src/bitvector.rs:41-54
 *                                                -> bphf_num_levels / bphf_level_dims /
 *                                                   bphf_level_export / bphf_from_levels
 *   BoomHashMap::create_map / get                src/hashmap.rs:27-40,49-66
 *                                                -> bphf_map_permute_u64 / bphf_map_get_batch_u64
 *
 * A thin Rust shim binds exactly these symbols; all functions are plain C,
 * take plain pointers and sizes, return 0 on success and a negative status on error, and
 * never unwind across the boundary.  There is NO CPU fallback: every compute entry point
 * fails with BPHF_ERR_CUDA when no sm_100 device is usable.
 *
 * Results are bit-exact with the reference algorithm for T = u64 (same per-level bit
 * vectors, same rank arrays, same hash value for every key) -- see DESIGN.md for the
 * parity argument and its one unpinned assumption (the third-party wyhash 8-byte read).
 */
#ifndef BOOMPHF_CUDA_H
#define BOOMPHF_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BPHF_API __attribute__((visibility("default")))
#else
#define BPHF_API
#endif

typedef struct bphf_mphf bphf_mphf; /* opaque; owns device memory on one GPU */

enum {
    BPHF_OK = 0,
    BPHF_ERR_ARG = -1,       /* null / inconsistent argument                                     */
    BPHF_ERR_GAMMA = -2,     /* reference: assert!(gamma > 1.01)            src/lib.rs:236,364   */
    BPHF_ERR_MAX_ITERS = -3, /* reference: panic!("counldn't find unique hashes") :267,400       */
    BPHF_ERR_CUDA = -4,      /* CUDA runtime error or no usable device (no CPU fallback exists)  */
    BPHF_ERR_NOMEM = -5,     /* device or host allocation failed                                 */
    BPHF_ERR_INTERNAL = -6   /* invariant violated (would be a bug)                              */
};

/* bphf_hash_batch_u64 writes this where Mphf::try_hash returns None
 * (Mphf::hash would hit unreachable!("must find a hash value"), src/lib.rs:334). */
#define BPHF_NONE UINT64_MAX

#define BPHF_MAX_ITERS 100u /* src/lib.rs:98 */
#define BPHF_MAX_LEVELS 101u

/* where a pointer argument lives */
enum { BPHF_MEM_HOST = 0, BPHF_MEM_DEVICE = 1 };

/* ---- library ------------------------------------------------------------------------- */
BPHF_API int bphf_abi_version(void);
BPHF_API const char *bphf_last_error(void); /* thread-local text of the last failure */
BPHF_API int bphf_device_count(int *count);
/* 1 when the 8-byte wyhash tail is read with swapped 32-bit halves (the primary candidate,
 * see DESIGN.md "parity unpinned"); must equal the oracle's bo_wy_swap_8byte_tail(). */
BPHF_API int bphf_wy_swap_8byte_tail(void);

/* ---- build: Mphf::new / Mphf::new_parallel for T = u64 ---------------------------------
 * keys            n distinct u64 keys (borrowed, never modified), host or device memory
 * gamma           must be > 1.01
 * has_seed/seed   Option<u64> starting_seed of new_parallel (None and Mphf::new: has_seed=0).
 *                 Level l is built with seed index seed+l (src/lib.rs:370,385); lookups always
 *                 use seed index l (src/lib.rs:327) exactly like the reference.
 * device          CUDA device ordinal
 * keys_mem        BPHF_MEM_HOST or BPHF_MEM_DEVICE (device pointer must be on `device`)
 * stream          cudaStream_t to run on, or NULL for a library-owned stream.  The call is
 *                 blocking either way (it returns after the structure is complete).
 * The result does not depend on key order, thread or GPU schedule. */
BPHF_API int bphf_build_u64(const uint64_t *keys, uint64_t n, double gamma, int has_seed, uint64_t seed,
                            int device, int keys_mem, void *stream, bphf_mphf **out);

/* ---- keys other than u64 ----------------------------------------------
 * Mphf<T> for the T whose Hash byte stream fits one wyhash tail word (the types of the reference's own
 * quickcheck tests, src/lib.rs:857-879, minus strings):
 *   BPHF_KEY_U64   (width ignored)  u64, i64, usize, isize            -- identical to bphf_build_u64
 *   BPHF_KEY_UINT  width 1 | 2 | 4  u8/i8, u16/i16, u32/i32           -- one write of `width` bytes
 *   BPHF_KEY_BYTES width 1..8       [u8; N], &[u8], Vec<u8> of length N -- write_usize(N), then N bytes
 * keys: n keys packed back to back, `width` bytes each.  Lookups take keys of the handle's kind. */
enum { BPHF_KEY_U64 = 0, BPHF_KEY_UINT = 1, BPHF_KEY_BYTES = 2, BPHF_KEY_STREAM = 3 };
BPHF_API int bphf_build_keys(const void *keys, uint64_t n, int key_kind, int key_width, double gamma, int has_seed,
                             uint64_t seed, int device, int keys_mem, bphf_mphf **out);
BPHF_API int bphf_hash_batch_keys(const bphf_mphf *m, const void *keys, uint64_t n, uint64_t *out, int mem,
                                  void *stream);
BPHF_API int bphf_from_levels_keys(uint32_t n_levels, const uint64_t *bits, const uint64_t *const *words,
                                   const uint64_t *const *ranks, int key_kind, int key_width, int device,
                                   bphf_mphf **out);
BPHF_API int bphf_key_kind(const bphf_mphf *m, int *key_kind, int *key_width);

/* ---- any other T: Hash -- stream keys (BPHF_KEY_STREAM) ------------------------------------
 * hash_with_seed (src/lib.rs:60-65) feeds `T::hash` to a wyhash Hasher; the hasher only ever sees a sequence of
 * `write(&[u8])` calls (u128: one 16-byte write; Vec<u8> / &[u8] / [u8; N]: write_usize(len) then one write of the
 * bytes; String / &str: write(bytes) then write_u8(0xff); Vec<String>: write_usize(len) then that per element; tuples:
 * their fields in order; ...).  The Rust shim records those calls with a recording `Hasher` and hands them over:
 *   bytes[n_bytes]       all writes of all keys back to back
 *   seg_ends[n_segs]     end offset (exclusive) of write j in `bytes`; write j starts where write j-1 ends
 *   key_segs[n_keys+1]   key i made writes [key_segs[i], key_segs[i+1]);  key_segs[0] = 0, key_segs[n_keys] = n_segs
 * The device runs the wyhash v1 core once per write, carrying the state, and finish(total bytes of the key) -- what
 * the crate's Hasher does -- so this covers the remaining quickcheck types of the reference (Vec<String>, Vec<u8>,
 * src/lib.rs:851-879) and every other Hash impl.  A zero-length write leaves the state unchanged.
 * All three tables live in `mem` (host or device).  Malformed tables -> BPHF_ERR_ARG.  Duplicate keys end in
 * BPHF_ERR_MAX_ITERS like any other kind.  bphf_from_levels_keys(..., BPHF_KEY_STREAM, 0, ...) re-creates a handle;
 * the data model and the level / rank export are the same as for u64 keys. */
BPHF_API int bphf_build_stream(const uint8_t *bytes, uint64_t n_bytes, const uint64_t *seg_ends, uint64_t n_segs,
                               const uint64_t *key_segs, uint64_t n_keys, double gamma, int has_seed, uint64_t seed,
                               int device, int mem, bphf_mphf **out);
/* Mphf::try_hash over a batch of stream keys: out[i] = index of key i or BPHF_NONE */
BPHF_API int bphf_hash_batch_stream(const bphf_mphf *m, const uint8_t *bytes, uint64_t n_bytes, const uint64_t *seg_ends,
                                    uint64_t n_segs, const uint64_t *key_segs, uint64_t n_keys, uint64_t *out, int mem,
                                    void *stream);

/* ---- chunked construction ---------------------------------------------
 * Mphf::from_chunked_iterator(gamma, &chunks, n)                       src/lib.rs:112   (parallel_variant = 0)
 * Mphf::from_chunked_iterator_parallel(gamma, &chunks, max_iters, n, num_threads)  :539 (parallel_variant = 1)
 * for T = u64: n_chunks arrays (all host or all device memory) whose lengths add up to n.  Results follow the
 * reference bit for bit, including the parallel variant's quirks: its levels use `gamma` only until fewer than
 * min(50 M, 1 % of n) keys are left and 1.7 afterwards (src/lib.rs:583-585, 654), an empty input yields zero
 * levels, and max_iters only guards the chunked levels.  The key set is gathered into HBM (it must fit the GPU);
 * the out-of-core re-iteration of the reference is not reproduced. */
BPHF_API int bphf_build_chunked_u64(const uint64_t *const *chunks, const uint64_t *chunk_lens, uint64_t n_chunks,
                                    uint64_t n, double gamma, int parallel_variant, int has_max_iters,
                                    uint64_t max_iters, int device, int keys_mem, bphf_mphf **out);

/* ---- sharded build over several GPUs --------------------------------------
 * The reference's new_parallel shares one (a, collide) pair per level between all rayon workers
 * (src/lib.rs:373-377).  Here the key set is split into `world` shards, one per GPU (one process or
 * thread each); every shard marks its keys into its own full-width pair and the pairs are merged per
 * level with the OR-collide monoid (a,c)+(a',c') = (a|a', c|c'|(a&a')) by a peer-memory (NVLink)
 * reduce-scatter / all-gather kernel.  Every shard ends up with the identical, complete structure
 * (a replica for lookups), bit-exact with an unsharded build of the union of the shards.
 *
 * Life cycle: every shard calls bphf_comm_create (allocates its peer-visible arena), publishes its
 * 64-byte handle (bphf_comm_ipc_handle; any host transport, e.g. torch.distributed all_gather),
 * connects with the handles of all shards in rank order (bphf_comm_connect_ipc), then all shards
 * call bphf_build_sharded_u64 collectively (same n_global, gamma, seed; same number of calls).
 * Shards living in ONE process (threads, one or several devices) connect with
 * bphf_comm_connect_local instead.  world <= 8.  After an error status the comm must be destroyed. */
typedef struct bphf_comm bphf_comm;
#define BPHF_IPC_HANDLE_BYTES 64
/* the arena is sized for builds of up to max_keys_global keys (all shards together) at gamma <= max_gamma */
BPHF_API int bphf_comm_create(int device, int rank, int world, uint64_t max_keys_global, double max_gamma,
                              bphf_comm **out);
BPHF_API int bphf_comm_ipc_handle(const bphf_comm *c, void *handle_out /* BPHF_IPC_HANDLE_BYTES */);
BPHF_API int bphf_comm_connect_ipc(bphf_comm *c, const void *handles /* world * BPHF_IPC_HANDLE_BYTES */);
BPHF_API int bphf_comm_connect_local(bphf_comm *c, bphf_comm *const *all /* world comms, rank order */);
BPHF_API void bphf_comm_destroy(bphf_comm *c);
/* keys: this shard's n_local keys (host or device memory on the comm's device); n_global = sum of all shards'
 * n_local.  Blocking; collective over the comm. */
BPHF_API int bphf_build_sharded_u64(bphf_comm *c, const uint64_t *keys, uint64_t n_local, uint64_t n_global,
                                    double gamma, int has_seed, uint64_t seed, int keys_mem, bphf_mphf **out);

/* ---- data model: (BitVector{bits, vector}, ranks) per level ---------------------------- */
BPHF_API int bphf_num_levels(const bphf_mphf *m, uint32_t *n_levels);
BPHF_API int bphf_level_dims(const bphf_mphf *m, uint32_t level, uint64_t *bits, uint64_t *n_words,
                             uint64_t *n_ranks);
/* copy one level out: n_words u64 words (BitVector.vector) and n_ranks u64 ranks */
BPHF_API int bphf_level_export(const bphf_mphf *m, uint32_t level, uint64_t *words_out, uint64_t *ranks_out,
                               int out_mem);
/* all levels at once into caller buffers laid out level after level (no padding):
 * words_out has sum(n_words) entries, ranks_out sum(n_ranks) */
BPHF_API int bphf_export_all(const bphf_mphf *m, uint64_t *words_out, uint64_t *ranks_out, int out_mem);
BPHF_API int bphf_total_dims(const bphf_mphf *m, uint64_t *total_words, uint64_t *total_ranks, uint64_t *n_keys);
/* upload a deserialised Mphf (host arrays) to `device` */
BPHF_API int bphf_from_levels(uint32_t n_levels, const uint64_t *bits, const uint64_t *const *words,
                              const uint64_t *const *ranks, int device, bphf_mphf **out);
/* Releases the handle and its device memory (stream-ordered on a library-owned stream).  Lookups that were launched
 * on this handle with a caller-supplied stream and device buffers return before they finish: they must have
 * completed (or that stream must have been synchronised) before the handle is freed. */
BPHF_API void bphf_free(bphf_mphf *m);

/* ---- lookup: Mphf::hash / try_hash over a batch ----------------------------------------
 * out[i] = rank of keys[i], or BPHF_NONE when no level matched.  keys/out both live in
 * `mem`.  Re-entrant on a shared handle. */
BPHF_API int bphf_hash_batch_u64(const bphf_mphf *m, const uint64_t *keys, uint64_t n, uint64_t *out, int mem,
                                 void *stream);

/* ---- BoomHashMap helpers -------------------------------------------
 * create_map (src/hashmap.rs:27-40): out-of-place scatter into MPHF order,
 *   keys_out[hash(k)] = k, values_out[hash(k)] = v  (values may be NULL). */
BPHF_API int bphf_map_permute_u64(const bphf_mphf *m, const uint64_t *keys, const uint64_t *values, uint64_t n,
                                  uint64_t *keys_out, uint64_t *values_out, int mem, void *stream);
/* get (src/hashmap.rs:49-66): pos = try_hash(q); found = pos != None && map_keys[pos] == q;
 * values_out[i] = map_values[pos] (0 when not found; pos itself when map_values is NULL =
 * get_key_id, src/hashmap.rs:88-105), found_out[i] = 0/1. */
BPHF_API int bphf_map_get_batch_u64(const bphf_mphf *m, const uint64_t *map_keys, const uint64_t *map_values,
                                    uint64_t n_map, const uint64_t *queries, uint64_t n_queries,
                                    uint64_t *values_out, uint8_t *found_out, int mem, void *stream);

/* ---- build statistics (measurement only) ------------------------------------------------ */
typedef struct bphf_build_stats {
    uint32_t n_levels;
    uint32_t tail_level;                   /* first level finished by the single-CTA tail kernel  */
    uint64_t keys_in[BPHF_MAX_LEVELS + 1]; /* keys entering each level                             */
    uint64_t bits[BPHF_MAX_LEVELS + 1];
    uint32_t kernel_launches;              /* kernels launched by the build                        */
    uint32_t host_syncs;                   /* host round trips                                     */
    float pass_ms[BPHF_MAX_LEVELS + 1];    /* per-level mark(+filter) kernel time, profile>=1: level 0 only; 2: all */
    float fin_ms[BPHF_MAX_LEVELS + 1];     /* per-level finalize kernel time (profile 2)            */
    float tail_ms, ranks_ms, total_ms;     /* profile 2 (total_ms always)                           */
    float h2d_ms;                          /* reserved                                              */
    uint32_t bucketed_levels;              /* levels built by the shared-memory (bucketed) kernels; sharded builds:
                                            * + 1000 * levels built by the exchange (owner computes) kernels */
    uint32_t rebuilds;                     /* 1 when a bucket overflow forced an unbucketed rebuild */
    float merge_ms[BPHF_MAX_LEVELS + 1];   /* sharded build: barrier + OR-collide merge + barrier (profile 2) */
    float cascade_ms;                      /* persistent kernel that runs every level >= 1 (build mode 0)     */
    uint32_t streamed_ingest;              /* 1: the host keys went through the device ring (bphf_set_ingest_mode) */
    uint64_t peak_device_bytes;            /* high-water mark of the device's default memory pool over the build
                                              (streamed builds and profile 2; else 0)                             */
} bphf_build_stats;
BPHF_API int bphf_get_build_stats(const bphf_mphf *m, bphf_build_stats *out);
/* 0: no events; 1: CUDA events around the kernels of levels 0 and 1; 2: around every kernel.
 * pass_ms[l] = entry step of level l (scatter / filter+mark), fin_ms[l] = completion step (bucket mark / finalize) */
BPHF_API int bphf_set_profile(int level);
/* construction strategy (all produce identical structures):
 *   mode 0 (default) = L2-atomic kernels; every level >= 1 runs inside one persistent, cooperatively launched
 *                      kernel with grid-wide barriers (no per-level launch gaps);
 *   mode 1           = L2-atomic kernels, two launches per level (what a sharded build uses between its merges);
 *   mode 2           = large levels are built in shared memory after grouping the keys by bit range ("bucketed"),
 *                      levels with fewer than bucket_min_keys keys (0 = default 2^20) by the mode-1 kernels. */
BPHF_API int bphf_set_build_mode(int mode, uint64_t bucket_min_keys);
/* sharded builds (mode 0): levels with more than `bits` bits are built "owner computes" -- every shard routes its keys'
 * slot offsets (4 bytes per key over NVLink) to the shard that owns the slot range, the owners mark their L2-resident
 * slices and return one survivor bit per key -- instead of marking full-width levels on every shard and folding them
 * with the OR-collide merge.  Default 4.5e8 (a full-width (a, collide) pair beyond the L2); 0 = never.  Identical
 * results either way; call it with the same value on every shard. */
BPHF_API int bphf_set_xchg_bits(double bits);
/* ingest of HOST key sets (bphf_build_u64 and bphf_build_chunked_u64 with BPHF_MEM_HOST; identical results):
 *   mode 0 (default) = automatic: a key set whose 8 n bytes plus the work buffers of an in-core build do not fit the
 *                      device's free memory is STREAMED -- the host keys pass through a 3-slot device ring twice (mark of
 *                      level 0; filter of level 0 + mark of level 1) and only the ~45 % that survive level 0 ever become
 *                      device resident; everything else is copied whole, once, and built in core;
 *   mode 1           = always streamed (tests, memory-constrained callers);  mode 2 = never.
 * ring_keys = keys per ring slot (0 = default 16 Mi = 128 MiB).  This is the out-of-core side of
 * Mphf::from_chunked_iterator[_parallel] (src/lib.rs:103-111: "the iterator is re-created once per level"): here the
 * source is read twice in total, not once per level. */
BPHF_API int bphf_set_ingest_mode(int mode, uint64_t ring_keys);
/* lookup strategy (measurement only; identical results): 0 (default) = level-synchronous kernel over the packed
 * lookup structure, 1 = one thread per key over the same structure */
BPHF_API int bphf_set_query_mode(int mode);
/* Lookups against structures larger than the L2 (BASELINE configs[3]; Mphf::hash, src/lib.rs:324-335, over a batch):
 * the queries are grouped by the slice of level 0 (then level 1, ...) their slot falls in, every slice is probed while
 * it is L2 resident, and the ranks are put back in the caller's order -- identical results, no random DRAM probe.
 * mode 0 (default) = used for device-resident batches >= 32 Mi keys against >= 200 MB of packed lookup structure
 * (~3.5e8 keys at gamma 1.7: below that the L2 still serves the plain kernel better), 1 = never, 2 = whenever the
 * structure has a packed form and more than one level (tests).  The other arguments tune it, 0 = default:
 * slice_shift = log2 of the bits of a level per slice (28), walk_bytes = the trailing levels that fit this many bytes are
 * walked unpartitioned (40 MiB), batch_keys = queries per pass (2^27; scratch is ~55 B per key and stage),
 * cap_scale = slice list capacity relative to its plan (1.0; below 1 forces the overflow path: the batch is then redone
 * by the plain kernel, on the device, without a host round trip). */
BPHF_API int bphf_set_query_partition(int mode, uint32_t slice_shift, uint64_t walk_bytes, uint64_t batch_keys,
                                      double cap_scale);
/* counters of the partitioned lookup: batches launched by this process, batches redone by the plain kernel on `device`
 * (synchronises the device), stages of the last plan.  Any pointer may be null. */
BPHF_API int bphf_query_partition_stats(int device, uint64_t *batches, uint64_t *fallbacks, uint32_t *stages);

/* ---- utilities used by tests and bench (device side) ------------------------------------ */
/* distinct-by-construction synthetic keys: kind 0 = splitmix64 finaliser of
 * seed + i*golden (bijection on u64); 1 = 62-bit 2-bit-packed 31-mers (bijection on 2^62);
 * 2 = 2*i (benches/build.rs:11).  out is a device pointer, i runs over [start, start+n). */
BPHF_API int bphf_keygen_u64(uint64_t *out_dev, uint64_t start, uint64_t n, uint64_t seed, int kind, int device,
                             void *stream);
/* the "L2-atomic ceiling" of this device, measured: best-of-3 time of n_ops random 4-byte operations on an array of
 * array_bytes bytes -- kind 0: non-returning OR (REDG), 1: returning fetch_or (ATOMG), 2: L1-bypassing load */
BPHF_API int bphf_l2_probe(int kind, uint64_t array_bytes, uint64_t n_ops, int device, float *ms_out);
/* MPHF property at full size: *n_bad = number of values that are >= n or repeat an earlier one */
BPHF_API int bphf_check_permutation(const uint64_t *vals_dev, uint64_t n, int device, uint64_t *n_bad);
/* order-independent checksum (sum and xor of splitmix64(v)) of a device u64 array */
BPHF_API int bphf_checksum_u64(const uint64_t *vals_dev, uint64_t n, int device, uint64_t *sum_out,
                               uint64_t *xor_out);
/* pinned host memory for end-to-end runs */
BPHF_API void *bphf_host_alloc(size_t bytes);
BPHF_API void bphf_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* BOOMPHF_CUDA_H */
