// boomphf_cuda.cu -- C ABI (include/boomphf_cuda.h) over the sm_100a kernels.
// Host driver of the level cascade: Mphf::new / new_parallel  (/root/reference/src/lib.rs:235-275, 363-408).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <chrono>
#include <vector>

#include "build_kernels.cuh"
#include "bucket_kernels.cuh"
#include "query_kernels.cuh"
#include "qpart_kernels.cuh"
#include "shard_kernels.cuh"
#include "xchg_kernels.cuh"

using namespace bphf;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

struct CudaFailure {
    int code;
    std::string msg;
};

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            char buf__[512];                                                                             \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            throw CudaFailure{e__ == cudaErrorMemoryAllocation ? BPHF_ERR_NOMEM : BPHF_ERR_CUDA, buf__}; \
        }                                                                                                \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        CK(cudaGetDevice(&prev));
        if (prev != dev) CK(cudaSetDevice(dev));
    }
    ~DeviceGuard() {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != prev && prev >= 0) cudaSetDevice(prev);
    }
};

constexpr int MAX_DEVICES = 64;
struct DeviceInfo {
    bool init = false;
    int sm_count = 0;
    int cc_major = 0;
    uint64_t total_mem = 0;
};
static DeviceInfo g_dev[MAX_DEVICES];
static std::mutex g_dev_mutex;
static int g_profile = 0;
static int g_query_mode = 0;  // 0: level-synchronous packed kernel, 1: thread-per-key packed kernel (measurement)

// partitioned lookup (qpart_kernels.cuh): structures larger than the L2
struct QPartCfg {
    int mode = 0;                      // 0 automatic, 1 never, 2 whenever the structure allows it (tests)
    uint32_t slice_shift = 28;         // bits of a level per bucket: 2^28 bits = 45 MB of granules
    uint64_t walk_bytes = 40ull << 20; // the levels left to the walk stage fit this many bytes of granules
    uint64_t batch_keys = 1ull << 27;  // queries per pass (scratch: ~55 B per key and stage)
    double cap_scale = 1.0;            // bucket list capacity relative to the planned one (tests: force overflows)
    // automatic mode: a slice only pays off when it takes many more probes than it has sectors (all of level 0: 9e6
    // sectors at 1e9 keys), so smaller batches go to the plain kernel ...
    uint64_t min_keys = 1ull << 25;
    // ... and so do structures that the L2 still serves well: measured crossover (1e8+ lookups, plain vs partitioned
    // Gkeys/s) 61 / 44 at 115 MB of granules, 48 / 45 at 172 MB, 38 / 44 at 230 MB, 34 / 44 at 287 MB, 27 / 40 at 580 MB
    uint64_t min_struct_bytes = 200000000ull;
};
static QPartCfg g_qpart;
static std::mutex g_qpart_mutex;
__device__ unsigned long long g_qp_fallbacks;  // batches redone by qfallback_kernel on this device
static std::atomic<uint64_t> g_qp_batches{0};
static std::atomic<uint32_t> g_qp_last_stages{0};

static const DeviceInfo& device_info(int dev) {
    if (dev < 0 || dev >= MAX_DEVICES) throw CudaFailure{BPHF_ERR_ARG, "device ordinal out of range"};
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    DeviceInfo& d = g_dev[dev];
    if (!d.init) {
        int count = 0;
        CK(cudaGetDeviceCount(&count));
        if (dev >= count) throw CudaFailure{BPHF_ERR_CUDA, "no such CUDA device (this library has no CPU fallback)"};
        int prev = -1;
        CK(cudaGetDevice(&prev));
        CK(cudaSetDevice(dev));
        CK(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
        CK(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        {
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            d.total_mem = total_b;
        }
        if (d.cc_major != 10)
            throw CudaFailure{BPHF_ERR_CUDA, "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only"};
        // keep freed build scratch in the stream-ordered pool instead of returning it to the OS
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t thresh = UINT64_MAX;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
        CK(cudaFuncSetAttribute(tail_kernel<KeyCfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TailSmem)));
        CK(cudaFuncSetAttribute(tail_kernel<StreamCfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TailSmem)));
        // the level kernels keep 32 KB of TMA ring per CTA; leave the L1 room for their random probes in flight
        // (with the carve-out maxed for occupancy the probes starve: measured 2x slower)
        {
            const void* ring_kernels[] = {(const void*)pass_kernel<false>, (const void*)pass_kernel<true>,
                                          (const void*)pass_kernel<true, StreamCfg>,
                                          (const void*)pass_chunk_kernel<false>, (const void*)pass_chunk_kernel<true>,
                                          (const void*)cascade_kernel<false>, (const void*)cascade_kernel<true>,
                                          (const void*)shard_gather_kernel<false>, (const void*)shard_gather_kernel<true>,
                                          (const void*)shard_gather_x_kernel<false>, (const void*)shard_gather_x_kernel<true>,
                                          (const void*)pass_x_kernel<false>, (const void*)pass_x_kernel<true>,
                                          (const void*)xscatter_kernel<false>, (const void*)xscatter_kernel<true>};
            for (const void* k : ring_kernels) {
                CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XS_SMEM_BYTES));
                CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 60));
            }
        }
        if (const char* e = getenv("BPHF_QP_CARVEOUT"))  // measurement aid: shared-memory carve-out of the probing lookup kernel
            CK(cudaFuncSetAttribute(qprobe_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
        CK(cudaFuncSetAttribute(scatter0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(bucket_mark_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        // CUDA loads kernels lazily, and loading one may have to wait for running work: a shard spinning at a
        // barrier would then block the first launch of the very kernel it waits for.  Load everything now.
        {
            cudaFuncAttributes fa;
            const void* all_kernels[] = {
                (const void*)mark_list_kernel<false>, (const void*)mark_list_kernel<true>, (const void*)pass_kernel<false>,
                (const void*)pass_kernel<true>, (const void*)finalize_kernel, (const void*)tail_kernel<KeyCfg>,
                (const void*)mark_list_kernel<true, StreamCfg>, (const void*)pass_kernel<true, StreamCfg>,
                (const void*)tail_kernel<StreamCfg>, (const void*)query_kernel<false, StreamCfg>,
                (const void*)query_kernel<true, StreamCfg>, (const void*)iota_kernel, (const void*)stream_check_kernel,
                (const void*)pass_chunk_kernel<false>, (const void*)pass_chunk_kernel<true>,
                (const void*)cascade_kernel<false>, (const void*)cascade_kernel<true>,
                (const void*)scan_sums_kernel, (const void*)scan_top_kernel, (const void*)scan_apply_kernel,
                (const void*)scatter0_kernel, (const void*)scatter_kernel, (const void*)bucket_mark_kernel,
                (const void*)query_kernel<false>, (const void*)query_kernel<true>, (const void*)map_permute_kernel<false>,
                (const void*)map_get_kernel<false>, (const void*)map_scatter_kernel, (const void*)map_gather_kernel,
                (const void*)pack_query_kernel, (const void*)set_base_rank_kernel, (const void*)query_sync_kernel,
                (const void*)qfirst_kernel, (const void*)qprobe_kernel, (const void*)qwalk_kernel,
                (const void*)qrestore_kernel<true>, (const void*)qrestore_kernel<false>, (const void*)qfallback_kernel,
                (const void*)key_words_kernel, (const void*)l2_probe_kernel, (const void*)keygen_kernel, (const void*)check_perm_kernel, (const void*)checksum_kernel,
                (const void*)shard_barrier_kernel, (const void*)or_collide_merge_kernel, (const void*)tail_gather_kernel,
                (const void*)shard_gather_kernel<false>, (const void*)shard_gather_kernel<true>,
                (const void*)shard_collect_kernel, (const void*)shard_switch_kernel,
                (const void*)shard_gather_x_kernel<false>, (const void*)shard_gather_x_kernel<true>,
                (const void*)pass_x_kernel<false>, (const void*)pass_x_kernel<true>,
                (const void*)xscatter_kernel<false>, (const void*)xscatter_kernel<true>, (const void*)xmark_kernel,
                (const void*)xfinalize_kernel, (const void*)xpush_kernel, (const void*)xbits_kernel, (const void*)xpop_kernel};
            for (const void* k : all_kernels) CK(cudaFuncGetAttributes(&fa, k));
        }
        if (prev != dev) CK(cudaSetDevice(prev));
        d.init = true;
    }
    return d;
}

// library-owned streams, one set per (thread, device)
struct ThreadStreams {
    cudaStream_t main[MAX_DEVICES] = {};
    cudaStream_t copy[MAX_DEVICES] = {};
    ~ThreadStreams() {
        for (int d = 0; d < MAX_DEVICES; d++) {
            if (main[d]) cudaStreamDestroy(main[d]);
            if (copy[d]) cudaStreamDestroy(copy[d]);
        }
    }
};
static thread_local ThreadStreams g_streams;

// pinned staging for the device -> host read of the build state: a copy into pageable memory is synchronous
// inside the driver, and a shard blocked there can keep another shard's thread of the same process from launching
struct PinnedState {
    void* p = nullptr;
    ~PinnedState() {
        if (p) cudaFreeHost(p);
    }
};
static thread_local PinnedState g_pinned_state;
static void* pinned_state_buffer(size_t bytes) {
    if (!g_pinned_state.p) CK(cudaHostAlloc(&g_pinned_state.p, bytes, cudaHostAllocDefault));
    return g_pinned_state.p;
}
static cudaStream_t own_stream(int dev, bool copy) {
    cudaStream_t& s = copy ? g_streams.copy[dev] : g_streams.main[dev];
    if (!s) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return s;
}

// Memory that a handle keeps (arena words, ranks, lookup tables) is allocated AND freed on one process-wide stream per
// device, whatever stream the build ran on and whichever thread frees the handle.  Measured reason: the stream-
// ordered pool only recycles a block cheaply on the stream that freed it; handles allocated on a caller's stream and
// freed on another (the freeing thread's) made every following build spend 2-90 ms in cudaMallocAsync, and the first
// build after a device-wide synchronize 1-60 ms (tools/host_trace.py).
static cudaStream_t g_handle_stream[MAX_DEVICES] = {};
static std::mutex g_handle_stream_mutex;
static cudaStream_t handle_stream(int dev) {
    std::lock_guard<std::mutex> lock(g_handle_stream_mutex);
    if (!g_handle_stream[dev]) CK(cudaStreamCreateWithFlags(&g_handle_stream[dev], cudaStreamNonBlocking));
    return g_handle_stream[dev];
}
struct ThreadEvent {
    cudaEvent_t e[MAX_DEVICES] = {};
    ~ThreadEvent() {
        for (auto x : e)
            if (x) cudaEventDestroy(x);
    }
};
static thread_local ThreadEvent g_order_event;
// order stream `after` behind everything enqueued on `before` so far
static void order_streams(int dev, cudaStream_t before, cudaStream_t after) {
    if (before == after) return;
    cudaEvent_t& e = g_order_event.e[dev];
    if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventRecord(e, before));
    CK(cudaStreamWaitEvent(after, e, 0));
}
// handle-owned allocation, usable on stream `use` from here on
template <typename T>
static void handle_alloc(T** out, size_t bytes, int dev, cudaStream_t use) {
    const cudaStream_t hs = handle_stream(dev);
    CK(cudaMallocAsync((void**)out, bytes, hs));
    order_streams(dev, hs, use);
}
// give handle-owned memory back once `last_use` has passed everything enqueued on it so far
static void handle_free(void* p, int dev, cudaStream_t last_use) {
    if (!p) return;
    const cudaStream_t hs = handle_stream(dev);
    order_streams(dev, last_use, hs);
    CK(cudaFreeAsync(p, hs));
}

// ------------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------------
struct bphf_mphf {
    int device = 0;
    uint32_t n_levels = 0;
    uint64_t n_keys = 0;       // popcount of all levels == number of keys
    uint64_t words_used = 0;   // arena words (multiple of 8)
    uint64_t* d_words = nullptr;
    uint64_t* d_ranks = nullptr;
    QLevel* d_lv = nullptr;
    QLevelP* d_lvp = nullptr;   // packed lookup structure (query_kernels.cuh); null when a level has >= 2^32 bits
    uint4* d_packed = nullptr;
    uint64_t total_gran = 0;
    KeyCfg kc = make_key_cfg(BPHF_KIND_U64, 8);
    std::vector<LevelDesc> lv;
    bphf_build_stats stats;
    bphf_mphf() { memset(&stats, 0, sizeof stats); }
};

// ------------------------------------------------------------------------------------------------
// sharded build: the peer-visible arena of one shard
// ------------------------------------------------------------------------------------------------
struct bphf_comm {
    int device = 0;
    int rank = 0, world = 1;
    uint64_t phys_words = 0;     // words per arena (incl. the tail reserve)
    unsigned char* base = nullptr;  // [words | coll | xkeys | flags], one cudaMalloc (IPC-exportable)
    size_t bytes = 0;
    void* peer_base[SHARD_MAX] = {};  // mapped bases of all shards (own base at [rank])
    bool peer_is_ipc[SHARD_MAX] = {};
    bool connected = false;
    bool co_located = false;     // another shard shares this GPU (tests): no cooperative launches, they could not
                                 // become resident next to a peer's spinning barrier kernel
    uint64_t epoch = 0;          // barrier counter; advances identically on every shard
    cudaStream_t stream = nullptr;  // the shard's build stream (owned: one hardware queue per live shard)
    cudaStream_t push_stream = nullptr;  // exchange levels: final a slices go to the peers next to the build stream's work
    cudaEvent_t push_ev = nullptr;
    void* pinned = nullptr;         // pinned staging for the build state read-back
    CommDev dev;
    // exchange levels (xchg_kernels.cuh): world x x_regions blocks of x_sub_cap entries each way
    uint32_t x_regions = 0, x_sub_cap = 0;
    size_t x_entries() const { return (size_t)world * x_regions * x_sub_cap; }
    size_t off_coll() const { return (size_t)phys_words * 8; }
    size_t off_xkeys() const { return (size_t)phys_words * 16; }
    size_t off_flags() const { return off_xkeys() + XKEYS_WORDS * 8; }
    size_t off_xidx() const { return off_flags() + (size_t)SHARD_MAX * FLAG_STRIDE * 8; }
    size_t off_xcnt() const { return off_xidx() + x_entries() * 4; }
    size_t off_xbits() const { return (off_xcnt() + (size_t)world * x_regions * 4 + 127) & ~(size_t)127; }
    size_t total_bytes() const { return off_xbits() + x_entries() / 8 + 128; }
    void map_peer(int p, void* b) {
        peer_base[p] = b;
        unsigned char* u = static_cast<unsigned char*>(b);
        dev.words[p] = reinterpret_cast<uint64_t*>(u);
        dev.coll[p] = reinterpret_cast<uint64_t*>(u + off_coll());
        dev.xkeys[p] = reinterpret_cast<uint64_t*>(u + off_xkeys());
        dev.flags[p] = reinterpret_cast<uint64_t*>(u + off_flags());
        dev.xidx[p] = reinterpret_cast<uint32_t*>(u + off_xidx());
        dev.xcnt[p] = reinterpret_cast<uint32_t*>(u + off_xcnt());
        dev.xbits[p] = reinterpret_cast<uint32_t*>(u + off_xbits());
    }
};

static void upload_level_table(bphf_mphf* m, cudaStream_t s) {
    std::vector<QLevel> q(std::max<uint32_t>(m->n_levels, 1));
    for (uint32_t l = 0; l < m->n_levels; l++) {
        q[l].bits = m->lv[l].bits;
        q[l].bit_off = m->lv[l].word_off * 64;
        q[l].sp0 = m->lv[l].sp0_query;
    }
    handle_alloc(&m->d_lv, q.size() * sizeof(QLevel), m->device, s);
    CK(cudaMemcpyAsync(m->d_lv, q.data(), q.size() * sizeof(QLevel), cudaMemcpyHostToDevice, s));
    // packed lookup structure: needs the ranks (already enqueued on s) and 32-bit slot indices
    bool can_pack = m->n_levels > 0;
    for (uint32_t l = 0; l < m->n_levels; l++) can_pack = can_pack && (m->lv[l].bits >> 32) == 0;
    std::vector<QLevelP> qp(std::max<uint32_t>(m->n_levels, 1));
    if (can_pack) {
        uint64_t off = 0;
        for (uint32_t l = 0; l < m->n_levels; l++) {
            qp[l].sp0 = m->lv[l].sp0_query;
            qp[l].base_rank = 0;  // first rank entry of the level: filled in on the device (no host round trip)
            qp[l].gran_off = off;
            qp[l].bits = (uint32_t)m->lv[l].bits;
            qp[l].n_gran = (uint32_t)((m->lv[l].bits + GRAN_BITS - 1) / GRAN_BITS);
            off += qp[l].n_gran;
        }
        m->total_gran = off;
        handle_alloc(&m->d_lvp, qp.size() * sizeof(QLevelP), m->device, s);
        CK(cudaMemcpyAsync(m->d_lvp, qp.data(), qp.size() * sizeof(QLevelP), cudaMemcpyHostToDevice, s));
        handle_alloc(&m->d_packed, std::max<uint64_t>(off, 1) * sizeof(uint4), m->device, s);
        const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((off + 255) / 256, (uint64_t)g_dev[m->device].sm_count * 16));
        set_base_rank_kernel<<<1, 128, 0, s>>>(m->d_lvp, m->d_lv, m->n_levels, m->d_ranks);
        pack_query_kernel<<<grid, 256, 0, s>>>(m->d_lvp, m->d_lv, m->n_levels, m->d_words, m->d_ranks, off, m->d_packed);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(s));  // q / qp are stack vectors
}

static QView query_view(const bphf_mphf* m) {
    QView v;
    v.lv = m->d_lv;
    v.lvp = m->d_lvp;
    v.words = m->d_words;
    v.ranks = m->d_ranks;
    v.packed = m->d_packed;
    v.n_levels = m->n_levels;
    v.kc = m->kc;
    return v;
}

// ------------------------------------------------------------------------------------------------
// plan: simulation of the cascade with upper- and lower-biased key counts (expected colliding
// fraction 1 - exp(-k/bits), +-6 sigma).  It only sizes buffers, grids and shared memory and decides
// which kernel variants get enqueued; the kernels read the real sizes from the device state.
// ------------------------------------------------------------------------------------------------
// Mphf::from_chunked_iterator_parallel semantics (src/lib.rs:539-667); mode 0 = plain new / new_parallel
struct ChunkSpec {
    KeyCfg kc = make_key_cfg(BPHF_KIND_U64, 8);  // key kind of the build
    int mode = 0;
    uint64_t thr = 0;        // (0.01f32 * n as f32) as u64
    int has_max = 0;
    uint64_t max_iters = 0;
    // streamed ingest (ring_keys > 0): the level-0 keys stay in HOST memory as n_segs arrays and pass through a small
    // device ring twice (mark of level 0, then filter of level 0 + mark of level 1); level 1's list is where the key
    // set becomes device resident
    const uint64_t* const* seg_ptr = nullptr;
    const uint64_t* seg_len = nullptr;
    uint64_t n_segs = 0;
    uint64_t ring_keys = 0;
};

struct LevelPlan {
    uint64_t k_up = 0, k_lo = 0;
    bool maybe_bucket = false, surely_bucket = false;
    uint32_t range_max = 0, count_max = 1;  // bucket geometry upper bounds (for shared-memory sizing)
    uint64_t need_keys = 0;                 // capacity needed in buf[level & 1]
};
struct Plan {
    uint64_t words_cap = 0;
    uint64_t list_cap[2] = {0, 0};
    std::vector<LevelPlan> lv;  // non-tail levels
    bool maybe_big = false;
};

// 0: L2-atomic kernels, levels >= 1 in one persistent cascade kernel (default, fastest measured)
// 1: L2-atomic kernels, two launches per level   2: bucketed shared-memory kernels for the large levels
static int g_build_mode = 0;
// sharded builds (mode 0): a level is built by the exchange ("owner computes") kernels when its bit count exceeds this
// (the full-width (a, collide) pair of the merge path would not stay L2 resident); 0 disables the exchange path
static double g_xchg_bits = [] {
    const char* e = getenv("BPHF_XCHG_BITS");  // measurement override (bphf_set_xchg_bits is the API)
    return e ? atof(e) : 4.5e8;
}();
static uint64_t g_bucket_min_keys = 1ull << 20;   // levels below this are built by the plain kernels

// a level is built by the bucketed kernels when its (a, collide) pair clearly exceeds the L2: random L2 atomics fall
// from ~126 G/s to ~30 G/s once they miss to HBM.  Measured crossover (level-0 cost per 1e8 keys, plain vs
// bucketed): 1.7e8 bits 0.97 / 1.05 ms, 3.4e8 bits 1.1-1.3 / 1.1-1.2 ms, 6.8e8 bits 2.7 / 1.5 ms.
static double bucket_fit_bits() {
    static const double v = [] {
        const char* e = getenv("BPHF_FIT_BITS");  // measurement override
        const double x = e ? atof(e) : 0.0;
        return x > 0.0 ? x : 4.5e8;
    }();
    return v;
}

static Plan make_plan(uint64_t n, double gamma, uint32_t slots, bool bucket_enable, uint64_t min_keys, double frac = 1.0,
                      const ChunkSpec* chunk = nullptr) {
    Plan p;
    uint64_t k_up = n, k_lo = n;
    uint64_t words = 0;
    bool switched = false;  // chunked-parallel: the levels after the first small one use gamma 1.7
    for (uint32_t level = 0; level <= BPHF_MAX_LEVELS && k_up > 0; level++) {
        const double g = switched ? 1.7 : gamma;
        if (chunk && chunk->mode && !switched && k_lo < 50000000ULL && k_lo < chunk->thr) switched = true;
        const uint64_t bits_up = level_size(g, k_up), bits_lo = level_size(g, k_lo);
        if (level == 0 && (bits_up >> 32)) p.maybe_big = true;
        if (k_up <= TAIL_MAX_KEYS && bits_up <= TAIL_MAX_BITS) break;
        words += round_up8((bits_up + 63) / 64);
        LevelPlan lp;
        lp.k_up = k_up;
        lp.k_lo = k_lo;
        const bool prev_maybe = level == 0 || p.lv[level - 1].maybe_bucket;
        const bool prev_surely = level == 0 || p.lv[level - 1].surely_bucket;
        BucketGeom gu{}, gl{};
        const bool bu = bucket_enable && bucket_geometry(bits_up, k_up, slots, min_keys, frac, &gu);
        const bool bl = bucket_enable && bucket_geometry(bits_lo, k_lo, slots, min_keys, frac, &gl);
        lp.maybe_bucket = prev_maybe && (bu || bl);
        lp.surely_bucket = prev_surely && bu && bl;
        lp.need_keys = level >= 1 ? k_up : 0;
        if (lp.maybe_bucket && frac < 1.0) lp.need_keys = 0;  // sharded: plain lists are sized by the caller
        if (lp.maybe_bucket) {
            // the real geometry is derived from the real k; size for the worst of both estimates
            if (bu && bl && gu.range == gl.range) {
                lp.range_max = gu.range;
                lp.count_max = std::max(gu.count, gl.count);
            } else {
                lp.range_max = BUCKET_R_MAX;
                lp.count_max = std::max(bu ? gu.count : 1u, bl ? gl.count : 1u) + slots;
                if (lp.count_max > BUCKET_MAX_COUNT) lp.count_max = BUCKET_MAX_COUNT;
            }
            const uint64_t cap_max = (uint64_t)std::max(bu ? gu.cap : 0u, bl ? gl.cap : 0u) + 64;
            lp.need_keys = std::max<uint64_t>(lp.need_keys, (uint64_t)lp.count_max * cap_max);
        }
        p.lv.push_back(lp);
        // Expected number of colliding keys of a level.  hashmod (src/lib.rs:78-88) is a fast-range reduction of a
        // 32-bit hash when bits < 2^32: a slot owns floor(2^32/bits) or ceil(2^32/bits) hash values, so the slots are
        // NOT equally likely and more keys collide than the uniform 1 - exp(-k/bits) predicts (+0.3 % at 6.8e8 bits,
        // far more near 2^32).  Model both slot classes (Poisson), then +-6 sigma.
        auto next = [](uint64_t k, uint64_t bits, double sign) {
            double frac;
            if (bits >> 32) {
                frac = 1.0 - std::exp(-(double)k / (double)bits);
            } else {
                const double two32 = 4294967296.0;
                const double q = two32 / (double)bits, lo = std::floor(q), f = q - lo;
                const double p_lo = lo / two32, p_hi = (lo + 1.0) / two32;
                const double share_hi = f * (double)bits * p_hi, share_lo = (1.0 - f) * (double)bits * p_lo;
                frac = share_hi * (1.0 - std::exp(-(double)k * p_hi)) + share_lo * (1.0 - std::exp(-(double)k * p_lo));
            }
            const double kn = (double)k * frac;
            const double v = kn + sign * (6.0 * std::sqrt(kn) + 64.0 + 1e-4 * kn);
            return v <= 0.0 ? (uint64_t)0 : (uint64_t)v;
        };
        k_up = std::min(k_up, next(k_up, bits_up, +1.0));
        k_lo = std::min(k_lo, next(k_lo, bits_lo, -1.0));
    }
    p.words_cap = words + 1024;
    for (size_t l = 0; l < p.lv.size(); l++) p.list_cap[l & 1] = std::max(p.list_cap[l & 1], p.lv[l].need_keys);
    if (bucket_enable && !p.lv.empty() && p.lv.back().maybe_bucket)  // first tail level right after a bucketed one
        p.list_cap[p.lv.size() & 1] = std::max<uint64_t>(p.list_cap[p.lv.size() & 1], TAIL_MAX_KEYS);
    return p;
}

// ------------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------------
namespace {

struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    int kind = 0;  // 0 entry step of a level (scatter / pass / mark_list), 1 completion step (bucket_mark / finalize), 2 tail, 3 ranks
    uint32_t level = 0;
};

constexpr size_t MAX_DYN_SMEM = 200 * 1024;

// The two key-list buffers of a build (the only large scratch) are parked per thread and device between builds instead
// of going back to the stream-ordered pool: measured (tools/pool_probe.py), the first LARGE cudaMallocAsync after a
// device-wide synchronize takes 2 ms to 0.5 s (small and medium blocks come back in microseconds), i.e. a caller that
// synchronises the device between two builds paid for it in the next one.  A parked buffer is idle: the build that
// used it has returned, and builds end with a synchronisation of their stream.  At most SCRATCH_CACHE_BYTES per buffer.
constexpr uint64_t SCRATCH_CACHE_BYTES = 2ull << 30;
struct ScratchCache {
    uint64_t* buf[MAX_DEVICES][2] = {};
    uint64_t cap[MAX_DEVICES][2] = {};  // in keys
    ~ScratchCache() {
        for (int d = 0; d < MAX_DEVICES; d++)
            for (int i = 0; i < 2; i++)
                if (buf[d][i] && g_handle_stream[d]) cudaFreeAsync(buf[d][i], g_handle_stream[d]);
    }
};
static thread_local ScratchCache g_scratch;

struct Builder {
    int device;
    const DeviceInfo& info;
    cudaStream_t s;
    int profile;
    BuildState hst;
    BuildState* d_state = nullptr;
    uint64_t* d_keys_owned = nullptr;
    const uint64_t* d_keys = nullptr;
    uint64_t* d_words = nullptr;
    uint64_t* d_coll = nullptr;
    uint32_t* d_blockpop = nullptr;
    uint32_t* d_cursors = nullptr;  // [2][MAX_BUCKETS_ALLOC]
    uint32_t* d_regcnt = nullptr;   // [2][n_regions]: keys per region of the striped key lists (pass_body)
    uint32_t* d_xregcnt = nullptr;  // sharded: [2][n_regions][SHARD_MAX] keys per (region, owner) block of an exchange level
    uint64_t* d_buf[2] = {nullptr, nullptr};   // key lists / bucket key regions, level l in d_buf[l & 1]
    uint64_t buf_cap[2] = {0, 0};               // allocated size of d_buf[] in keys (>= the planned list_cap)
    uint64_t phys_words = 0;  // words_cap + TAIL_RESERVE_WORDS
    uint64_t n = 0;
    bool maybe_big = false;
    bool stream_idle = false;  // set once the build stream has been synchronised after the last use of the list buffers
    bool stream_keys = false;  // BPHF_KIND_STREAM: key lists hold key indices, StreamCfg kernel instantiations
    uint32_t launches = 0, syncs = 0;
    bphf_comm* comm = nullptr;  // sharded build: the arenas live in the comm's peer-visible allocation
    bool gathered = false;      // sharded build: the remaining keys were collected on every shard; unsharded from here
    uint32_t xchg_levels = 0;   // sharded build: leading levels built by the exchange kernels (same on every shard)
    uint32_t xchg_launched = 0; // exchange levels enqueued so far
    std::vector<EventPair> events;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;

    Builder(int dev, const DeviceInfo& di, cudaStream_t st, int prof) : device(dev), info(di), s(st), profile(prof) {
        memset(&hst, 0, sizeof hst);
    }
    ~Builder() {
        // scratch goes back to the pool (stream ordered); the handle took ownership of what it keeps
        if (d_state) cudaFreeAsync(d_state, s);
        if (d_keys_owned) cudaFreeAsync(d_keys_owned, s);
        if (d_words && !comm) {
            try {
                handle_free(d_words, device, s);
            } catch (...) {
            }
        }
        if (d_coll && !comm) cudaFreeAsync(d_coll, s);
        if (d_blockpop) cudaFreeAsync(d_blockpop, s);
        if (d_cursors) cudaFreeAsync(d_cursors, s);
        if (d_regcnt) cudaFreeAsync(d_regcnt, s);
        if (d_xregcnt) cudaFreeAsync(d_xregcnt, s);
        for (int i = 0; i < 2; i++) {
            try {
                release_list(i);
            } catch (...) {
            }
        }
        for (auto& e : events) {
            if (e.a) cudaEventDestroy(e.a);
            if (e.b) cudaEventDestroy(e.b);
        }
        if (ev_begin) cudaEventDestroy(ev_begin);
        if (ev_end) cudaEventDestroy(ev_end);
    }

    // persistent grids: exactly the CTAs that are resident at once (a grid of 8 per SM with 6 resident runs a second,
    // two-thirds empty wave: 41 tiles per CTA either way)
    int resident_grid(const void* fn, int threads, size_t dyn_smem = 0) const {
        static std::mutex mu;
        static std::vector<std::pair<const void*, int>> cache;
        std::lock_guard<std::mutex> lock(mu);
        for (auto& e : cache)
            if (e.first == fn) return info.sm_count * e.second;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, dyn_smem) != cudaSuccess || per_sm < 1)
            per_sm = dyn_smem ? 1 : 4;
        if (dyn_smem) per_sm = std::min(per_sm, BPHF_PASS_CTAS);  // level kernels: the rest of the array stays L1
        cache.emplace_back(fn, per_sm);
        return info.sm_count * per_sm;
    }
    uint32_t slots() const { return (uint32_t)info.sm_count * 2; }
    // regions of the striped key lists: one per warp of the persistent cascade grid
    uint32_t n_regions() const { return (uint32_t)info.sm_count * BPHF_PASS_CTAS * WP_WARPS; }
    static uint64_t region_cap(uint64_t dense_keys, uint32_t R) {  // key slots per region for a dense source of that size
        const uint64_t tiles = (dense_keys + WP_TILE - 1) / WP_TILE;
        return std::max<uint64_t>(1, (tiles + R - 1) / R) * WP_TILE;
    }

    size_t begin_event(int kind, uint32_t level) {
        const bool want = profile >= 2 || (profile == 1 && (((kind <= 1 || kind == 4) && level <= 1) || kind == 5));
        if (!want) return (size_t)-1;
        EventPair e;
        e.kind = kind;
        e.level = level;
        CK(cudaEventCreate(&e.a));
        CK(cudaEventCreate(&e.b));
        CK(cudaEventRecord(e.a, s));
        events.push_back(e);
        return events.size() - 1;
    }
    void end_event(size_t id) {
        if (id != (size_t)-1) CK(cudaEventRecord(events[id].b, s));
    }

    void alloc_arena(uint64_t words_cap) {
        phys_words = words_cap + TAIL_RESERVE_WORDS;
        if (comm) {
            if (phys_words > comm->phys_words)
                throw CudaFailure{BPHF_ERR_NOMEM, "sharded build: the comm's arena is too small for this key count / gamma"};
            d_words = comm->dev.words[comm->rank];
            d_coll = comm->dev.coll[comm->rank];
        } else {
            handle_alloc(&d_words, phys_words * 8, device, s);  // the handle keeps the arena
            CK(cudaMallocAsync((void**)&d_coll, phys_words * 8, s));
        }
        CK(cudaMallocAsync((void**)&d_blockpop, (phys_words / 8) * 4, s));
        CK(cudaMallocAsync((void**)&d_cursors, (size_t)2 * MAX_BUCKETS_ALLOC * CURSOR_STRIDE * 4, s));
        CK(cudaMemsetAsync(d_words, 0, phys_words * 8, s));
        CK(cudaMemsetAsync(d_coll, 0, phys_words * 8, s));
        CK(cudaMemsetAsync(d_cursors, 0, (size_t)2 * MAX_BUCKETS_ALLOC * CURSOR_STRIDE * 4, s));
        CK(cudaMallocAsync((void**)&d_regcnt, (size_t)2 * n_regions() * 4, s));
        CK(cudaMemsetAsync(d_regcnt, 0, (size_t)2 * n_regions() * 4, s));
        if (comm) {
            CK(cudaMallocAsync((void**)&d_xregcnt, (size_t)2 * n_regions() * SHARD_MAX * 4, s));
            CK(cudaMemsetAsync(d_xregcnt, 0, (size_t)2 * n_regions() * SHARD_MAX * 4, s));
        }
    }
    // ST_NEED_WORDS: move to a larger arena, keeping everything built so far
    void grow_arena(uint64_t need_words) {
        if (comm) {
            char msg[256];
            snprintf(msg, sizeof msg, "sharded build: level %u (%llu keys) needs %llu arena words, the comm has %llu (duplicate keys?)",
                     hst.n_levels, (unsigned long long)hst.lv[hst.n_levels].k_in, (unsigned long long)need_words,
                     (unsigned long long)hst.words_cap);
            throw CudaFailure{BPHF_ERR_NOMEM, msg};
        }
        const uint64_t new_cap = std::max<uint64_t>(need_words + need_words / 2, hst.words_cap * 2);
        const uint64_t new_phys = new_cap + TAIL_RESERVE_WORDS;
        uint64_t *nw = nullptr, *nc = nullptr;
        uint32_t* nb = nullptr;
        handle_alloc(&nw, new_phys * 8, device, s);
        CK(cudaMallocAsync((void**)&nc, new_phys * 8, s));
        CK(cudaMallocAsync((void**)&nb, (new_phys / 8) * 4, s));
        CK(cudaMemsetAsync(nw, 0, new_phys * 8, s));
        CK(cudaMemsetAsync(nc, 0, new_phys * 8, s));
        const uint64_t used = hst.words_used;
        CK(cudaMemcpyAsync(nw, d_words, used * 8, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(nc, d_coll, used * 8, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(nb, d_blockpop, (used / 8) * 4, cudaMemcpyDeviceToDevice, s));
        handle_free(d_words, device, s);
        CK(cudaFreeAsync(d_coll, s));
        CK(cudaFreeAsync(d_blockpop, s));
        d_words = nw;
        d_coll = nc;
        d_blockpop = nb;
        phys_words = new_phys;
        hst.words_cap = new_cap;
    }
    // ST_NEED_LIST for level `level`: its key list / bucket regions (buf[level & 1]) are too small.  The buffer
    // holds nothing live at that point (level-2's keys are dead), so it is simply replaced.
    void grow_list(int which, uint64_t need) {
        const uint64_t cap = need + need / 8 + 1024;
        if (d_buf[which]) handle_free(d_buf[which], device, s);
        d_buf[which] = nullptr;
        buf_cap[which] = 0;
        acquire_list(which, cap);
        hst.list_cap[which] = cap;
        hst.reg_cap = std::min(hst.list_cap[0], hst.list_cap[1]) / hst.n_regions / WP_TILE * WP_TILE;
    }
    // list buffer `which` with room for `cap` keys: the parked one of this thread when it is large enough
    void acquire_list(int which, uint64_t cap) {
        uint64_t*& parked = g_scratch.buf[device][which];
        uint64_t& parked_cap = g_scratch.cap[device][which];
        if (parked && parked_cap >= cap) {
            d_buf[which] = parked;
            buf_cap[which] = parked_cap;
            parked = nullptr;
            parked_cap = 0;
            return;
        }
        if (parked) {  // too small: give it back for good
            handle_free(parked, device, s);
            parked = nullptr;
            parked_cap = 0;
        }
        handle_alloc(&d_buf[which], cap * 8, device, s);
        buf_cap[which] = cap;
    }
    // end of the build (its stream is synchronised, or the build failed and the buffer is ordered behind the stream)
    void release_list(int which) {
        if (!d_buf[which]) return;
        uint64_t*& parked = g_scratch.buf[device][which];
        if (!parked && buf_cap[which] * 8 <= SCRATCH_CACHE_BYTES && stream_idle) {
            parked = d_buf[which];
            g_scratch.cap[device][which] = buf_cap[which];
        } else {
            handle_free(d_buf[which], device, s);
        }
        d_buf[which] = nullptr;
        buf_cap[which] = 0;
    }

    const uint32_t* list_cursor(uint32_t level) const { return d_cursors + (size_t)(level & 1) * MAX_BUCKETS_ALLOC * CURSOR_STRIDE; }
    uint32_t* a32() const { return reinterpret_cast<uint32_t*>(d_words); }
    uint32_t* c32() const { return reinterpret_cast<uint32_t*>(d_coll); }

    // ---- plain (L2-atomic) kernels ----------------------------------------------------------------
    void launch_mark_list(uint32_t level, uint64_t begin, uint64_t end, uint64_t k_up, const uint64_t* keys_override = nullptr) {
        const uint64_t cnt = level == 0 ? end - begin : k_up;
        if (cnt == 0) return;
        const uint64_t* d_keys = keys_override ? keys_override : this->d_keys;
        const uint64_t tiles = (cnt + PASS_TILE - 1) / PASS_TILE;
        const void* fn = stream_keys ? (const void*)mark_list_kernel<true, StreamCfg>
                         : maybe_big ? (const void*)mark_list_kernel<true> : (const void*)mark_list_kernel<false>;
        const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)resident_grid(fn, PASS_THREADS));
        if (stream_keys)
            mark_list_kernel<true, StreamCfg><<<grid, PASS_THREADS, 0, s>>>(d_state, level, d_keys, begin, end, d_buf[0], d_buf[1], a32(), c32(), list_cursor(level));
        else if (maybe_big)
            mark_list_kernel<true><<<grid, PASS_THREADS, 0, s>>>(d_state, level, d_keys, begin, end, d_buf[0], d_buf[1], a32(), c32(), list_cursor(level));
        else
            mark_list_kernel<false><<<grid, PASS_THREADS, 0, s>>>(d_state, level, d_keys, begin, end, d_buf[0], d_buf[1], a32(), c32(), list_cursor(level));
        CK(cudaGetLastError());
        launches++;
    }
    void launch_pass(uint32_t level, uint64_t k_src_up) {
        const uint64_t tiles = std::max<uint64_t>(1, (k_src_up + WP_TILE - 1) / WP_TILE);
        const void* fn = stream_keys ? (const void*)pass_kernel<true, StreamCfg>
                         : maybe_big ? (const void*)pass_kernel<true> : (const void*)pass_kernel<false>;
        // one warp per region of the striped lists (fewer CTAs when there is less to read)
        const int grid = (int)std::min<uint64_t>((tiles + WP_WARPS - 1) / WP_WARPS, (uint64_t)resident_grid(fn, PASS_THREADS, PASS_SMEM_BYTES));
        if (stream_keys)
            pass_kernel<true, StreamCfg><<<grid, PASS_THREADS, PASS_SMEM_BYTES, s>>>(d_state, level, d_keys, d_buf[0], d_buf[1], a32(), c32());
        else if (maybe_big)
            pass_kernel<true><<<grid, PASS_THREADS, PASS_SMEM_BYTES, s>>>(d_state, level, d_keys, d_buf[0], d_buf[1], a32(), c32());
        else
            pass_kernel<false><<<grid, PASS_THREADS, PASS_SMEM_BYTES, s>>>(d_state, level, d_keys, d_buf[0], d_buf[1], a32(), c32());
        CK(cudaGetLastError());
        launches++;
    }
    // streamed ingest: filter(0) + mark(1) over one chunk of the level-0 keys
    void launch_pass_chunk(const uint64_t* chunk, uint64_t cnt) {
        if (cnt == 0) return;
        const uint64_t tiles = (cnt + WP_TILE - 1) / WP_TILE;
        const void* fn = maybe_big ? (const void*)pass_chunk_kernel<true> : (const void*)pass_chunk_kernel<false>;
        const int grid = (int)std::min<uint64_t>((tiles + WP_WARPS - 1) / WP_WARPS, (uint64_t)resident_grid(fn, PASS_THREADS, PASS_SMEM_BYTES));
        if (maybe_big) pass_chunk_kernel<true><<<grid, PASS_THREADS, PASS_SMEM_BYTES, s>>>(d_state, chunk, cnt, d_buf[0], d_buf[1], a32(), c32());
        else pass_chunk_kernel<false><<<grid, PASS_THREADS, PASS_SMEM_BYTES, s>>>(d_state, chunk, cnt, d_buf[0], d_buf[1], a32(), c32());
        CK(cudaGetLastError());
        launches++;
    }
    void launch_finalize(uint32_t level, uint64_t k_up) {
        const uint64_t bits = level_size(hst.gamma, k_up);
        const uint64_t pairs = round_up8((bits + 63) / 64) / 2;
        const int grid = (int)std::min<uint64_t>(std::max<uint64_t>(1, (pairs + FIN_THREADS - 1) / FIN_THREADS),
                                                 (uint64_t)resident_grid((const void*)finalize_kernel, FIN_THREADS));
        finalize_kernel<<<grid, FIN_THREADS, 0, s>>>(d_state, level, d_words, d_coll, d_blockpop);
        CK(cudaGetLastError());
        launches++;
    }
    // levels >= 1 of an unsharded L2-atomic build: one cooperative launch, every CTA resident
    bool use_cascade() const {
        static const bool no_coop = getenv("BPHF_NO_COOP") != nullptr;  // debugging aid
        if (no_coop || stream_keys) return false;  // stream keys: per-level kernels (their own instantiations)
        return comm ? (gathered && !comm->co_located) : g_build_mode == 0;
    }
    bool sharded_now() const { return comm && !gathered; }
    // sharded: collect the remaining keys on every shard, then mark + finalize level `level` from the collected list
    void launch_gather(uint32_t level, uint64_t k_src_up, uint64_t k_up) {
        const size_t e0 = begin_event(0, level);
        const uint64_t tiles = std::max<uint64_t>(1, (k_src_up + WP_TILE - 1) / WP_TILE);
        const void* gfn = maybe_big ? (const void*)shard_gather_kernel<true> : (const void*)shard_gather_kernel<false>;
        const int grid = (int)std::min<uint64_t>((tiles + WP_WARPS - 1) / WP_WARPS, (uint64_t)resident_grid(gfn, PASS_THREADS, PASS_SMEM_BYTES));
        uint64_t* xk = comm->dev.xkeys[comm->rank];
        if (maybe_big) shard_gather_kernel<true><<<grid, PASS_THREADS, PASS_SMEM_BYTES, s>>>(d_state, level, d_keys, d_buf[0], d_buf[1], c32(), xk);
        else shard_gather_kernel<false><<<grid, PASS_THREADS, PASS_SMEM_BYTES, s>>>(d_state, level, d_keys, d_buf[0], d_buf[1], c32(), xk);
        CK(cudaGetLastError());
        launches++;
        if (level == xchg_levels && xchg_levels > 0) {  // the collect follows an exchange level: filter by the owners' bits
            const void* xfn = maybe_big ? (const void*)shard_gather_x_kernel<true> : (const void*)shard_gather_x_kernel<false>;
            const int xgrid = resident_grid(xfn, PASS_THREADS, PASS_SMEM_BYTES);
            const uint32_t* xb = comm->dev.xbits[comm->rank];
            if (maybe_big) shard_gather_x_kernel<true><<<xgrid, PASS_THREADS, PASS_SMEM_BYTES, s>>>(d_state, level, d_buf[0], d_buf[1], xk, xb);
            else shard_gather_x_kernel<false><<<xgrid, PASS_THREADS, PASS_SMEM_BYTES, s>>>(d_state, level, d_buf[0], d_buf[1], xk, xb);
            CK(cudaGetLastError());
            launches++;
        }
        end_event(e0);
        const size_t em = begin_event(4, level);
        launch_barrier(BAR_LIST_SUM, level);
        const int cgrid = (int)std::min<uint64_t>(std::max<uint64_t>(1, (k_up + 255) / 256), (uint64_t)info.sm_count * 4);
        shard_collect_kernel<<<cgrid, 256, 0, s>>>(d_state, level, comm->dev, d_buf[0], d_buf[1]);
        shard_switch_kernel<<<1, 1, 0, s>>>(d_state, level);
        CK(cudaGetLastError());
        launches += 2;
        end_event(em);
        gathered = true;
        launch_mark_list(level, 0, 0, k_up);
        launch_finalize(level, k_up);
    }
    void launch_cascade() {
        static std::atomic<int> per_sm[2];  // CTAs per SM that a cooperative launch accepts (0: not probed yet)
        const int v = maybe_big ? 1 : 0;
        const void* fn = maybe_big ? (const void*)cascade_kernel<true> : (const void*)cascade_kernel<false>;
        int want = per_sm[v].load();
        if (!want) {
            int n = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, PASS_THREADS, PASS_SMEM_BYTES));
            want = std::max(1, std::min(n, BPHF_PASS_CTAS));  // 64 KB of warp rings per CTA; the rest of the array stays L1
        }
        const size_t ev = begin_event(5, 0);
        void* args[] = {(void*)&d_state, (void*)&d_keys, (void*)&d_buf[0], (void*)&d_buf[1], (void*)&d_words, (void*)&d_coll, (void*)&d_blockpop};
        for (;;) {
            const cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(info.sm_count * want), dim3(PASS_THREADS), args, PASS_SMEM_BYTES, s);
            if (e == cudaSuccess) break;
            // the occupancy query and the launch can disagree (shared-memory carve-out): every CTA must be resident
            if (e == cudaErrorCooperativeLaunchTooLarge && want > 1) {
                cudaGetLastError();
                want--;
                continue;
            }
            CK(e);
        }
        per_sm[v].store(want);
        end_event(ev);
        launches++;
    }
    void launch_tail() {
        const size_t ev = begin_event(2, 0);
        CommDev cd;
        memset(&cd, 0, sizeof cd);
        if (comm) {
            cd = comm->dev;
            tail_gather_kernel<<<1, TAIL_THREADS, 0, s>>>(d_state, d_keys, d_buf[0], d_buf[1], c32(), cd.xkeys[cd.rank], d_cursors);
            CK(cudaGetLastError());
            launches++;
            launch_barrier(BAR_PLAIN, 0);
        }
        if (stream_keys)
            tail_kernel<StreamCfg><<<1, TAIL_THREADS, sizeof(TailSmem), s>>>(d_state, d_keys, d_buf[0], d_buf[1], a32(), c32(), d_blockpop, cd);
        else
            tail_kernel<KeyCfg><<<1, TAIL_THREADS, sizeof(TailSmem), s>>>(d_state, d_keys, d_buf[0], d_buf[1], a32(), c32(), d_blockpop, cd);
        CK(cudaGetLastError());
        end_event(ev);
        launches++;
        // nobody may start the next build (and overwrite its exchange slot) before every shard has read it; the final a
        // slices of the exchange levels that this shard pushed on the side stream are complete before it signals
        if (comm) {
            join_pushes();
            launch_barrier(BAR_PLAIN, 0);
        }
    }
    // ---- sharded build: device-side barrier and OR-collide merge over peer memory ---------------------
    void launch_barrier(uint32_t kind, uint32_t level) {
        comm->epoch += 1;
        shard_barrier_kernel<<<1, 32, 0, s>>>(d_state, comm->dev, comm->epoch, kind, level);
        CK(cudaGetLastError());
        launches++;
    }
    void launch_merge(uint32_t level, uint64_t k_up) {
        const uint64_t bits = level_size(hst.gamma, k_up);
        const uint64_t pairs = round_up8((bits + 63) / 64) / 2;
        const uint64_t per = (pairs + comm->world - 1) / comm->world;
        const int grid = (int)std::min<uint64_t>(std::max<uint64_t>(1, (per + MERGE_THREADS - 1) / MERGE_THREADS),
                                                 (uint64_t)info.sm_count * 4);
        or_collide_merge_kernel<<<grid, MERGE_THREADS, 0, s>>>(d_state, level, comm->dev);
        CK(cudaGetLastError());
        launches++;
    }
    // ---- exchange levels (owner computes, xchg_kernels.cuh) ------------------------------------------
    void launch_xchg_level(uint32_t level, uint64_t k_up) {
        const size_t e0 = begin_event(0, level);
        {
            const void* fn = maybe_big ? (const void*)xscatter_kernel<true> : (const void*)xscatter_kernel<false>;
            const int grid = resident_grid(fn, PASS_THREADS, XS_SMEM_BYTES);
            if (maybe_big) xscatter_kernel<true><<<grid, PASS_THREADS, XS_SMEM_BYTES, s>>>(d_state, level, d_keys, d_buf[0], d_buf[1], comm->dev);
            else xscatter_kernel<false><<<grid, PASS_THREADS, XS_SMEM_BYTES, s>>>(d_state, level, d_keys, d_buf[0], d_buf[1], comm->dev);
            CK(cudaGetLastError());
            launches++;
        }
        end_event(e0);
        const size_t em = begin_event(4, level);
        launch_barrier(BAR_PLAIN, level);
        end_event(em);
        const size_t e1 = begin_event(1, level);
        const int wgrid = info.sm_count * 6;
        xmark_kernel<<<wgrid, PASS_THREADS, 0, s>>>(d_state, level, comm->dev, a32(), c32());
        const uint64_t bits = level_size(hst.gamma, k_up);
        const uint64_t pairs = round_up8((bits + 63) / 64) / 2;
        const uint64_t per = (pairs + comm->world - 1) / comm->world;
        const int fgrid = (int)std::min<uint64_t>(std::max<uint64_t>(1, (per + FIN_THREADS - 1) / FIN_THREADS), (uint64_t)info.sm_count * 4);
        xfinalize_kernel<<<fgrid, FIN_THREADS, 0, s>>>(d_state, level, comm->dev);
        // the final slice goes to the peers on the side stream, next to xbits and everything that follows
        // (shards that share one GPU -- tests -- keep to their one stream: with more streams than hardware queues a
        // peer's spinning barrier kernel can sit in front of the very kernel it waits for)
        cudaStream_t ps = comm->co_located ? s : comm->push_stream;
        if (ps != s) {
            CK(cudaEventRecord(comm->push_ev, s));
            CK(cudaStreamWaitEvent(ps, comm->push_ev, 0));
        }
        xpush_kernel<<<std::max(1, fgrid / 2), FIN_THREADS, 0, ps>>>(d_state, level, comm->dev);
        xbits_kernel<<<wgrid, PASS_THREADS, 0, s>>>(d_state, level, comm->dev, c32());
        CK(cudaGetLastError());
        launches += 4;
        end_event(e1);
        const size_t em2 = begin_event(4, level);
        launch_barrier(BAR_XDONE, level);
        end_event(em2);
        xchg_launched = std::max(xchg_launched, level + 1);
    }
    // every shard's pushes have to be complete before anybody reads a replica: this shard's join the build stream in front
    // of the barrier that ends the build; the rank popcounts of the exchange levels follow that barrier
    void join_pushes() {
        if (!comm || !xchg_launched || comm->co_located) return;
        CK(cudaEventRecord(comm->push_ev, comm->push_stream));
        CK(cudaStreamWaitEvent(s, comm->push_ev, 0));
    }
    void launch_xpop() {
        for (uint32_t level = 0; level < xchg_launched; level++) {
            const int pgrid = info.sm_count * 8;
            xpop_kernel<<<pgrid, FIN_THREADS, 0, s>>>(d_state, level, d_words, d_blockpop);
            CK(cudaGetLastError());
            launches++;
        }
    }
    // first level after the exchange levels: filter by the owners' survivor bits, then as any merge-mode level
    void launch_pass_x(uint32_t level) {
        const void* fn = maybe_big ? (const void*)pass_x_kernel<true> : (const void*)pass_x_kernel<false>;
        const int grid = resident_grid(fn, PASS_THREADS, PASS_SMEM_BYTES);
        const uint32_t* xb = comm->dev.xbits[comm->rank];
        if (maybe_big) pass_x_kernel<true><<<grid, PASS_THREADS, PASS_SMEM_BYTES, s>>>(d_state, level, d_buf[0], d_buf[1], a32(), c32(), xb);
        else pass_x_kernel<false><<<grid, PASS_THREADS, PASS_SMEM_BYTES, s>>>(d_state, level, d_buf[0], d_buf[1], a32(), c32(), xb);
        CK(cudaGetLastError());
        launches++;
    }
    // ---- bucketed (shared-memory) kernels -----------------------------------------------------------
    void launch_scatter0(uint64_t begin, uint64_t end, uint32_t count_max) {
        if (end <= begin) return;
        const size_t smem = std::min(MAX_DYN_SMEM, scatter_smem_bytes(count_max));
        const uint64_t tiles = (end - begin + BK_TILE - 1) / BK_TILE;
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (smem + 1024)));
        const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)info.sm_count * per_sm);
        scatter0_kernel<<<grid, BK_THREADS, smem, s>>>(d_state, d_keys, begin, end, d_buf[0], d_cursors, (uint32_t)smem);
        CK(cudaGetLastError());
        launches++;
    }
    void launch_scatter(uint32_t level, uint32_t src_range_max, uint32_t src_count_max, uint32_t dst_count_max) {
        const size_t smem = std::min(MAX_DYN_SMEM, (size_t)src_range_max / 8 + scatter_smem_bytes(dst_count_max));
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (smem + 1024)));
        const int grid = (int)std::max<uint32_t>(1, std::min<uint32_t>(src_count_max, (uint32_t)(info.sm_count * per_sm)));
        scatter_kernel<<<grid, BK_THREADS, smem, s>>>(d_state, level, d_buf[0], d_buf[1], c32(), d_cursors, (uint32_t)smem);
        CK(cudaGetLastError());
        launches++;
    }
    void launch_bucket_mark(uint32_t level, uint32_t range_max, uint32_t count_max) {
        const size_t smem = std::min(MAX_DYN_SMEM, (size_t)range_max / 4);
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / (smem + 1024)));
        const int grid = (int)std::max<uint32_t>(1, std::min<uint32_t>(count_max, (uint32_t)(info.sm_count * per_sm)));
        bucket_mark_kernel<<<grid, BK_THREADS, smem, s>>>(d_state, level, d_buf[0], d_buf[1], a32(), c32(), d_blockpop,
                                                         d_cursors, (uint32_t)smem);
        CK(cudaGetLastError());
        launches++;
    }

    // Everything level `level` might need, in dependency order; each kernel checks the device state and returns
    // at once when it is not the variant that applies.  level 0's entry step is issued by the caller (chunks).
    void launch_level(uint32_t level, const LevelPlan* prev, const LevelPlan& cur, bool entry_done = false) {
        const bool prev_maybe_b = level >= 1 && prev->maybe_bucket, prev_surely_b = level >= 1 && prev->surely_bucket;
        if (level >= 1 && !entry_done) {
            const size_t e0 = begin_event(0, level);
            if (prev_maybe_b) launch_scatter(level, prev->range_max, prev->count_max, cur.maybe_bucket ? cur.count_max : 1);
            if (!cur.surely_bucket && prev_maybe_b) launch_mark_list(level, 0, 0, cur.k_up);
            if (!prev_surely_b) launch_pass(level, prev->k_up);
            if (comm && !gathered && level == xchg_levels && xchg_levels > 0) launch_pass_x(level);
            end_event(e0);
        }
        const size_t e1 = begin_event(1, level);
        if (cur.maybe_bucket) launch_bucket_mark(level, cur.range_max, cur.count_max);
        end_event(e1);
        if (sharded_now()) {
            // every shard has marked its keys -> fold the (a, collide) pairs of all shards -> everybody holds
            // the merged level.  The first barrier also checks that the shards' lists add up to k(level).
            const size_t em = begin_event(4, level);
            launch_barrier(BAR_LIST_SUM, level);
            launch_merge(level, cur.k_up);
            launch_barrier(BAR_PLAIN, level);
            end_event(em);
        }
        const size_t e2 = begin_event(1, level);
        if (!cur.surely_bucket || sharded_now()) launch_finalize(level, cur.k_up);
        end_event(e2);
    }

    void upload_state() { CK(cudaMemcpyAsync(d_state, &hst, sizeof hst, cudaMemcpyHostToDevice, s)); }
    void download_state() {
        void* pin = comm ? comm->pinned : pinned_state_buffer(sizeof hst);
        CK(cudaMemcpyAsync(pin, d_state, sizeof hst, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        memcpy(&hst, pin, sizeof hst);
        syncs++;
    }
};

}  // namespace

enum { BUILD_RETRY_PLAIN = 1000, BUILD_RETRY_INCORE = 1001 };

// BPHF_HOST_TRACE=1: host-side timeline of every build on stderr (microseconds since the call started)
struct HostTrace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    char line[512];
    size_t len = 0;
    HostTrace() : t0(std::chrono::steady_clock::now()) {
        static const bool enabled = getenv("BPHF_HOST_TRACE") != nullptr;
        on = enabled;
        line[0] = 0;
    }
    void mark(const char* what) {
        if (!on) return;
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        len += (size_t)snprintf(line + len, len < sizeof line ? sizeof line - len : 0, " %s=%.0f", what, us);
        if (len >= sizeof line) len = sizeof line - 1;
    }
    ~HostTrace() {
        if (on) fprintf(stderr, "[bphf host trace, us]%s\n", line);
    }
};

// device ring of a streamed ingest; released on every exit path
struct IngestRing {
    static constexpr int SLOTS = 3;
    cudaStream_t s;
    uint64_t* buf[SLOTS] = {};
    cudaEvent_t free_ev[SLOTS] = {};
    bool used[SLOTS] = {};
    explicit IngestRing(cudaStream_t st) : s(st) {}
    ~IngestRing() {
        for (int i = 0; i < SLOTS; i++) {
            if (buf[i]) cudaFreeAsync(buf[i], s);
            if (free_ev[i]) cudaEventDestroy(free_ev[i]);
        }
    }
};

// n = keys of this shard; comm == nullptr: unsharded (n_global == n)
static int build_once(const uint64_t* keys, uint64_t n, double gamma, int has_seed, uint64_t seed, int device,
                      int keys_mem, void* stream, bool bucket_enable, uint64_t min_keys, uint32_t rebuilds, bphf_mphf** out,
                      bphf_comm* comm = nullptr, uint64_t n_global_arg = 0, const ChunkSpec* chunk = nullptr) {
    HostTrace trace;
    const DeviceInfo& info = device_info(device);
    DeviceGuard guard(device);
    cudaStream_t s = stream ? (cudaStream_t)stream : own_stream(device, false);
    Builder b(device, info, s, g_profile);
    b.n = n;
    b.comm = comm;
    b.stream_keys = chunk && chunk->kc.kind == BPHF_KIND_STREAM;
    if (b.stream_keys && (comm || bucket_enable)) return fail(BPHF_ERR_INTERNAL, "stream keys take the plain unsharded path");
    const bool ingest = chunk && chunk->ring_keys > 0;
    if (ingest && (comm || bucket_enable || b.stream_keys || chunk->kc.kind != BPHF_KIND_U64))
        return fail(BPHF_ERR_INTERNAL, "streamed ingest takes the plain unsharded u64 path");
    const uint64_t n_global = comm ? n_global_arg : n;
    const double shard_frac = (comm && n_global) ? (double)n / (double)n_global : 1.0;

    CK(cudaEventCreate(&b.ev_begin));
    CK(cudaEventCreate(&b.ev_end));
    CK(cudaEventRecord(b.ev_begin, s));
    const bool want_peak = (chunk && chunk->ring_keys > 0) || g_profile >= 2;
    if (want_peak) {
        // statistics: high-water mark of the stream-ordered pool over this build
        cudaMemPool_t pool;
        uint64_t zero = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &zero);
    }

    trace.mark("events");
    Plan plan = make_plan(n_global, gamma, b.slots(), bucket_enable, min_keys, shard_frac, chunk);
    trace.mark("plan");
    b.maybe_big = plan.maybe_big;
    uint32_t gather_level = NO_LEVEL;
    if (comm) {
        // this shard's redo lists: its expected share of every level + 8 sigma (there is no host round trip to
        // grow a list in the middle of a collective build; overflow ends in BPHF_ERR_NOMEM)
        plan.list_cap[0] = plan.list_cap[1] = 1;
        const double frac = shard_frac;
        for (size_t l = 0; l <= plan.lv.size(); l++) {
            uint64_t cap = 0;
            if (l >= 1) {
                const double est = (l < plan.lv.size() ? (double)plan.lv[l].k_up : (double)TAIL_MAX_KEYS / 0.4) * frac * 1.05;
                cap = std::min<uint64_t>(n, (uint64_t)(est + 8.0 * std::sqrt(est) + 4096.0));
            }
            // bucket regions of a bucketed level (already sized for this shard's share of the keys)
            if (l < plan.lv.size() && plan.lv[l].maybe_bucket) cap = std::max<uint64_t>(cap, plan.lv[l].need_keys);
            plan.list_cap[l & 1] = std::max<uint64_t>(plan.list_cap[l & 1], cap);
        }
        // collect level: first non-tail level >= 1 with at most GATHER_MAX_KEYS keys and plain neighbours
        static const bool no_gather = getenv("BPHF_NO_GATHER") != nullptr;  // measurement aid
        for (size_t l = 1; l < plan.lv.size() && !no_gather; l++) {
            if (plan.lv[l].k_up <= GATHER_MAX_KEYS && !plan.lv[l - 1].maybe_bucket && !plan.lv[l].maybe_bucket) {
                gather_level = (uint32_t)l;
                break;
            }
        }
        if (gather_level != NO_LEVEL)
            for (int i = 0; i < 2; i++) plan.list_cap[i] = std::max<uint64_t>(plan.list_cap[i], GATHER_MAX_KEYS + 1024);
    }

    if (comm && g_build_mode == 0 && g_xchg_bits > 0.0 && !bucket_enable && !b.stream_keys) {
        // exchange levels: a prefix of levels that surely exceed the threshold (both plan estimates), whose slices can
        // be addressed with 32-bit offsets, and that leave two ordinary levels in front of the tail
        if (comm->x_regions != b.n_regions()) return fail(BPHF_ERR_INTERNAL, "comm and builder disagree on the region count");
        uint32_t xl = 0;
        for (size_t l = 0; l + 2 < plan.lv.size(); l++) {
            const uint64_t bits_lo = level_size(gamma, plan.lv[l].k_lo), bits_up = level_size(gamma, plan.lv[l].k_up);
            const uint64_t slice_words = round_up8(((bits_up + 63) / 64 + comm->world - 1) / comm->world);
            if ((double)bits_lo > g_xchg_bits && slice_words * 64 < (1ull << 32)) xl = (uint32_t)l + 1;
            else break;
        }
        b.xchg_levels = xl;
        if (xl > 0) {
            // (region, owner) key blocks of the exchange levels live in the list buffers
            const uint64_t need = (uint64_t)b.n_regions() * comm->world * comm->x_sub_cap;
            for (int i = 0; i < 2; i++) plan.list_cap[i] = std::max<uint64_t>(plan.list_cap[i], need);
            // the collect level may directly follow the exchange levels, but not be one of them
            if (gather_level != NO_LEVEL && gather_level < xl) gather_level = NO_LEVEL;
        }
    }
    if (ingest && plan.lv.size() < 2) return BUILD_RETRY_INCORE;  // level 1 may already belong to the tail kernel
    // ---- striped key lists (pass_body): every region must be able to hold what its share of the first dense source
    // holds -- the input, or the list the last bucketed level leaves behind; later levels only shrink
    const uint32_t n_regions = b.n_regions();
    uint64_t reg_cap = 0;
    if (!plan.lv.empty()) {
        if (ingest) {
            // the chunks of the input pass through pass_chunk_kernel one by one and accumulate in the regions
            for (uint64_t g = 0; g < chunk->n_segs; g++)
                for (uint64_t off = 0; off < chunk->seg_len[g]; off += chunk->ring_keys)
                    reg_cap += Builder::region_cap(std::min(chunk->ring_keys, chunk->seg_len[g] - off), n_regions);
        } else {
            const bool l0_bucketed = plan.lv[0].maybe_bucket;
            const uint64_t dense = comm ? n : (l0_bucketed && plan.lv.size() > 1 ? plan.lv[1].k_up : n_global);
            reg_cap = Builder::region_cap(dense, n_regions);
        }
        for (int i = 0; i < 2; i++) plan.list_cap[i] = std::max<uint64_t>(plan.list_cap[i], (uint64_t)n_regions * reg_cap);
    }
    CK(cudaMallocAsync((void**)&b.d_state, sizeof(BuildState), s));
    trace.mark("a_state");
    b.alloc_arena(plan.words_cap);
    trace.mark("a_arena");
    for (int i = 0; i < 2; i++) {
        if (plan.list_cap[i]) b.acquire_list(i, plan.list_cap[i]);
    }

    trace.mark("alloc");
    BuildState& st = b.hst;
    st.kc = chunk ? chunk->kc : make_key_cfg(BPHF_KIND_U64, 8);
    st.status = ST_RUNNING;
    st.tail_level = NO_TAIL;
    st.words_cap = plan.words_cap;
    st.list_cap[0] = plan.list_cap[0];
    st.list_cap[1] = plan.list_cap[1];
    st.seed0 = has_seed ? seed : 0;  // starting_seed.unwrap_or(0)
    st.gamma = gamma;
    st.bucket_enable = bucket_enable ? 1 : 0;
    st.bucket_slots = b.slots();
    st.bucket_min_keys = min_keys;
    st.shard_rank = comm ? (uint32_t)comm->rank : 0;
    st.shard_world = comm ? (uint32_t)comm->world : 1;
    st.shard_frac = shard_frac;
    st.switch_level = NO_LEVEL;
    if (chunk && chunk->mode) {
        st.chunk_mode = 1;
        st.chunk_thr = chunk->thr;
        st.chunk_has_max = (uint32_t)chunk->has_max;
        st.chunk_max_iters = chunk->max_iters;
    }
    st.gather_level = gather_level;
    st.list_ready_level = NO_LEVEL;
    st.n_regions = n_regions;
    st.reg_cap = reg_cap;
    st.reg_cnt = b.d_regcnt;
    st.xchg_levels = b.xchg_levels;
    st.x_sub_cap = comm ? comm->x_sub_cap : 0;
    st.xreg_cnt = b.d_xregcnt;
    if (!try_make_level(&st, 0, n_global)) return fail(BPHF_ERR_INTERNAL, "level 0 does not fit the planned buffers");
    st.lv[0].k_loc = n;
    b.upload_state();
    trace.mark("prologue");
    const bool level0_bucketed = st.lv[0].b_range != 0;
    const uint32_t count0 = st.lv[0].b_count;

    // ---- level 0 entry step: keys (H2D in chunks when they live on the host, each chunk is consumed as it lands)
    auto level0_entry = [&](uint64_t lo, uint64_t hi) {
        if (level0_bucketed) b.launch_scatter0(lo, hi, count0);
        else b.launch_mark_list(0, lo, hi, 0);
    };
    // ST_NEED_WORDS / ST_NEED_LIST for level `level`: grow what is missing and redo the level set-up on the host until
    // it fits, then hand the state back to the device (words_used still points at the end of the previous level)
    auto grow_until_fits = [&](uint32_t level) {
        const uint64_t k = st.lv[level].k_in;
        for (;;) {
            if (st.status == ST_NEED_WORDS) b.grow_arena(st.need);
            else b.grow_list(level & 1, st.need);
            st.status = ST_RUNNING;
            if (try_make_level(&st, level, k)) break;
        }
        b.upload_state();
    };
    std::vector<LevelPlan> lps = plan.lv;
    uint32_t first_level = 1;  // first level the common launch sequence below still has to run
    IngestRing ring(s);
    if (ingest) {
        // ---- streamed ingest: two passes of the host keys through a 3-slot device ring ------------------------
        cudaStream_t cs = own_stream(device, true);
        for (int i = 0; i < IngestRing::SLOTS; i++) {
            CK(cudaMallocAsync((void**)&ring.buf[i], chunk->ring_keys * 8, s));
            CK(cudaEventCreateWithFlags(&ring.free_ev[i], cudaEventDisableTiming));
        }
        cudaEvent_t ready;
        CK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        CK(cudaEventRecord(ready, s));  // allocations + state upload precede the copies
        CK(cudaStreamWaitEvent(cs, ready, 0));
        cudaEventDestroy(ready);
        uint64_t slot = 0;
        auto pass_over_host_keys = [&](auto&& consume) {
            for (uint64_t g = 0; g < chunk->n_segs; g++) {
                for (uint64_t off = 0; off < chunk->seg_len[g]; off += chunk->ring_keys, slot++) {
                    const uint64_t cnt = std::min(chunk->ring_keys, chunk->seg_len[g] - off);
                    const int i = (int)(slot % IngestRing::SLOTS);
                    if (ring.used[i]) CK(cudaStreamWaitEvent(cs, ring.free_ev[i], 0));  // its last consumer is done
                    CK(cudaMemcpyAsync(ring.buf[i], chunk->seg_ptr[g] + off, cnt * 8, cudaMemcpyHostToDevice, cs));
                    cudaEvent_t landed;
                    CK(cudaEventCreateWithFlags(&landed, cudaEventDisableTiming));
                    CK(cudaEventRecord(landed, cs));
                    CK(cudaStreamWaitEvent(s, landed, 0));
                    cudaEventDestroy(landed);
                    consume(ring.buf[i], cnt);
                    CK(cudaEventRecord(ring.free_ev[i], s));
                    ring.used[i] = true;
                }
            }
        };
        const size_t pe = b.begin_event(0, 0);
        pass_over_host_keys([&](const uint64_t* p, uint64_t cnt) { b.launch_mark_list(0, 0, cnt, 0, p); });
        b.end_event(pe);
        LevelPlan lp0;
        lp0.k_up = lp0.k_lo = n_global;
        lps[0] = lp0;
        b.launch_level(0, nullptr, lps[0]);
        b.download_state();  // one extra host round trip, nothing next to a second pass over PCIe
        if (st.status == ST_NEED_WORDS || st.status == ST_NEED_LIST) grow_until_fits(st.n_levels);
        if (st.status == ST_RUNNING) {
            if (st.tail_level <= 1) return BUILD_RETRY_INCORE;  // the tail kernel would want the level-0 keys resident
            const size_t p1 = b.begin_event(0, 1);
            pass_over_host_keys([&](const uint64_t* p, uint64_t cnt) { b.launch_pass_chunk(p, cnt); });
            b.end_event(p1);
            LevelPlan l1;
            l1.k_up = l1.k_lo = st.lv[1].k_in;
            lps[1] = l1;
            b.launch_level(1, &lps[0], lps[1], true);
            first_level = 2;
        } else {
            first_level = (uint32_t)lps.size();  // done or failed at level 0: the loop below reports it
        }
    } else if (keys_mem == BPHF_MEM_HOST && n && b.xchg_levels > 0) {
        // exchange level 0 routes the shard's whole key set in one pass
        CK(cudaMallocAsync((void**)&b.d_keys_owned, n * 8, s));
        b.d_keys = b.d_keys_owned;
        CK(cudaMemcpyAsync(b.d_keys_owned, keys, n * 8, cudaMemcpyHostToDevice, s));
    } else if (keys_mem == BPHF_MEM_HOST && n) {
        CK(cudaMallocAsync((void**)&b.d_keys_owned, n * 8, s));
        b.d_keys = b.d_keys_owned;
        cudaStream_t cs = own_stream(device, true);
        const uint64_t chunk = 16ull << 20;  // 16 Mi keys = 128 MiB per copy
        const uint64_t n_chunks = (n + chunk - 1) / chunk;
        std::vector<cudaEvent_t> evs(n_chunks);
        cudaEvent_t ready;
        CK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        CK(cudaEventRecord(ready, s));  // allocation + state upload precede the copies
        CK(cudaStreamWaitEvent(cs, ready, 0));
        const size_t pe = b.begin_event(0, 0);
        for (uint64_t c = 0; c < n_chunks; c++) {
            const uint64_t lo = c * chunk, hi = std::min(n, lo + chunk);
            CK(cudaMemcpyAsync(b.d_keys_owned + lo, keys + lo, (hi - lo) * 8, cudaMemcpyHostToDevice, cs));
            CK(cudaEventCreateWithFlags(&evs[c], cudaEventDisableTiming));
            CK(cudaEventRecord(evs[c], cs));
            CK(cudaStreamWaitEvent(s, evs[c], 0));
            level0_entry(lo, hi);
        }
        b.end_event(pe);
        for (auto e : evs) cudaEventDestroy(e);
        cudaEventDestroy(ready);
    } else {
        b.d_keys = keys;
        if (b.xchg_levels == 0) {
            const size_t pe = b.begin_event(0, 0);
            level0_entry(0, n);
            b.end_event(pe);
        }
    }

    // ---- all planned levels, then the tail, without a host round trip
    LevelPlan lp0;
    lp0.k_up = lp0.k_lo = n_global;
    lp0.maybe_bucket = lp0.surely_bucket = level0_bucketed;
    lp0.range_max = st.lv[0].b_range;
    lp0.count_max = level0_bucketed ? count0 : 1;
    if (lps.empty()) lps.push_back(lp0);
    else if (!ingest) lps[0] = lp0;
    if (b.xchg_levels > 0) b.launch_xchg_level(0, n_global);
    else if (!ingest) b.launch_level(0, nullptr, lps[0]);
    {
        // the persistent cascade kernel takes over at the first level whose predecessor is surely plain; the
        // bucketed prefix (and the level right after it) is launched step by step
        uint32_t l = first_level;
        for (; l < lps.size(); l++) {
            if (l < b.xchg_levels) {
                b.launch_xchg_level(l, lps[l].k_up);
                continue;
            }
            if (comm && l == st.gather_level) {
                b.launch_gather(l, lps[l - 1].k_up, lps[l].k_up);
                l++;
            }
            if (l >= lps.size()) break;
            if (b.use_cascade() && !lps[l - 1].maybe_bucket) break;
            b.launch_level(l, &lps[l - 1], lps[l]);
        }
        if (l < lps.size()) b.launch_cascade();
    }
    // the first tail-owned level right after a bucketed one gets its keys from the scatter step as a plain list
    if (lps.back().maybe_bucket)
        b.launch_scatter((uint32_t)lps.size(), lps.back().range_max, lps.back().count_max, 1);

    trace.mark("enqueued");
    for (;;) {
        b.launch_tail();
        b.download_state();
        trace.mark("state");
        if (st.status == ST_DONE) break;
        if (st.status == ST_BUCKET_OVERFLOW) return BUILD_RETRY_PLAIN;
        if (st.status == ST_ERR_MAX_ITERS) {
            char msg[128];
            snprintf(msg, sizeof msg, "counldn't find unique hashes (ran out of key space, %llu keys left)",
                     (unsigned long long)st.err_info);
            return fail(BPHF_ERR_MAX_ITERS, msg);
        }
        if (st.status == ST_ERR_INTERNAL) {
            char msg[128];
            snprintf(msg, sizeof msg, "internal consistency check failed (info %llx)", (unsigned long long)st.err_info);
            return fail(BPHF_ERR_INTERNAL, msg);
        }
        if (st.status == ST_ERR_COMM) {
            char msg[160];
            snprintf(msg, sizeof msg, "sharded build: a peer shard failed or did not reach a barrier in time (level %llu, peers %llx)",
                     (unsigned long long)(st.err_info >> 48), (unsigned long long)(st.err_info & 0xffffffffULL));
            return fail(BPHF_ERR_CUDA, msg);
        }
        if (st.status == ST_ERR_LIST_OVERFLOW)
            return fail(BPHF_ERR_NOMEM, "sharded build: this shard's share of a level outgrew its redo list (unbalanced shards)");
        const uint32_t level = st.n_levels;  // the level that could not be set up / has not run yet
        if (level == 0 || level > BPHF_MAX_LEVELS) return fail(BPHF_ERR_INTERNAL, "bad resume level");
        if (st.status == ST_NEED_WORDS || st.status == ST_NEED_LIST) {
            grow_until_fits(level);
        } else if (st.status != ST_RUNNING) {
            return fail(BPHF_ERR_INTERNAL, "unknown build status");
        }
        if (b.use_cascade() && st.lv[level - 1].b_range == 0 && st.lv[level].b_range == 0) {
            b.launch_cascade();
            continue;
        }
        // enqueue a few more levels from where the device stopped (k only shrinks from here); both kernel
        // families are enqueued because the modes of the coming levels are decided on the device
        const bool prev_b = st.lv[level - 1].b_range != 0, cur_b = st.lv[level].b_range != 0;
        LevelPlan pv, cu;
        pv.k_up = pv.k_lo = st.lv[level - 1].k_in;
        pv.maybe_bucket = pv.surely_bucket = prev_b;
        pv.range_max = prev_b ? st.lv[level - 1].b_range : 0;
        pv.count_max = prev_b ? st.lv[level - 1].b_count : 1;
        cu.k_up = cu.k_lo = st.lv[level].k_in;
        cu.maybe_bucket = cu.surely_bucket = cur_b;
        cu.range_max = cur_b ? st.lv[level].b_range : 0;
        cu.count_max = cur_b ? st.lv[level].b_count : 1;
        b.launch_level(level, &pv, cu);
        for (uint32_t l = level + 1; l < level + 4 && l <= BPHF_MAX_LEVELS; l++) {
            LevelPlan nx;
            nx.k_up = cu.k_up;
            nx.k_lo = 0;
            nx.maybe_bucket = cu.maybe_bucket;
            nx.surely_bucket = false;
            nx.range_max = BUCKET_R_MAX;
            nx.count_max = std::min<uint32_t>(BUCKET_MAX_COUNT, cu.count_max + b.slots());
            b.launch_level(l, &cu, nx);
            cu = nx;
        }
    }

    // ---- compute_ranks: exclusive scan of block popcounts over the whole arena
    std::unique_ptr<bphf_mphf> m(new (std::nothrow) bphf_mphf());
    if (!m) return fail(BPHF_ERR_NOMEM, "host allocation failed");
    m->device = device;
    m->kc = st.kc;
    m->kc.s_bytes = nullptr;  // the caller's segment tables were only borrowed for the build
    m->kc.s_seg_ends = nullptr;
    m->kc.s_key_segs = nullptr;
    m->n_levels = st.n_levels;
    m->words_used = st.words_used;
    m->lv.assign(st.lv, st.lv + st.n_levels);
    b.launch_xpop();  // exchange levels: rank popcounts of the complete level (after the barrier that ended the build)
    const uint64_t n_blocks = st.words_used / 8;
    handle_alloc(&m->d_ranks, std::max<uint64_t>(n_blocks, 1) * 8, device, s);
    {
        const uint64_t n_chunks = (n_blocks + SCAN_CHUNK - 1) / SCAN_CHUNK;
        uint64_t* d_sums = nullptr;
        CK(cudaMallocAsync((void**)&d_sums, std::max<uint64_t>(n_chunks, 1) * 8, s));
        const size_t ev = b.begin_event(3, 0);
        if (n_chunks) {
            scan_sums_kernel<<<(unsigned)n_chunks, SCAN_THREADS, 0, s>>>(b.d_blockpop, n_blocks, d_sums);
            scan_top_kernel<<<1, 1024, 0, s>>>(d_sums, n_chunks);
            scan_apply_kernel<<<(unsigned)n_chunks, SCAN_THREADS, 0, s>>>(b.d_blockpop, n_blocks, d_sums, m->d_ranks);
            CK(cudaGetLastError());
            b.launches += 3;
        }
        b.end_event(ev);
        CK(cudaFreeAsync(d_sums, s));
    }
    if (comm) {
        // the arena belongs to the comm (peers write into it during the next build): the handle gets a copy
        handle_alloc(&m->d_words, std::max<uint64_t>(st.words_used, 8) * 8, device, s);
        CK(cudaMemcpyAsync(m->d_words, b.d_words, st.words_used * 8, cudaMemcpyDeviceToDevice, s));
    } else {
        m->d_words = b.d_words;  // the handle keeps the arena (padded, <= planned capacity)
        b.d_words = nullptr;
    }
    upload_level_table(m.get(), s);  // level tables + packed lookup structure
    CK(cudaEventRecord(b.ev_end, s));
    trace.mark("epilogue");
    CK(cudaStreamSynchronize(s));
    trace.mark("synced");
    b.syncs++;
    b.stream_idle = true;  // nothing of this build is in flight any more: its list buffers may be parked

    // ---- statistics
    bphf_build_stats& bs = m->stats;
    bs.n_levels = st.n_levels;
    bs.tail_level = st.tail_level;
    uint64_t total_keys = 0;
    for (uint32_t l = 0; l < st.n_levels; l++) {
        bs.keys_in[l] = st.lv[l].k_in;
        bs.bits[l] = st.lv[l].bits;
    }
    total_keys = n_global;
    m->n_keys = total_keys;
    bs.kernel_launches = b.launches;
    bs.host_syncs = b.syncs;
    CK(cudaEventElapsedTime(&bs.total_ms, b.ev_begin, b.ev_end));
    for (auto& e : b.events) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e.a, e.b));
        if (e.kind == 0 && e.level <= BPHF_MAX_LEVELS) bs.pass_ms[e.level] += ms;
        else if (e.kind == 1 && e.level <= BPHF_MAX_LEVELS) bs.fin_ms[e.level] += ms;
        else if (e.kind == 2) bs.tail_ms += ms;
        else if (e.kind == 3) bs.ranks_ms += ms;
        else if (e.kind == 4 && e.level <= BPHF_MAX_LEVELS) bs.merge_ms[e.level] += ms;
        else if (e.kind == 5) bs.cascade_ms += ms;
    }
    bs.h2d_ms = 0.f;
    uint32_t nb = 0;
    for (uint32_t l = 0; l < st.n_levels; l++) nb += (st.lv[l].b_range != 0) + 1000u * (st.lv[l].x_slice_bits != 0);
    bs.bucketed_levels = nb;
    bs.rebuilds = rebuilds;
    bs.streamed_ingest = ingest ? 1 : 0;
    if (want_peak) {
        cudaMemPool_t pool;
        uint64_t high = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, &high) == cudaSuccess)
            bs.peak_device_bytes = high;
    }
    trace.mark("stats");
    *out = m.release();
    return BPHF_OK;
}

// ---- ingest of HOST key sets ---------------------------------------------------------------------------------
// 0 (default): a key set whose 8 n bytes plus the work buffers of an in-core build do not fit the device's free memory
// is streamed through a device ring (two passes over PCIe; level 1's list is where it becomes resident), anything
// else is copied whole (one pass); 1: always stream; 2: never.
static int g_ingest_mode = 0;
static uint64_t g_ingest_ring_keys = 16ull << 20;  // 128 MiB per ring slot

static bool want_streamed_ingest(uint64_t n, double gamma, int device) {
    if (g_ingest_mode == 2 || n == 0) return false;
    if (g_ingest_mode == 1) return true;
    // in-core: the keys, the two key lists (bucket regions of level 0 and 1 when level 0 exceeds the L2: ~1.1 n and
    // ~0.5 n keys; plain: ~0.45 n and ~0.2 n) and the (a, collide) arenas (~3.1 gamma/1.7 bits per key each)
    const bool bucketed = g_build_mode == 2 || (g_build_mode == 0 && (double)level_size(gamma, n) > bucket_fit_bits());
    const double lists = bucketed ? 1.6 : 0.65;
    const double need = 8.0 * (double)n * (1.0 + lists) + 2.0 * (double)n * gamma * 1.85 / 8.0 + 512e6;
    if (need < 0.25 * (double)device_info(device).total_mem) return false;  // small next to the device: skip the query
    DeviceGuard guard(device);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
    cudaMemPool_t pool;
    uint64_t reserved = 0, used = 0;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
        free_b += reserved - used;  // cached by the stream-ordered pool: available to the build
    return need > 0.92 * (double)free_b;
}

static int build_impl(const uint64_t* keys, uint64_t n, double gamma, int has_seed, uint64_t seed, int device,
                      int keys_mem, void* stream, bphf_mphf** out, const ChunkSpec* chunk = nullptr) {
    if (!out) return fail(BPHF_ERR_ARG, "out is null");
    *out = nullptr;
    if (n && !keys) return fail(BPHF_ERR_ARG, "keys is null");
    if (keys_mem != BPHF_MEM_HOST && keys_mem != BPHF_MEM_DEVICE) return fail(BPHF_ERR_ARG, "bad keys_mem");
    // assert!(gamma > 1.01)  src/lib.rs:236,364  (NaN fails the assertion too)
    if (!(gamma > 1.01)) return fail(BPHF_ERR_GAMMA, "assertion failed: gamma > 1.01");
    if (keys_mem == BPHF_MEM_HOST && (!chunk || (chunk->kc.kind == BPHF_KIND_U64 && chunk->ring_keys == 0)) &&
        want_streamed_ingest(n, gamma, device)) {
        ChunkSpec spec = chunk ? *chunk : ChunkSpec();
        spec.seg_ptr = &keys;
        spec.seg_len = &n;
        spec.n_segs = 1;
        spec.ring_keys = std::max<uint64_t>(g_ingest_ring_keys, 4096);
        const int rc = build_once(nullptr, n, gamma, has_seed, seed, device, keys_mem, stream, false, 0, 0, out, nullptr, 0, &spec);
        if (rc != BUILD_RETRY_INCORE) return rc;  // else: too few keys survive level 0 for the streamed schedule
    }
    // mode 0: levels whose (a, collide) pair does not fit the L2 are built by the bucketed kernels
    bool bucket = g_build_mode == 2;
    uint64_t min_keys = g_bucket_min_keys;
    if (chunk && chunk->kc.kind == BPHF_KIND_STREAM) bucket = false;  // the bucketed kernels exist for word keys only
    else if (g_build_mode == 0) {
        const double fit_bits = bucket_fit_bits();
        bucket = (double)level_size(gamma, n) > fit_bits;
        min_keys = (uint64_t)(fit_bits / gamma);
    }
    int rc = build_once(keys, n, gamma, has_seed, seed, device, keys_mem, stream, bucket, min_keys, 0, out, nullptr, 0, chunk);
    if (rc == BUILD_RETRY_PLAIN) {
        // a bucket region overflowed (heavily duplicated or otherwise non-random input): the plain L2-atomic
        // kernels have no such limit.  Still the GPU path -- there is no CPU fallback anywhere.
        rc = build_once(keys, n, gamma, has_seed, seed, device, keys_mem, stream, false, 0, 1, out, nullptr, 0, chunk);
        if (rc == BUILD_RETRY_PLAIN) return fail(BPHF_ERR_INTERNAL, "bucket overflow reported by an unbucketed build");
    }
    return rc;
}

// ------------------------------------------------------------------------------------------------
// exported functions
// ------------------------------------------------------------------------------------------------
#define API_BEGIN try {
#define API_END                                      \
    }                                                \
    catch (const CudaFailure& f) {                   \
        return fail(f.code, f.msg);                  \
    }                                                \
    catch (const std::bad_alloc&) {                  \
        return fail(BPHF_ERR_NOMEM, "host allocation failed"); \
    }                                                \
    catch (...) {                                    \
        return fail(BPHF_ERR_INTERNAL, "unexpected exception"); \
    }

// stage host arrays on the device for the map helpers (simple, unpipelined)
struct Staged {
    cudaStream_t s;
    std::vector<void*> ptrs;
    explicit Staged(cudaStream_t st) : s(st) {}
    ~Staged() {
        for (void* p : ptrs) cudaFreeAsync(p, s);
    }
    template <typename T>
    T* in(const T* host, uint64_t count) {
        if (!host || !count) return nullptr;
        T* d = nullptr;
        CK(cudaMallocAsync((void**)&d, count * sizeof(T), s));
        ptrs.push_back(d);
        CK(cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
        return d;
    }
    template <typename T>
    T* scratch(uint64_t count) {
        T* d = nullptr;
        CK(cudaMallocAsync((void**)&d, std::max<uint64_t>(count, 1) * sizeof(T), s));
        ptrs.push_back(d);
        return d;
    }
};

extern "C" {

BPHF_API int bphf_abi_version(void) { return 1; }
BPHF_API const char* bphf_last_error(void) { return g_last_error.c_str(); }
BPHF_API int bphf_wy_swap_8byte_tail(void) { return BPHF_WY_SWAP_8BYTE_TAIL; }

BPHF_API int bphf_device_count(int* count) {
    API_BEGIN
    if (!count) return fail(BPHF_ERR_ARG, "count is null");
    *count = 0;
    CK(cudaGetDeviceCount(count));
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_set_build_mode(int mode, uint64_t bucket_min_keys) {
    g_build_mode = (mode == 1 || mode == 2) ? mode : 0;
    g_bucket_min_keys = bucket_min_keys ? bucket_min_keys : (1ull << 20);
    return BPHF_OK;
}

BPHF_API int bphf_set_xchg_bits(double bits) {
    g_xchg_bits = bits > 0.0 ? bits : 0.0;
    return BPHF_OK;
}

BPHF_API int bphf_set_ingest_mode(int mode, uint64_t ring_keys) {
    if (mode < 0 || mode > 2) return fail(BPHF_ERR_ARG, "ingest mode is 0 (automatic), 1 (always streamed) or 2 (never)");
    g_ingest_mode = mode;
    g_ingest_ring_keys = ring_keys ? ring_keys : (16ull << 20);
    return BPHF_OK;
}

BPHF_API int bphf_set_query_mode(int mode) {
    g_query_mode = mode == 1 ? 1 : 0;
    return BPHF_OK;
}

BPHF_API int bphf_set_query_partition(int mode, uint32_t slice_shift, uint64_t walk_bytes, uint64_t batch_keys, double cap_scale) {
    if (mode < 0 || mode > 2) return fail(BPHF_ERR_ARG, "partition mode is 0 (automatic), 1 (never) or 2 (whenever possible)");
    if (slice_shift > 31) return fail(BPHF_ERR_ARG, "slice_shift is at most 31 (levels of the packed structure have < 2^32 bits)");
    std::lock_guard<std::mutex> lock(g_qpart_mutex);
    const QPartCfg def;
    g_qpart.mode = mode;
    g_qpart.slice_shift = slice_shift ? slice_shift : def.slice_shift;
    g_qpart.walk_bytes = walk_bytes ? walk_bytes : def.walk_bytes;
    g_qpart.batch_keys = batch_keys ? batch_keys : def.batch_keys;
    g_qpart.cap_scale = cap_scale > 0.0 ? std::min(cap_scale, 1.5) : def.cap_scale;
    return BPHF_OK;
}

BPHF_API int bphf_query_partition_stats(int device, uint64_t* batches, uint64_t* fallbacks, uint32_t* stages) {
    API_BEGIN
    DeviceGuard guard(device);
    if (batches) *batches = g_qp_batches.load();
    if (stages) *stages = g_qp_last_stages.load();
    if (fallbacks) {
        unsigned long long f = 0;
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpyFromSymbol(&f, g_qp_fallbacks, sizeof f));
        *fallbacks = f;
    }
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_set_profile(int level) {
    g_profile = level < 0 ? 0 : (level > 2 ? 2 : level);
    return BPHF_OK;
}

BPHF_API int bphf_build_u64(const uint64_t* keys, uint64_t n, double gamma, int has_seed, uint64_t seed, int device,
                            int keys_mem, void* stream, bphf_mphf** out) {
    API_BEGIN
    return build_impl(keys, n, gamma, has_seed, seed, device, keys_mem, stream, out);
    API_END
}

// ---- chunked construction --------------------------------------------------------
// Mphf::from_chunked_iterator (src/lib.rs:112-226) and from_chunked_iterator_parallel (:539-667) for T = u64.
// The chunks are gathered into one device array as they are handed over (H2D per chunk for host chunks) and the
// cascade runs with the reference's chunked semantics.  The key set has to fit the GPU's memory (180 GB ~ 1e10
// keys with the work buffers); what the reference adds beyond that -- re-iterating an out-of-core source once per
// level -- is not reproduced.
BPHF_API int bphf_build_chunked_u64(const uint64_t* const* chunks, const uint64_t* chunk_lens, uint64_t n_chunks, uint64_t n,
                                    double gamma, int parallel_variant, int has_max_iters, uint64_t max_iters, int device,
                                    int keys_mem, bphf_mphf** out) {
    API_BEGIN
    if (!out) return fail(BPHF_ERR_ARG, "out is null");
    *out = nullptr;
    if (n_chunks && (!chunks || !chunk_lens)) return fail(BPHF_ERR_ARG, "null chunk table");
    if (keys_mem != BPHF_MEM_HOST && keys_mem != BPHF_MEM_DEVICE) return fail(BPHF_ERR_ARG, "bad keys_mem");
    uint64_t total = 0;
    for (uint64_t c = 0; c < n_chunks; c++) {
        if (chunk_lens[c] && !chunks[c]) return fail(BPHF_ERR_ARG, "null chunk");
        total += chunk_lens[c];
    }
    // the reference walks exactly n items: fewer panics ("max number of items overflowed"), more never terminates
    if (total != n) return fail(BPHF_ERR_ARG, "n does not match the total length of the chunks");
    if (!(gamma > 1.01)) return fail(BPHF_ERR_GAMMA, "assertion failed: gamma > 1.01");  // src/lib.rs:125,562
    device_info(device);
    DeviceGuard guard(device);
    cudaStream_t s = own_stream(device, false);
    if (parallel_variant && n == 0) {
        // keys_remaining == 0 on entry: the loop breaks before the first level (src/lib.rs:580-582) -> no levels
        std::unique_ptr<bphf_mphf> m(new bphf_mphf());
        m->device = device;
        handle_alloc(&m->d_words, 64, device, s);
        handle_alloc(&m->d_ranks, 8, device, s);
        upload_level_table(m.get(), s);
        *out = m.release();
        return BPHF_OK;
    }
    ChunkSpec cs;
    if (parallel_variant) {
        cs.mode = 1;
        cs.thr = (uint64_t)(0.01f * (float)n);  // f32 arithmetic as in src/lib.rs:556-557
        cs.has_max = has_max_iters ? 1 : 0;
        cs.max_iters = max_iters;
    }
    // host chunks that would not fit the device next to the work buffers: streamed, chunk by chunk, twice -- the
    // reference's own reason for this constructor (it re-iterates the source once per level, src/lib.rs:103-111)
    if (keys_mem == BPHF_MEM_HOST && want_streamed_ingest(n, gamma, device)) {
        ChunkSpec spec = cs;
        spec.seg_ptr = chunks;
        spec.seg_len = chunk_lens;
        spec.n_segs = n_chunks;
        spec.ring_keys = std::max<uint64_t>(g_ingest_ring_keys, 4096);
        const int rc = build_once(nullptr, n, gamma, 0, 0, device, keys_mem, s, false, 0, 0, out, nullptr, 0, &spec);
        if (rc != BUILD_RETRY_INCORE) return rc;
    }
    // gather: one contiguous device array (a single device chunk is used in place)
    const uint64_t* d_keys = nullptr;
    uint64_t* owned = nullptr;
    if (keys_mem == BPHF_MEM_DEVICE && n_chunks == 1) {
        d_keys = chunks[0];
    } else if (n) {
        CK(cudaMallocAsync((void**)&owned, n * 8, s));
        uint64_t off = 0;
        for (uint64_t c = 0; c < n_chunks; c++) {
            if (!chunk_lens[c]) continue;
            CK(cudaMemcpyAsync(owned + off, chunks[c], chunk_lens[c] * 8,
                               keys_mem == BPHF_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
            off += chunk_lens[c];
        }
        d_keys = owned;
    }
    int rc;
    try {
        rc = build_impl(d_keys, n, gamma, 0, 0, device, BPHF_MEM_DEVICE, s, out, parallel_variant ? &cs : nullptr);
    } catch (...) {
        if (owned) cudaFreeAsync(owned, s);
        throw;
    }
    if (owned) CK(cudaFreeAsync(owned, s));
    return rc;
    API_END
}

// ---- sharded build -------------------------------------------------------------------
BPHF_API int bphf_comm_create(int device, int rank, int world, uint64_t max_keys_global, double max_gamma, bphf_comm** out) {
    API_BEGIN
    if (!out) return fail(BPHF_ERR_ARG, "out is null");
    *out = nullptr;
    if (world < 1 || world > SHARD_MAX || rank < 0 || rank >= world) return fail(BPHF_ERR_ARG, "bad rank / world (world <= 8)");
    if (!(max_gamma > 1.01)) return fail(BPHF_ERR_GAMMA, "assertion failed: gamma > 1.01");
    const DeviceInfo& info = device_info(device);
    DeviceGuard guard(device);
    const Plan plan = make_plan(max_keys_global, max_gamma, (uint32_t)info.sm_count * 2, false, 0);
    std::unique_ptr<bphf_comm> c(new bphf_comm());
    c->device = device;
    c->rank = rank;
    c->world = world;
    c->phys_words = round_up8(plan.words_cap + plan.words_cap / 64 + TAIL_RESERVE_WORDS);
    // exchange blocks: a (sender, region, owner) block holds the keys of one warp region that fall into one owner's slice
    // -- about n_local / (R * world) of them; sized for shards up to twice the even share, + 6 sigma, in whole tiles
    c->x_regions = (uint32_t)info.sm_count * BPHF_PASS_CTAS * WP_WARPS;
    {
        const double mean = 2.0 * (double)max_keys_global / (double)world / (double)world / (double)c->x_regions;
        const uint64_t cap = (uint64_t)(mean + 6.0 * std::sqrt(mean) + 64.0);
        c->x_sub_cap = (uint32_t)std::min<uint64_t>(((cap + WP_TILE - 1) / WP_TILE) * WP_TILE, 1u << 30);
    }
    c->bytes = c->total_bytes();
    memset(&c->dev, 0, sizeof c->dev);
    c->dev.rank = (uint32_t)rank;
    c->dev.world = (uint32_t)world;
    c->dev.x_sub_cap = c->x_sub_cap;
    c->dev.x_regions = c->x_regions;
    {
        const char* e = getenv("BPHF_BARRIER_TIMEOUT_MS");
        const double ms = e ? atof(e) : 0.0;
        c->dev.barrier_timeout_ns = ms > 0.0 ? (uint64_t)(ms * 1e6) : BARRIER_TIMEOUT_NS_DEFAULT;
    }
    // plain cudaMalloc: exportable with cudaIpcGetMemHandle (pool memory is not)
    CK(cudaMalloc((void**)&c->base, c->bytes));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->push_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->push_ev, cudaEventDisableTiming));
    CK(cudaHostAlloc(&c->pinned, sizeof(BuildState), cudaHostAllocDefault));
    CK(cudaMemset(c->base, 0, c->bytes));
    CK(cudaDeviceSynchronize());
    c->map_peer(rank, c->base);
    if (world == 1) c->connected = true;
    *out = c.release();
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_comm_ipc_handle(const bphf_comm* c, void* handle_out) {
    API_BEGIN
    if (!c || !handle_out) return fail(BPHF_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == BPHF_IPC_HANDLE_BYTES, "IPC handle size");
    DeviceGuard guard(c->device);
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->base));
    memcpy(handle_out, &h, sizeof h);
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_comm_connect_ipc(bphf_comm* c, const void* handles) {
    API_BEGIN
    if (!c || !handles) return fail(BPHF_ERR_ARG, "null argument");
    DeviceGuard guard(c->device);
    for (int p = 0; p < c->world; p++) {
        if (p == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)p * BPHF_IPC_HANDLE_BYTES, sizeof h);
        void* b = nullptr;
        CK(cudaIpcOpenMemHandle(&b, h, cudaIpcMemLazyEnablePeerAccess));
        c->map_peer(p, b);
        c->peer_is_ipc[p] = true;
    }
    c->connected = true;
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_comm_connect_local(bphf_comm* c, bphf_comm* const* all) {
    API_BEGIN
    if (!c || !all) return fail(BPHF_ERR_ARG, "null argument");
    DeviceGuard guard(c->device);
    for (int p = 0; p < c->world; p++) {
        if (p == c->rank) continue;
        const bphf_comm* o = all[p];
        if (!o || o->rank != p || o->world != c->world || o->phys_words != c->phys_words || o->x_sub_cap != c->x_sub_cap ||
            o->x_regions != c->x_regions)
            return fail(BPHF_ERR_ARG, "comms do not match (rank order, world, arena size)");
        if (o->device == c->device) c->co_located = true;
        if (o->device != c->device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, c->device, o->device));
            if (!can) return fail(BPHF_ERR_CUDA, "devices cannot access each other's memory");
            cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            cudaGetLastError();
        }
        c->map_peer(p, o->base);
    }
    c->connected = true;
    return BPHF_OK;
    API_END
}

BPHF_API void bphf_comm_destroy(bphf_comm* c) {
    if (!c) return;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(c->device) == cudaSuccess) {
        cudaDeviceSynchronize();
        for (int p = 0; p < c->world; p++)
            if (c->peer_is_ipc[p] && c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
        if (c->base) cudaFree(c->base);
        if (c->stream) cudaStreamDestroy(c->stream);
        if (c->push_stream) cudaStreamDestroy(c->push_stream);
        if (c->push_ev) cudaEventDestroy(c->push_ev);
        if (c->pinned) cudaFreeHost(c->pinned);
    }
    if (prev >= 0) cudaSetDevice(prev);
    delete c;
}

BPHF_API int bphf_build_sharded_u64(bphf_comm* c, const uint64_t* keys, uint64_t n_local, uint64_t n_global, double gamma,
                                    int has_seed, uint64_t seed, int keys_mem, bphf_mphf** out) {
    API_BEGIN
    if (!out) return fail(BPHF_ERR_ARG, "out is null");
    *out = nullptr;
    if (!c) return fail(BPHF_ERR_ARG, "null comm");
    if (!c->connected) return fail(BPHF_ERR_ARG, "comm is not connected to its peers");
    if (n_local && !keys) return fail(BPHF_ERR_ARG, "keys is null");
    if (n_local > n_global) return fail(BPHF_ERR_ARG, "n_local > n_global");
    if (keys_mem != BPHF_MEM_HOST && keys_mem != BPHF_MEM_DEVICE) return fail(BPHF_ERR_ARG, "bad keys_mem");
    if (!(gamma > 1.01)) return fail(BPHF_ERR_GAMMA, "assertion failed: gamma > 1.01");
    // a one-shard "group" is an ordinary build -- unless BPHF_XCHG_WORLD1 asks for the sharded schedule anyway: the shard
    // then owns every slot and routes to itself, which is how the exchange kernels are profiled on one GPU (under ncu a
    // real peer would wait at a device barrier for kernels that the profiler replays one at a time)
    if (c->world == 1 && !getenv("BPHF_XCHG_WORLD1"))
        return build_impl(keys, n_local, gamma, has_seed, seed, c->device, keys_mem, nullptr, out);
    // same policy as an unsharded build, on the GLOBAL level sizes (every shard marks full-width levels)
    bool bucket = g_build_mode == 2;
    uint64_t min_keys = g_bucket_min_keys;
    if (g_build_mode == 0 && g_xchg_bits <= 0.0) {
        // exchange path switched off: every shard marks full-width levels, grouped into shared-memory buckets while they
        // exceed the L2 (the round-1 schedule)
        const double fit_bits = bucket_fit_bits();
        bucket = (double)level_size(gamma, n_global) > fit_bits;
        min_keys = (uint64_t)(fit_bits / gamma);
    }
    const int rc = build_once(keys, n_local, gamma, has_seed, seed, c->device, keys_mem, c->stream, bucket, min_keys, 0, out, c, n_global);
    if (rc == BUILD_RETRY_PLAIN)
        return fail(BPHF_ERR_NOMEM, "sharded build: a bucket region overflowed (unbalanced shards or heavily duplicated keys); "
                                    "use bphf_set_build_mode(1, 0) on every shard");
    return rc;
    API_END
}

static_assert((int)BPHF_KIND_U64 == (int)BPHF_KEY_U64 && (int)BPHF_KIND_UINT == (int)BPHF_KEY_UINT &&
                  (int)BPHF_KIND_BYTES == (int)BPHF_KEY_BYTES && (int)BPHF_KIND_STREAM == (int)BPHF_KEY_STREAM,
              "key kinds of hash.cuh and of the public header must agree");
static bool valid_key_kind(int kind, int width) {
    if (kind == BPHF_KIND_U64) return true;
    if (kind == BPHF_KIND_UINT) return width == 1 || width == 2 || width == 4;
    if (kind == BPHF_KIND_BYTES) return width >= 1 && width <= 8;
    return false;
}

// ---- keys other than u64 -----------------------------------------------------------
// packed keys of `width` bytes -> tail words (hash.cuh: key_tail_word), one u64 per key, on the device
static uint64_t* keys_to_words(const void* keys, uint64_t n, const KeyCfg& kc, int mem, cudaStream_t s, const DeviceInfo& info,
                               std::vector<void*>& scratch) {
    const size_t bytes = (size_t)n * kc.width;
    const uint8_t* d_in = static_cast<const uint8_t*>(keys);
    if (mem == BPHF_MEM_HOST) {
        void* tmp = nullptr;
        CK(cudaMallocAsync(&tmp, std::max<size_t>(bytes, 8), s));
        scratch.push_back(tmp);
        CK(cudaMemcpyAsync(tmp, keys, bytes, cudaMemcpyHostToDevice, s));
        d_in = static_cast<const uint8_t*>(tmp);
    }
    uint64_t* d_words = nullptr;
    CK(cudaMallocAsync((void**)&d_words, std::max<uint64_t>(n, 1) * 8, s));
    scratch.push_back(d_words);
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)info.sm_count * 16));
    key_words_kernel<<<grid, 256, 0, s>>>(d_in, kc.width, n, d_words);
    CK(cudaGetLastError());
    return d_words;
}

BPHF_API int bphf_build_keys(const void* keys, uint64_t n, int key_kind, int key_width, double gamma, int has_seed,
                             uint64_t seed, int device, int keys_mem, bphf_mphf** out) {
    API_BEGIN
    if (!out) return fail(BPHF_ERR_ARG, "out is null");
    *out = nullptr;
    if (key_kind == BPHF_KIND_STREAM) return fail(BPHF_ERR_ARG, "stream keys are built by bphf_build_stream");
    if (!valid_key_kind(key_kind, key_width)) return fail(BPHF_ERR_ARG, "unsupported key kind / width");
    if (key_kind == BPHF_KIND_U64)
        return build_impl(static_cast<const uint64_t*>(keys), n, gamma, has_seed, seed, device, keys_mem, nullptr, out);
    if (n && !keys) return fail(BPHF_ERR_ARG, "keys is null");
    if (keys_mem != BPHF_MEM_HOST && keys_mem != BPHF_MEM_DEVICE) return fail(BPHF_ERR_ARG, "bad keys_mem");
    if (!(gamma > 1.01)) return fail(BPHF_ERR_GAMMA, "assertion failed: gamma > 1.01");
    const DeviceInfo& info = device_info(device);
    DeviceGuard guard(device);
    cudaStream_t s = own_stream(device, false);
    ChunkSpec spec;
    spec.kc = make_key_cfg((uint32_t)key_kind, (uint32_t)key_width);
    std::vector<void*> scratch;
    int rc;
    try {
        const uint64_t* d_words = keys_to_words(keys, n, spec.kc, keys_mem, s, info, scratch);
        rc = build_impl(d_words, n, gamma, has_seed, seed, device, BPHF_MEM_DEVICE, s, out, &spec);
    } catch (...) {
        for (void* p : scratch) cudaFreeAsync(p, s);
        throw;
    }
    for (void* p : scratch) cudaFreeAsync(p, s);
    return rc;
    API_END
}

// ---- stream keys: any T: Hash as the recorded sequence of Hasher::write calls (hash.cuh) -------------------
// The caller's three tables on the device (copied when they live on the host), structurally checked.
static KeyCfg stage_stream(const uint8_t* bytes, uint64_t n_bytes, const uint64_t* seg_ends, uint64_t n_segs,
                           const uint64_t* key_segs, uint64_t n_keys, int mem, cudaStream_t s, const DeviceInfo& info,
                           std::vector<void*>& scratch) {
    if (!key_segs || (n_segs && !seg_ends) || (n_bytes && !bytes))
        throw CudaFailure{BPHF_ERR_ARG, "stream keys: null table"};
    if (mem != BPHF_MEM_HOST && mem != BPHF_MEM_DEVICE) throw CudaFailure{BPHF_ERR_ARG, "bad mem"};
    KeyCfg kc = make_key_cfg(BPHF_KIND_STREAM, 0);
    auto stage = [&](const void* src, size_t size) -> const void* {
        if (mem == BPHF_MEM_DEVICE) return src;
        void* d = nullptr;
        CK(cudaMallocAsync(&d, std::max<size_t>(size, 8), s));
        scratch.push_back(d);
        if (size) CK(cudaMemcpyAsync(d, src, size, cudaMemcpyHostToDevice, s));
        return d;
    };
    kc.s_bytes = static_cast<const uint8_t*>(stage(bytes, n_bytes));
    kc.s_seg_ends = static_cast<const uint64_t*>(stage(seg_ends, n_segs * 8));
    kc.s_key_segs = static_cast<const uint64_t*>(stage(key_segs, (n_keys + 1) * 8));
    uint32_t* d_bad = nullptr;
    CK(cudaMallocAsync((void**)&d_bad, 4, s));
    scratch.push_back(d_bad);
    CK(cudaMemsetAsync(d_bad, 0, 4, s));
    const uint64_t work = std::max(n_segs, n_keys);
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((work + 255) / 256, (uint64_t)info.sm_count * 8));
    stream_check_kernel<<<grid, 256, 0, s>>>(kc.s_seg_ends, n_segs, kc.s_key_segs, n_keys, n_bytes, d_bad);
    CK(cudaGetLastError());
    uint32_t bad = 0;
    CK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (bad) throw CudaFailure{BPHF_ERR_ARG, "stream keys: segment tables are not monotonic / do not add up"};
    return kc;
}

BPHF_API int bphf_build_stream(const uint8_t* bytes, uint64_t n_bytes, const uint64_t* seg_ends, uint64_t n_segs,
                               const uint64_t* key_segs, uint64_t n_keys, double gamma, int has_seed, uint64_t seed,
                               int device, int mem, bphf_mphf** out) {
    API_BEGIN
    if (!out) return fail(BPHF_ERR_ARG, "out is null");
    *out = nullptr;
    if (!(gamma > 1.01)) return fail(BPHF_ERR_GAMMA, "assertion failed: gamma > 1.01");
    const DeviceInfo& info = device_info(device);
    DeviceGuard guard(device);
    cudaStream_t s = own_stream(device, false);
    ChunkSpec spec;
    std::vector<void*> scratch;
    int rc;
    try {
        spec.kc = stage_stream(bytes, n_bytes, seg_ends, n_segs, key_segs, n_keys, mem, s, info, scratch);
        uint64_t* d_index = nullptr;  // level 0's key list: every key's index
        CK(cudaMallocAsync((void**)&d_index, std::max<uint64_t>(n_keys, 1) * 8, s));
        scratch.push_back(d_index);
        const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n_keys + 255) / 256, (uint64_t)info.sm_count * 16));
        iota_kernel<<<grid, 256, 0, s>>>(d_index, n_keys);
        CK(cudaGetLastError());
        rc = build_impl(d_index, n_keys, gamma, has_seed, seed, device, BPHF_MEM_DEVICE, s, out, &spec);
    } catch (...) {
        for (void* p : scratch) cudaFreeAsync(p, s);
        throw;
    }
    for (void* p : scratch) cudaFreeAsync(p, s);
    return rc;
    API_END
}

BPHF_API int bphf_key_kind(const bphf_mphf* m, int* key_kind, int* key_width) {
    if (!m) return fail(BPHF_ERR_ARG, "null handle");
    if (key_kind) *key_kind = (int)m->kc.kind;
    if (key_width) *key_width = (int)m->kc.width;
    return BPHF_OK;
}

BPHF_API int bphf_num_levels(const bphf_mphf* m, uint32_t* n_levels) {
    if (!m || !n_levels) return fail(BPHF_ERR_ARG, "null argument");
    *n_levels = m->n_levels;
    return BPHF_OK;
}

BPHF_API int bphf_level_dims(const bphf_mphf* m, uint32_t level, uint64_t* bits, uint64_t* n_words, uint64_t* n_ranks) {
    if (!m) return fail(BPHF_ERR_ARG, "null handle");
    if (level >= m->n_levels) return fail(BPHF_ERR_ARG, "that level doesn't exist");
    if (bits) *bits = m->lv[level].bits;
    if (n_words) *n_words = m->lv[level].n_words;
    if (n_ranks) *n_ranks = (m->lv[level].n_words + 7) / 8;
    return BPHF_OK;
}

BPHF_API int bphf_total_dims(const bphf_mphf* m, uint64_t* total_words, uint64_t* total_ranks, uint64_t* n_keys) {
    if (!m) return fail(BPHF_ERR_ARG, "null handle");
    uint64_t w = 0, r = 0;
    for (uint32_t l = 0; l < m->n_levels; l++) {
        w += m->lv[l].n_words;
        r += (m->lv[l].n_words + 7) / 8;
    }
    if (total_words) *total_words = w;
    if (total_ranks) *total_ranks = r;
    if (n_keys) *n_keys = m->n_keys;
    return BPHF_OK;
}

BPHF_API int bphf_level_export(const bphf_mphf* m, uint32_t level, uint64_t* words_out, uint64_t* ranks_out, int out_mem) {
    API_BEGIN
    if (!m) return fail(BPHF_ERR_ARG, "null handle");
    if (level >= m->n_levels) return fail(BPHF_ERR_ARG, "that level doesn't exist");
    DeviceGuard guard(m->device);
    cudaStream_t s = own_stream(m->device, false);
    const LevelDesc& d = m->lv[level];
    const cudaMemcpyKind kind = out_mem == BPHF_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (words_out && d.n_words) CK(cudaMemcpyAsync(words_out, m->d_words + d.word_off, d.n_words * 8, kind, s));
    if (ranks_out && d.n_words)
        CK(cudaMemcpyAsync(ranks_out, m->d_ranks + d.word_off / 8, ((d.n_words + 7) / 8) * 8, kind, s));
    CK(cudaStreamSynchronize(s));
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_export_all(const bphf_mphf* m, uint64_t* words_out, uint64_t* ranks_out, int out_mem) {
    API_BEGIN
    if (!m) return fail(BPHF_ERR_ARG, "null handle");
    DeviceGuard guard(m->device);
    cudaStream_t s = own_stream(m->device, false);
    const cudaMemcpyKind kind = out_mem == BPHF_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    uint64_t wo = 0, ro = 0;
    for (uint32_t l = 0; l < m->n_levels; l++) {
        const LevelDesc& d = m->lv[l];
        const uint64_t nr = (d.n_words + 7) / 8;
        if (words_out && d.n_words) CK(cudaMemcpyAsync(words_out + wo, m->d_words + d.word_off, d.n_words * 8, kind, s));
        if (ranks_out && nr) CK(cudaMemcpyAsync(ranks_out + ro, m->d_ranks + d.word_off / 8, nr * 8, kind, s));
        wo += d.n_words;
        ro += nr;
    }
    CK(cudaStreamSynchronize(s));
    return BPHF_OK;
    API_END
}

static int from_levels_impl(uint32_t n_levels, const uint64_t* bits, const uint64_t* const* words,
                            const uint64_t* const* ranks, KeyCfg kc, int device, bphf_mphf** out);

BPHF_API int bphf_from_levels(uint32_t n_levels, const uint64_t* bits, const uint64_t* const* words,
                              const uint64_t* const* ranks, int device, bphf_mphf** out) {
    API_BEGIN
    return from_levels_impl(n_levels, bits, words, ranks, make_key_cfg(BPHF_KIND_U64, 8), device, out);
    API_END
}

BPHF_API int bphf_from_levels_keys(uint32_t n_levels, const uint64_t* bits, const uint64_t* const* words,
                                   const uint64_t* const* ranks, int key_kind, int key_width, int device, bphf_mphf** out) {
    API_BEGIN
    if (key_kind != BPHF_KIND_STREAM && !valid_key_kind(key_kind, key_width))
        return fail(BPHF_ERR_ARG, "unsupported key kind / width");
    return from_levels_impl(n_levels, bits, words, ranks, make_key_cfg((uint32_t)key_kind, (uint32_t)key_width), device, out);
    API_END
}

static int from_levels_impl(uint32_t n_levels, const uint64_t* bits, const uint64_t* const* words,
                            const uint64_t* const* ranks, KeyCfg kc, int device, bphf_mphf** out) {
    if (!out) return fail(BPHF_ERR_ARG, "out is null");
    *out = nullptr;
    if (n_levels > BPHF_MAX_LEVELS || (n_levels && (!bits || !words || !ranks))) return fail(BPHF_ERR_ARG, "bad level arrays");
    device_info(device);
    DeviceGuard guard(device);
    cudaStream_t s = own_stream(device, false);
    std::unique_ptr<bphf_mphf> m(new bphf_mphf());
    m->device = device;
    m->kc = kc;
    m->n_levels = n_levels;
    m->lv.resize(n_levels);
    uint64_t off = 0;
    for (uint32_t l = 0; l < n_levels; l++) {
        LevelDesc& d = m->lv[l];
        d.bits = bits[l];
        d.n_words = (bits[l] + 63) / 64;
        d.word_off = off;
        d.k_in = 0;
        d.sp0 = d.sp0_query = level_spx(kc, l);
        off += round_up8(d.n_words);
        if (d.n_words && (!words[l] || !ranks[l])) return fail(BPHF_ERR_ARG, "null level buffer");
    }
    m->words_used = off;
    handle_alloc(&m->d_words, std::max<uint64_t>(off, 8) * 8, device, s);
    handle_alloc(&m->d_ranks, std::max<uint64_t>(off / 8, 1) * 8, device, s);
    CK(cudaMemsetAsync(m->d_words, 0, std::max<uint64_t>(off, 8) * 8, s));
    uint64_t n_keys = 0;
    for (uint32_t l = 0; l < n_levels; l++) {
        const LevelDesc& d = m->lv[l];
        if (!d.n_words) continue;
        CK(cudaMemcpyAsync(m->d_words + d.word_off, words[l], d.n_words * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(m->d_ranks + d.word_off / 8, ranks[l], ((d.n_words + 7) / 8) * 8, cudaMemcpyHostToDevice, s));
        for (uint64_t i = 0; i < d.n_words; i++) n_keys += (uint64_t)__builtin_popcountll(words[l][i]);
    }
    m->n_keys = n_keys;
    upload_level_table(m.get(), s);
    m->stats.n_levels = n_levels;
    *out = m.release();
    return BPHF_OK;
}

BPHF_API void bphf_free(bphf_mphf* m) {
    if (!m) return;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(m->device) == cudaSuccess) {
        cudaStream_t s = nullptr;
        try {
            s = handle_stream(m->device);
        } catch (...) {
        }
        if (m->d_words) cudaFreeAsync(m->d_words, s);
        if (m->d_ranks) cudaFreeAsync(m->d_ranks, s);
        if (m->d_lv) cudaFreeAsync(m->d_lv, s);
        if (m->d_lvp) cudaFreeAsync(m->d_lvp, s);
        if (m->d_packed) cudaFreeAsync(m->d_packed, s);
    }
    if (prev >= 0) cudaSetDevice(prev);
    delete m;
}

#ifdef BPHF_TRACE
// measurement builds only (not declared in the header): copies the cascade phase timestamps to the host
BPHF_API int bphf_debug_cascade_trace(uint64_t* out, uint64_t bytes) {
    API_BEGIN
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out, g_cascade_trace, std::min<uint64_t>(bytes, sizeof(g_cascade_trace))));
    return BPHF_OK;
    API_END
}
#endif

BPHF_API int bphf_get_build_stats(const bphf_mphf* m, bphf_build_stats* out) {
    if (!m || !out) return fail(BPHF_ERR_ARG, "null argument");
    *out = m->stats;
    return BPHF_OK;
}

// ---- lookups ---------------------------------------------------------------------------------
static int query_grid(const DeviceInfo& info, uint64_t n) {
    const uint64_t blocks = (n + QUERY_THREADS - 1) / QUERY_THREADS;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(blocks, (uint64_t)info.sm_count * 16));
}

// ---- partitioned lookup (qpart_kernels.cuh): structures larger than the L2 ---------------------------------------
struct QPartPlan {
    int n_stages = 0;                  // partitioned stages (levels 0..n_stages-1); the walk stage starts at level n_stages
    uint32_t shift[QP_MAX_STAGES] = {}, n_buckets[QP_MAX_STAGES] = {};
};
static QPartPlan qpart_plan(const bphf_mphf* m, const QPartCfg& cfg, uint64_t n) {
    QPartPlan p;
    if (cfg.mode == 1 || !m->d_lvp || m->kc.kind == BPHF_KIND_STREAM || g_query_mode != 0) return p;
    if (m->n_keys >= 0xfffffffeULL || m->n_levels < 2) return p;  // ranks travel as 32-bit values, 0xffffffff = None
    if (cfg.mode == 0 && (n < cfg.min_keys || m->total_gran * 16 < cfg.min_struct_bytes)) return p;
    uint64_t rest = m->total_gran * 16;  // bytes of granules of the levels not yet assigned to a stage
    for (uint32_t l = 0; l + 1 < m->n_levels && p.n_stages < QP_MAX_STAGES; l++) {
        if (rest <= cfg.walk_bytes) break;
        const uint64_t bits = m->lv[l].bits;
        uint32_t shift = std::max<uint32_t>(cfg.slice_shift, 5);
        while (((bits + (1ull << shift) - 1) >> shift) > QP_MAX_BUCKETS) shift++;
        const uint32_t nb = (uint32_t)((bits + (1ull << shift) - 1) >> shift);
        if (nb < 2) break;  // the level is one slice: nothing to group
        p.shift[p.n_stages] = shift;
        p.n_buckets[p.n_stages] = nb;
        p.n_stages++;
        rest -= (bits + GRAN_BITS - 1) / GRAN_BITS * 16;
    }
    return p;
}

static int qp_grid(const DeviceInfo& info, const void* fn, int threads) {
    static std::mutex mu;
    static std::vector<std::pair<const void*, int>> cache;
    std::lock_guard<std::mutex> lock(mu);
    for (auto& e : cache)
        if (e.first == fn) return info.sm_count * e.second;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    cache.emplace_back(fn, per_sm);
    return info.sm_count * per_sm;
}

static void launch_query_partitioned(const bphf_mphf* m, const DeviceInfo& info, const QPartCfg& cfg, const QPartPlan& plan,
                                     const uint64_t* keys, uint64_t n, uint64_t* out, cudaStream_t s) {
    const QView v = query_view(m);
    const int S = plan.n_stages;
    // (at most 2^29 keys per pass: the slots of a stage, ~1.1-2 per key, are 32-bit numbers)
    const uint64_t B = std::min<uint64_t>(n, std::min<uint64_t>(std::max<uint64_t>(cfg.batch_keys, QP_UNIT), 1ull << 29));
    auto round_unit = [](uint64_t x) { return (x + QP_UNIT - 1) / QP_UNIT * QP_UNIT; };
    // st[0] = the caller's keys, st[1..S] = partitioned stages (levels 0..S-1), st[S+1] = walk stage
    std::vector<QPStage> st(S + 2);
    Staged scratch(s);
    const size_t ctl_stage = (size_t)QP_MAX_LISTS * QP_CURSOR_STRIDE;
    const size_t ctl_words = (size_t)(S + 2) * ctl_stage + 32;
    uint32_t* d_ctl = scratch.scratch<uint32_t>(ctl_words);
    uint32_t* d_overflow = d_ctl + (size_t)(S + 2) * ctl_stage;
    for (int i = 0; i <= S + 1; i++) {
        QPStage& q = st[i];
        memset(&q, 0, sizeof q);
        q.cursor = d_ctl + (size_t)i * ctl_stage;
        if (i == 0) {
            q.n_buckets = 1;
            q.cap = (uint32_t)round_unit(B);
            q.res = scratch.scratch<uint32_t>(q.cap);  // the slot of stage 1 every key went to
            continue;
        }
        const bool walk = i == S + 1;
        q.level = (uint32_t)(i - 1);
        q.n_buckets = walk ? 1 : plan.n_buckets[i - 1];
        q.shift = walk ? 0 : plan.shift[i - 1];
        // a list receives B * (its bucket's share of the level's bits) / QP_SUBS keys at most when the queries hash
        // uniformly (every key of the batch may reach the stage: absent keys miss everywhere)
        const double share = walk ? 1.0 : std::min(1.0, (double)(1ull << q.shift) / (double)m->lv[q.level].bits);
        const double mean = (double)B * share / QP_SUBS;
        const uint64_t cap = (uint64_t)((mean * 1.03 + 6.0 * std::sqrt(mean) + QP_UNIT) * cfg.cap_scale);
        q.cap = (uint32_t)round_unit(std::max<uint64_t>(cap, 1));
        const uint64_t slots = (uint64_t)q.cap * q.n_buckets * QP_SUBS;  // < 2^32: slots travel as 32-bit words
        q.keys = scratch.scratch<uint64_t>(slots);
        q.res = scratch.scratch<uint32_t>(slots);
        if (!walk) q.went = scratch.scratch<uint8_t>(slots / 8);
    }
    const int g_split0 = qp_grid(info, (const void*)qfirst_kernel, QP_THREADS);
    const int g_split = qp_grid(info, (const void*)qprobe_kernel, QP_THREADS);
    const int g_rest0 = qp_grid(info, (const void*)qrestore_kernel<true>, QP_THREADS);
    const int g_rest = qp_grid(info, (const void*)qrestore_kernel<false>, QP_THREADS);
    const int g_walk = qp_grid(info, (const void*)qwalk_kernel, QS_THREADS);
    const int g_fall = qp_grid(info, (const void*)qfallback_kernel, QS_THREADS);
    unsigned long long* d_fallbacks = nullptr;
    CK(cudaGetSymbolAddress((void**)&d_fallbacks, g_qp_fallbacks));
    // BPHF_Q_PROFILE=1 (measurement aid): CUDA events around every kernel of the first batch, printed to stderr
    static const bool q_profile = getenv("BPHF_Q_PROFILE") != nullptr;
    std::vector<cudaEvent_t> ev;
    std::vector<std::string> ev_name;
    auto mark = [&](const char* name) {
        if (!q_profile) return;
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        CK(cudaEventRecord(e, s));
        ev.push_back(e);
        ev_name.push_back(name);
    };
    for (uint64_t lo = 0; lo < n; lo += B) {
        const uint64_t cnt = std::min(B, n - lo);
        const bool prof = q_profile && lo == 0;
        CK(cudaMemsetAsync(d_ctl, 0, ctl_words * 4, s));
        QPStage first = st[0];
        first.keys = const_cast<uint64_t*>(keys + lo);
        if (prof) mark("begin");
        qfirst_kernel<<<g_split0, QP_THREADS, 0, s>>>(v, first, st[1], cnt, d_overflow);
        if (prof) mark("split first");
        for (int i = 1; i <= S; i++) {
            qprobe_kernel<<<g_split, QP_THREADS, 0, s>>>(v, st[i], st[i + 1], d_overflow);
            if (prof) mark("split stage");
        }
        qwalk_kernel<<<g_walk, QS_THREADS, 0, s>>>(v, st[S + 1]);
        if (prof) mark("walk");
        for (int i = S; i >= 1; i--) {
            qrestore_kernel<false><<<g_rest, QP_THREADS, 0, s>>>(st[i], st[i + 1], 0, nullptr);
            if (prof) mark("restore stage");
        }
        qrestore_kernel<true><<<g_rest0, QP_THREADS, 0, s>>>(first, st[1], cnt, out + lo);
        if (prof) mark("restore first");
        qfallback_kernel<<<g_fall, QS_THREADS, 0, s>>>(v, keys + lo, cnt, out + lo, d_overflow, d_fallbacks);
        if (prof) mark("fallback check");
        CK(cudaGetLastError());
        g_qp_batches++;
    }
    if (!ev.empty()) {
        CK(cudaStreamSynchronize(s));
        std::string line = "[bphf q-profile] stages " + std::to_string(S) + ", batch " + std::to_string(std::min(B, n)) + " keys:";
        float total = 0;
        for (size_t i = 1; i < ev.size(); i++) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, ev[i - 1], ev[i]));
            total += ms;
            char buf[96];
            snprintf(buf, sizeof buf, " %s %.3f", ev_name[i].c_str(), ms);
            line += buf;
        }
        fprintf(stderr, "%s | total %.3f ms\n", line.c_str(), total);
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
    }
    g_qp_last_stages = (uint32_t)S;
}

// stream_kc: segment tables of the query batch for a BPHF_KIND_STREAM handle (keys == nullptr: key i is index i)
static void launch_query(const bphf_mphf* m, const DeviceInfo& info, const uint64_t* keys, uint64_t n, uint64_t* out, cudaStream_t s,
                         const KeyCfg* stream_kc = nullptr) {
    QView v = query_view(m);
    if (m->kc.kind == BPHF_KIND_STREAM) {
        if (!stream_kc) throw CudaFailure{BPHF_ERR_ARG, "this handle hashes stream keys: use bphf_hash_batch_stream"};
        v.kc = *stream_kc;
        if (v.lvp) query_kernel<true, StreamCfg><<<query_grid(info, n), QUERY_THREADS, 0, s>>>(v, keys, n, out);
        else query_kernel<false, StreamCfg><<<query_grid(info, n), QUERY_THREADS, 0, s>>>(v, keys, n, out);
        CK(cudaGetLastError());
        return;
    }
    if (v.lvp && g_query_mode == 0) {
        QPartCfg cfg;
        {
            std::lock_guard<std::mutex> lock(g_qpart_mutex);
            cfg = g_qpart;
        }
        const QPartPlan plan = qpart_plan(m, cfg, n);
        // (in place -- out == keys -- stays with the plain kernel: the fallback of a batch re-reads its keys)
        if (plan.n_stages > 0 && (const void*)keys != (const void*)out) {
            try {
                launch_query_partitioned(m, info, cfg, plan, keys, n, out, s);
                return;
            } catch (const CudaFailure& f) {
                // no room for the slice lists (~55 B per key and stage): the plain kernel needs no scratch.  Nothing
                // of a batch is visible before its last kernel, and batches already done are simply redone.
                if (f.code != BPHF_ERR_NOMEM) throw;
                (void)cudaGetLastError();
            }
        }
        const uint64_t tiles = (n + (uint64_t)QS_WTILE * QS_WARPS - 1) / ((uint64_t)QS_WTILE * QS_WARPS);
        // persistent CTAs: exactly one wave of resident CTAs, each walks tiles blockIdx, +gridDim, ...
        static int per_sm = 0;
        if (!per_sm) {
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, query_sync_kernel, QS_THREADS, 0));
            if (per_sm < 1) per_sm = 1;
        }
        const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)info.sm_count * per_sm));
        query_sync_kernel<<<grid, QS_THREADS, 0, s>>>(v, keys, n, out);
    } else if (v.lvp) query_kernel<true><<<query_grid(info, n), QUERY_THREADS, 0, s>>>(v, keys, n, out);
    else query_kernel<false><<<query_grid(info, n), QUERY_THREADS, 0, s>>>(v, keys, n, out);
    CK(cudaGetLastError());
}

BPHF_API int bphf_hash_batch_u64(const bphf_mphf* m, const uint64_t* keys, uint64_t n, uint64_t* out, int mem, void* stream) {
    API_BEGIN
    if (!m) return fail(BPHF_ERR_ARG, "null handle");
    if (n == 0) return BPHF_OK;
    if (!keys || !out) return fail(BPHF_ERR_ARG, "null buffer");
    if (m->kc.kind == BPHF_KIND_STREAM) return fail(BPHF_ERR_ARG, "this handle hashes stream keys: use bphf_hash_batch_stream");
    const DeviceInfo& info = device_info(m->device);
    DeviceGuard guard(m->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : own_stream(m->device, false);
    if (mem == BPHF_MEM_DEVICE) {
        launch_query(m, info, keys, n, out, s);
        if (!stream) CK(cudaStreamSynchronize(s));
        return BPHF_OK;
    }
    // host buffers: H2D -> kernel -> D2H pipelined in chunks over two streams
    const uint64_t chunk = 8ull << 20;  // 8 Mi keys = 64 MiB each way
    constexpr int NBUF = 3;
    const uint64_t cap = std::min(n, chunk);
    cudaStream_t cs = own_stream(m->device, true);
    // staging buffers and events; released on every exit path (an exception must not leak device memory)
    struct Pipe {
        cudaStream_t s;
        uint64_t* d_in[NBUF] = {};
        uint64_t* d_out[NBUF] = {};
        cudaEvent_t in_ready[NBUF] = {}, out_done[NBUF] = {};
        explicit Pipe(cudaStream_t st) : s(st) {}
        ~Pipe() {
            for (int i = 0; i < NBUF; i++) {
                if (d_in[i]) cudaFreeAsync(d_in[i], s);
                if (d_out[i]) cudaFreeAsync(d_out[i], s);
                if (in_ready[i]) cudaEventDestroy(in_ready[i]);
                if (out_done[i]) cudaEventDestroy(out_done[i]);
            }
        }
    } p(s);
    for (int i = 0; i < NBUF; i++) {
        CK(cudaMallocAsync((void**)&p.d_in[i], cap * 8, s));
        CK(cudaMallocAsync((void**)&p.d_out[i], cap * 8, s));
        CK(cudaEventCreateWithFlags(&p.in_ready[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p.out_done[i], cudaEventDisableTiming));
    }
    CK(cudaEventRecord(p.out_done[0], s));
    CK(cudaStreamWaitEvent(cs, p.out_done[0], 0));  // buffers exist before the copy stream touches them
    uint64_t c = 0;
    for (uint64_t lo = 0; lo < n; lo += chunk, c++) {
        const int bi = (int)(c % NBUF);
        const uint64_t cnt = std::min(chunk, n - lo);
        if (c >= (uint64_t)NBUF) CK(cudaStreamWaitEvent(cs, p.out_done[bi], 0));  // buffer free again
        CK(cudaMemcpyAsync(p.d_in[bi], keys + lo, cnt * 8, cudaMemcpyHostToDevice, cs));
        CK(cudaEventRecord(p.in_ready[bi], cs));
        CK(cudaStreamWaitEvent(s, p.in_ready[bi], 0));
        launch_query(m, info, p.d_in[bi], cnt, p.d_out[bi], s);
        CK(cudaMemcpyAsync(out + lo, p.d_out[bi], cnt * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(p.out_done[bi], s));
    }
    CK(cudaStreamSynchronize(cs));
    CK(cudaStreamSynchronize(s));
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_hash_batch_keys(const bphf_mphf* m, const void* keys, uint64_t n, uint64_t* out, int mem, void* stream) {
    API_BEGIN
    if (!m) return fail(BPHF_ERR_ARG, "null handle");
    if (m->kc.kind == BPHF_KIND_U64) return bphf_hash_batch_u64(m, static_cast<const uint64_t*>(keys), n, out, mem, stream);
    if (m->kc.kind == BPHF_KIND_STREAM) return fail(BPHF_ERR_ARG, "this handle hashes stream keys: use bphf_hash_batch_stream");
    if (n == 0) return BPHF_OK;
    if (!keys || !out) return fail(BPHF_ERR_ARG, "null buffer");
    const DeviceInfo& info = device_info(m->device);
    DeviceGuard guard(m->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : own_stream(m->device, false);
    std::vector<void*> scratch;
    try {
        const uint64_t* d_words = keys_to_words(keys, n, m->kc, mem, s, info, scratch);
        uint64_t* d_out = out;
        if (mem == BPHF_MEM_HOST) {
            CK(cudaMallocAsync((void**)&d_out, n * 8, s));
            scratch.push_back(d_out);
        }
        launch_query(m, info, d_words, n, d_out, s);
        if (mem == BPHF_MEM_HOST) CK(cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, s));
        for (void* p : scratch) CK(cudaFreeAsync(p, s));
        scratch.clear();
        CK(cudaStreamSynchronize(s));
    } catch (...) {
        for (void* p : scratch) cudaFreeAsync(p, s);
        throw;
    }
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_hash_batch_stream(const bphf_mphf* m, const uint8_t* bytes, uint64_t n_bytes, const uint64_t* seg_ends,
                                    uint64_t n_segs, const uint64_t* key_segs, uint64_t n_keys, uint64_t* out, int mem,
                                    void* stream) {
    API_BEGIN
    if (!m) return fail(BPHF_ERR_ARG, "null handle");
    if (m->kc.kind != BPHF_KIND_STREAM) return fail(BPHF_ERR_ARG, "this handle does not hash stream keys");
    if (n_keys == 0) return BPHF_OK;
    if (!out) return fail(BPHF_ERR_ARG, "null buffer");
    const DeviceInfo& info = device_info(m->device);
    DeviceGuard guard(m->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : own_stream(m->device, false);
    std::vector<void*> scratch;
    try {
        const KeyCfg kc = stage_stream(bytes, n_bytes, seg_ends, n_segs, key_segs, n_keys, mem, s, info, scratch);
        uint64_t* d_out = out;
        if (mem == BPHF_MEM_HOST) {
            CK(cudaMallocAsync((void**)&d_out, n_keys * 8, s));
            scratch.push_back(d_out);
        }
        launch_query(m, info, nullptr, n_keys, d_out, s, &kc);
        if (mem == BPHF_MEM_HOST) CK(cudaMemcpyAsync(out, d_out, n_keys * 8, cudaMemcpyDeviceToHost, s));
        for (void* p : scratch) CK(cudaFreeAsync(p, s));
        scratch.clear();
        CK(cudaStreamSynchronize(s));
    } catch (...) {
        for (void* p : scratch) cudaFreeAsync(p, s);
        throw;
    }
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_map_permute_u64(const bphf_mphf* m, const uint64_t* keys, const uint64_t* values, uint64_t n,
                                  uint64_t* keys_out, uint64_t* values_out, int mem, void* stream) {
    API_BEGIN
    if (!m) return fail(BPHF_ERR_ARG, "null handle");
    if (n == 0) return BPHF_OK;
    if (!keys || !keys_out || (values && !values_out)) return fail(BPHF_ERR_ARG, "null buffer");
    if (m->kc.kind != BPHF_KIND_U64) return fail(BPHF_ERR_ARG, "the map helpers take u64 keys");
    const DeviceInfo& info = device_info(m->device);
    DeviceGuard guard(m->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : own_stream(m->device, false);
    Staged st(s);
    const uint64_t *dk = keys, *dv = values;
    uint64_t *dko = keys_out, *dvo = values_out;
    if (mem == BPHF_MEM_HOST) {
        dk = st.in(keys, n);
        dv = st.in(values, n);
        dko = st.scratch<uint64_t>(n);
        dvo = values ? st.scratch<uint64_t>(n) : nullptr;
    }
    unsigned long long* d_bad = st.scratch<unsigned long long>(1);
    CK(cudaMemsetAsync(d_bad, 0, 8, s));
    if (m->d_lvp) {
        // ranks with the batched lookup kernel (warp-synchronous walk), then a plain scatter; 64 Mi keys at a time
        const uint64_t chunk = std::min<uint64_t>(n, 64ull << 20);
        uint64_t* d_pos = st.scratch<uint64_t>(chunk);
        for (uint64_t lo = 0; lo < n; lo += chunk) {
            const uint64_t cnt = std::min(chunk, n - lo);
            launch_query(m, info, dk + lo, cnt, d_pos, s);
            const int grid = (int)std::min<uint64_t>((cnt + 255) / 256, (uint64_t)info.sm_count * 16);
            map_scatter_kernel<<<grid, 256, 0, s>>>(d_pos, dk + lo, dv ? dv + lo : nullptr, cnt, n, dko, dvo, d_bad);
        }
    } else {
        map_permute_kernel<false><<<query_grid(info, n), QUERY_THREADS, 0, s>>>(query_view(m), dk, dv, n, dko, dvo, d_bad);
    }
    CK(cudaGetLastError());
    unsigned long long bad = 0;
    CK(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, s));
    if (mem == BPHF_MEM_HOST) {
        CK(cudaMemcpyAsync(keys_out, dko, n * 8, cudaMemcpyDeviceToHost, s));
        if (values) CK(cudaMemcpyAsync(values_out, dvo, n * 8, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    if (bad) return fail(BPHF_ERR_ARG, "a key was not in the construction set (hash out of range)");
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_map_get_batch_u64(const bphf_mphf* m, const uint64_t* map_keys, const uint64_t* map_values,
                                    uint64_t n_map, const uint64_t* queries, uint64_t nq, uint64_t* values_out,
                                    uint8_t* found_out, int mem, void* stream) {
    API_BEGIN
    if (!m) return fail(BPHF_ERR_ARG, "null handle");
    if (nq == 0) return BPHF_OK;
    if (!queries || !values_out || !found_out || (n_map && !map_keys)) return fail(BPHF_ERR_ARG, "null buffer");
    if (m->kc.kind != BPHF_KIND_U64) return fail(BPHF_ERR_ARG, "the map helpers take u64 keys");
    const DeviceInfo& info = device_info(m->device);
    DeviceGuard guard(m->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : own_stream(m->device, false);
    Staged st(s);
    const uint64_t *dmk = map_keys, *dmv = map_values, *dq = queries;
    uint64_t* dvo = values_out;
    uint8_t* dfo = found_out;
    if (mem == BPHF_MEM_HOST) {
        dmk = st.in(map_keys, n_map);
        dmv = st.in(map_values, n_map);
        dq = st.in(queries, nq);
        dvo = st.scratch<uint64_t>(nq);
        dfo = st.scratch<uint8_t>(nq);
    }
    if (m->d_lvp) {
        const uint64_t chunk = std::min<uint64_t>(nq, 64ull << 20);
        uint64_t* d_pos = st.scratch<uint64_t>(chunk);
        for (uint64_t lo = 0; lo < nq; lo += chunk) {
            const uint64_t cnt = std::min(chunk, nq - lo);
            launch_query(m, info, dq + lo, cnt, d_pos, s);
            const int grid = (int)std::min<uint64_t>((cnt + 255) / 256, (uint64_t)info.sm_count * 16);
            map_gather_kernel<<<grid, 256, 0, s>>>(d_pos, dmk, dmv, n_map, dq + lo, cnt, dvo + lo, dfo + lo);
        }
    } else {
        map_get_kernel<false><<<query_grid(info, nq), QUERY_THREADS, 0, s>>>(query_view(m), dmk, dmv, n_map, dq, nq, dvo, dfo);
    }
    CK(cudaGetLastError());
    if (mem == BPHF_MEM_HOST) {
        CK(cudaMemcpyAsync(values_out, dvo, nq * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(found_out, dfo, nq, cudaMemcpyDeviceToHost, s));
    }
    if (mem == BPHF_MEM_HOST || !stream) CK(cudaStreamSynchronize(s));
    return BPHF_OK;
    API_END
}

// ---- utilities --------------------------------------------------------------------------------
BPHF_API int bphf_l2_probe(int kind, uint64_t array_bytes, uint64_t n_ops, int device, float* ms_out) {
    API_BEGIN
    if (!ms_out || kind < 0 || kind > 2 || array_bytes < 4 || n_ops == 0) return fail(BPHF_ERR_ARG, "bad argument");
    const DeviceInfo& info = device_info(device);
    DeviceGuard guard(device);
    cudaStream_t s = own_stream(device, false);
    uint32_t *arr = nullptr, *sink = nullptr;
    CK(cudaMallocAsync((void**)&arr, array_bytes, s));
    CK(cudaMallocAsync((void**)&sink, 4, s));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaMemsetAsync(arr, 0, array_bytes, s));
        CK(cudaEventRecord(e0, s));
        l2_probe_kernel<<<info.sm_count * 8, 256, 0, s>>>(arr, array_bytes / 4, n_ops, kind, sink);
        CK(cudaEventRecord(e1, s));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CK(cudaFreeAsync(arr, s));
    CK(cudaFreeAsync(sink, s));
    CK(cudaStreamSynchronize(s));
    *ms_out = best;
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_keygen_u64(uint64_t* out_dev, uint64_t start, uint64_t n, uint64_t seed, int kind, int device, void* stream) {
    API_BEGIN
    if (n == 0) return BPHF_OK;
    if (!out_dev) return fail(BPHF_ERR_ARG, "null buffer");
    const DeviceInfo& info = device_info(device);
    DeviceGuard guard(device);
    cudaStream_t s = stream ? (cudaStream_t)stream : own_stream(device, false);
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)info.sm_count * 16);
    keygen_kernel<<<grid, 256, 0, s>>>(out_dev, start, n, seed, kind);
    CK(cudaGetLastError());
    if (!stream) CK(cudaStreamSynchronize(s));
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_check_permutation(const uint64_t* vals_dev, uint64_t n, int device, uint64_t* n_bad) {
    API_BEGIN
    if (!n_bad) return fail(BPHF_ERR_ARG, "null argument");
    *n_bad = 0;
    if (n == 0) return BPHF_OK;
    if (!vals_dev) return fail(BPHF_ERR_ARG, "null buffer");
    const DeviceInfo& info = device_info(device);
    DeviceGuard guard(device);
    cudaStream_t s = own_stream(device, false);
    Staged st(s);
    const uint64_t n32 = (n + 31) / 32;
    uint32_t* bitmap = st.scratch<uint32_t>(n32);
    unsigned long long* d_bad = st.scratch<unsigned long long>(1);
    CK(cudaMemsetAsync(bitmap, 0, n32 * 4, s));
    CK(cudaMemsetAsync(d_bad, 0, 8, s));
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)info.sm_count * 16);
    check_perm_kernel<<<grid, 256, 0, s>>>(vals_dev, n, bitmap, d_bad);
    CK(cudaGetLastError());
    unsigned long long bad = 0;
    CK(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *n_bad = bad;
    return BPHF_OK;
    API_END
}

BPHF_API int bphf_checksum_u64(const uint64_t* vals_dev, uint64_t n, int device, uint64_t* sum_out, uint64_t* xor_out) {
    API_BEGIN
    if (!sum_out || !xor_out) return fail(BPHF_ERR_ARG, "null argument");
    *sum_out = *xor_out = 0;
    if (n == 0) return BPHF_OK;
    const DeviceInfo& info = device_info(device);
    DeviceGuard guard(device);
    cudaStream_t s = own_stream(device, false);
    Staged st(s);
    unsigned long long* acc = st.scratch<unsigned long long>(2);
    CK(cudaMemsetAsync(acc, 0, 16, s));
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)info.sm_count * 16);
    checksum_kernel<<<grid, 256, 0, s>>>(vals_dev, n, acc);
    CK(cudaGetLastError());
    unsigned long long h[2];
    CK(cudaMemcpyAsync(h, acc, 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *sum_out = h[0];
    *xor_out = h[1];
    return BPHF_OK;
    API_END
}

BPHF_API void* bphf_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        g_last_error = "cudaMallocHost failed";
        return nullptr;
    }
    return p;
}
BPHF_API void bphf_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
