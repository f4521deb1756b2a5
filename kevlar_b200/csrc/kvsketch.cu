// kvsketch.cu -- host side of libkvsketch.so: the C ABI declared in include/kvsketch.h.
//
// One KvCtx per CUDA device holds the compute/copy streams, two input staging slots (so the
// H2D copy of batch i+1 overlaps the kernels of batch i) and the per-chunk scratch (hashes,
// valid bits).  Sketch tables live in ONE device allocation per sketch, khmer byte layout,
// each table 256-byte aligned.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include <cuda.h>   // driver API types only; entry points are fetched with cudaGetDriverEntryPoint (libcuda is never linked)

#include "kv_kernels.cuh"

// ------------------------------------------------------------------ errors

static thread_local std::string g_err;

static int kv_fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// error reporting for the other translation units of the library (kv_fastx.cpp)
int kv_fail_public(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return kv_fail(e_ == cudaErrorMemoryAllocation ? KV_ENOMEM : KV_ECUDA, "%s: %s (%s:%d)", #expr, \
                           cudaGetErrorString(e_), __FILE__, __LINE__);                              \
    } while (0)

#define KV_TRY(expr)            \
    do {                        \
        int rc_ = (expr);       \
        if (rc_ != KV_OK) return rc_; \
    } while (0)

extern "C" const char *kv_last_error(void) { return g_err.c_str(); }
extern "C" int kv_abi_version(void) { return KV_ABI_VERSION; }

// ------------------------------------------------------------------ context

enum { KV_PROF_OTHER = 0, KV_PROF_HASH = 1, KV_PROF_INCREMENT = 2, KV_PROF_UNIQUE = 3, KV_PROF_NOVEL = 4,
       KV_PROF_MERGE = 5, KV_PROF_FIXUP = 6, KV_PROF_PARTITION = 7 };   // KV_PROF_CLASSES (= 8) comes from kvsketch.h

struct KvBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct KvSlot {
    KvBuf bases, offsets;
    cudaEvent_t copied = nullptr, done = nullptr;
    bool used = false;
};

struct KvCtx {
    int device = -1;
    bool ready = false;
    std::mutex mu;
    cudaStream_t compute = nullptr, copy = nullptr;
    // the merge lane: between kv_merge_fork and kv_merge_join the peer-to-peer merge kernels (and the barriers of
    // lane-1 kv_peer_sync objects) run on `merge` next to whatever `compute` does meanwhile
    cudaStream_t merge = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool forked = false;
    int merge_lane_ctas = 1;                  // CTAs per SM of a merge kernel that runs next to the counting   [KV_MERGE_LANE_CTAS]
    KvSlot slot[2];
    int next_slot = 0;
    KvBuf tile_first, hashes, valid, fresh, first, hits, flags, discard, misc, added, part_items, part_small, hist;
    // sketches between these sizes take the region-partitioned update path (measured: 1.2-1.7x over
    // direct random atomics from 256 MB to 4 GB, break-even at 16 GB; profiles/r01_notes.md)
    uint64_t part_min_bytes = 128ull << 20, part_max_bytes = 8ull << 30;
    int part_region_log2 = 24;                // buckets per region (8-bit: 16 MB)
    // K3c, the tiled update path (default for sketches that do not fit L2): see kv_tile_apply_kernel
    int update_path = 0;                      // 0 auto, 1 direct, 2 region-partitioned (K3b), 3 tiled (K3c)   [KV_UPDATE_PATH]
    uint64_t tile_min_bytes = 128ull << 20;   // auto: sketches at least this large take the tiled path       [KV_TILE_MIN_BYTES]
    int tile_rb = 15;                         // log2(buckets per region)                                      [KV_TILE_RB]
    uint64_t tile_chunk_bases = 512ull << 20; // positions per chunk on the tiled path (256 M with n_unique tracking) [KV_TILE_CHUNK_BASES]
    int tile_direct_below = -1;               // regions with fewer offsets are updated in place (-1: region bytes / 64)
    int tile_block_log2 = 6;                  // slab layout: slots per interleave block (-1: run-major)               [KV_TILE_BLOCK_LOG2]
    KvBuf tile_cursor, tile_slab, tile_ovf, notes;
    unsigned *dirty = nullptr;   // device: one overflow flag per chunk, 64 slots used round-robin
    unsigned dirty_next = 0;
    unsigned long long *counters = nullptr;   // device: [0] n_valid  [1] n_unique  [2] n_hits  [3] occupied  [5] redone chunks
    unsigned long long *h_counters = nullptr; // pinned mirror
    uint8_t *h_stage = nullptr;               // pinned staging for sketch file I/O (lazily allocated)
    uint64_t launches = 0;
    // optional per-kernel-class timing (kv_profile): event pairs around launches
    bool profiling = false;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> prof_events;
    double prof_ms[KV_PROF_CLASSES] = {0};
    uint64_t prof_n[KV_PROF_CLASSES] = {0};
    int sm_count = 148;
    bool unique_fuse0 = true;            // table 0's first-touch pass inside the hash kernel + compact list (KV_NO_CLASSIFY switches it off)
    uint64_t first_range = 1ull << 31;   // buckets covered by the n_unique first[] scratch (4 B each; KV_FIRST_RANGE_LOG2)
    int first_epoch = 0;                 // epochs left before first[] must be memset again (0: memset first)
    KvBuf list_h, list_p, seg_cnt;
    size_t l2_persist = 0;     // bytes of L2 set aside for persisting accesses
    size_t l2_window_max = 0;
    uint64_t chunk_bases = 64ull << 20;
    // the batch whose hashes + valid bits still sit in `hashes` / `valid` (kv_unique_last_batch): set by a
    // single-chunk kv_consume_batch, dropped by whatever claims the scratch next
    struct { const kv_sketch *sketch = nullptr; uint64_t npos = 0; bool pass_a = false; uint32_t tag0 = 0; } last_hashed;
};

static KvCtx g_ctx[16];
static std::mutex g_ctx_mu;

static int kv_ctx_get(int device, KvCtx **out)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return kv_fail(KV_ENODEVICE, "no usable CUDA device (%s); libkvsketch has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n || device >= 16) return kv_fail(KV_EINVAL, "device %d out of range (have %d)", device, n);
    KvCtx &c = g_ctx[device];
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (!c.ready) {
        CU(cudaSetDevice(device));
        c.device = device;
        CU(cudaStreamCreateWithFlags(&c.compute, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking));
        {   // the merge lane outranks the compute stream: its few CTAs are placed as soon as slots free up
            // instead of queueing behind the half a million CTAs of a hash kernel
            int least = 0, greatest = 0;
            CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            CU(cudaStreamCreateWithPriority(&c.merge, cudaStreamNonBlocking, greatest));
            if (const char *env = getenv("KV_MERGE_LANE_CTAS")) c.merge_lane_ctas = std::max(1, atoi(env));
        }
        CU(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
        for (auto &s : c.slot) {
            CU(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        }
        CU(cudaMalloc(&c.counters, 8 * sizeof(unsigned long long)));
        CU(cudaMemset(c.counters, 0, 8 * sizeof(unsigned long long)));
        CU(cudaMallocHost(&c.h_counters, 8 * sizeof(unsigned long long)));
        CU(cudaMalloc(&c.dirty, 64 * sizeof(unsigned)));
        CU(cudaMemset(c.dirty, 0, 64 * sizeof(unsigned)));
        CU(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, device));
        {   // L2 residency control: let the structure a kernel hammers with random atomics persist in
            // L2 while its streaming inputs (hashes, bases) pass through
            int maxp = 0, maxw = 0;
            cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, device);
            cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, device);
            // measured on the benchmark config: +5 % on the increment kernel, -40 % on the novel scan
            // (less L2 left for its three sketches) -> opt-in only (profiles/r01_notes.md)
            if (!getenv("KV_L2_PERSIST")) maxp = 0;
            if (maxp > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp) == cudaSuccess) {
                c.l2_persist = (size_t)maxp;
                c.l2_window_max = (size_t)maxw;
            } else
                cudaGetLastError();
        }
        if (const char *env = getenv("KV_L2_FETCH")) {
            // L2 fetch granularity (32/64/128 B), opt-in.  Measured on 4 GB sketches and on C2
            // (profiles/r02r_l2fetch.jsonl): no difference at any setting, so the default is left alone.
            size_t want = (size_t)atoi(env);
            if (want && cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, want) != cudaSuccess) cudaGetLastError();
        }
        if (const char *env = getenv("KV_CHUNK_BASES")) {
            uint64_t v = strtoull(env, nullptr, 10);
            if (v >= KV_TILE) c.chunk_bases = v;
        }
        c.chunk_bases = (c.chunk_bases + KV_TILE - 1) / KV_TILE * KV_TILE;
        if (const char *env = getenv("KV_PART_MIN_BYTES")) c.part_min_bytes = strtoull(env, nullptr, 10);
        if (const char *env = getenv("KV_PART_MAX_BYTES")) c.part_max_bytes = strtoull(env, nullptr, 10);
        if (const char *env = getenv("KV_PART_REGION_LOG2")) c.part_region_log2 = std::max(4, std::min(30, atoi(env)));
        if (getenv("KV_NO_CLASSIFY")) c.unique_fuse0 = false;
        if (const char *env = getenv("KV_UPDATE_PATH"))
            c.update_path = !strcmp(env, "direct") ? 1 : !strcmp(env, "part") ? 2 : !strcmp(env, "tile") ? 3 : 0;
        if (const char *env = getenv("KV_TILE_MIN_BYTES")) c.tile_min_bytes = strtoull(env, nullptr, 10);
        if (const char *env = getenv("KV_TILE_RB")) c.tile_rb = std::max(6, std::min(16, atoi(env)));
        if (const char *env = getenv("KV_TILE_CHUNK_BASES")) c.tile_chunk_bases = std::max<uint64_t>(KV_TILE, strtoull(env, nullptr, 10));
        c.tile_chunk_bases = (c.tile_chunk_bases + KV_TILE - 1) / KV_TILE * KV_TILE;
        if (const char *env = getenv("KV_TILE_DIRECT_BELOW")) c.tile_direct_below = atoi(env);
        if (const char *env = getenv("KV_TILE_BLOCK_LOG2")) c.tile_block_log2 = std::max(-1, std::min(12, atoi(env)));
        if (const char *env = getenv("KV_FIRST_RANGE_LOG2")) c.first_range = 1ull << std::max(8, std::min(32, atoi(env)));
        c.ready = true;
    }
    *out = &c;
    return KV_OK;
}

static const size_t KV_IO_STAGE = 64u << 20;

static int kv_io_stage(KvCtx *ctx, uint8_t **out)
{
    if (!ctx->h_stage && cudaMallocHost((void **)&ctx->h_stage, KV_IO_STAGE) != cudaSuccess) {
        cudaGetLastError();
        ctx->h_stage = nullptr;
        return kv_fail(KV_ENOMEM, "cannot allocate pinned staging memory for sketch I/O");
    }
    *out = ctx->h_stage;
    return KV_OK;
}

// Sketch files are table bytes behind a small header, and production sketches are 4-72 GB (kevlar docs/tutorial.rst:51):
// one thread moves ~2 GB/s through the page cache, so the chunks that pass through the pinned staging buffer are read /
// written by several threads at once (pread / pwrite at their file offsets; KV_IO_THREADS, default min(8, cores / local
// ranks)), and the buffer is used in two halves so that the device copy of one chunk overlaps the file I/O of the other.
static int kv_io_threads()
{
    static int n = 0;
    if (!n) {
        unsigned hw = std::thread::hardware_concurrency();
        int ranks = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
        n = (int)std::max(1u, std::min(8u, (hw ? hw : 1u) / (unsigned)ranks));
        if (const char *e = getenv("KV_IO_THREADS")) n = std::max(1, std::min(64, atoi(e)));
    }
    return n;
}

struct KvFileJob {   // file I/O of one chunk, running on its own threads until join()
    std::vector<std::thread> workers;
    std::atomic<bool> failed{false};
    void start(int fd, uint8_t *buf, size_t n, uint64_t off, bool write)
    {
        const size_t min_part = 4u << 20;
        const int parts = (int)std::max<size_t>(1, std::min<size_t>((size_t)kv_io_threads(), (n + min_part - 1) / min_part));
        const size_t step = ((n + parts - 1) / parts + 4095) & ~(size_t)4095;
        for (int i = 0; i < parts; i++) {
            const size_t lo = std::min(n, (size_t)i * step), hi = std::min(n, lo + step);
            if (lo == hi) continue;
            workers.emplace_back([this, fd, buf, off, lo, hi, write] {
                size_t done = lo;
                while (done < hi) {
                    ssize_t got = write ? pwrite(fd, buf + done, hi - done, (off_t)(off + done)) : pread(fd, buf + done, hi - done, (off_t)(off + done));
                    if (got <= 0) { failed = true; return; }   // error, or (reading) the file ends early
                    done += (size_t)got;
                }
            });
        }
    }
    bool join()
    {
        for (std::thread &t : workers) t.join();
        workers.clear();
        return !failed;
    }
    ~KvFileJob() { join(); }
};

static int kv_buf_ensure(KvBuf &b, size_t need)
{
    if (need <= b.cap) return KV_OK;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t cap = need + need / 8 + 4096;
    cap = (cap + 255) & ~(size_t)255;
    CU(cudaMalloc(&b.p, cap));
    b.cap = cap;
    return KV_OK;
}

// every user of the hash scratch takes it through here: what an earlier batch left there is gone
static int kv_hash_scratch(KvCtx *ctx, uint64_t n_pos, bool want_hashes = true)
{
    ctx->last_hashed.sketch = nullptr;
    if (want_hashes) KV_TRY(kv_buf_ensure(ctx->hashes, n_pos * 8));
    return kv_buf_ensure(ctx->valid, (n_pos / 32 + 1) * 4);
}

#define LAUNCH_S(cls, ctx, strm, kern, grid, block, ...)                   \
    do {                                                                   \
        auto kfn_ = kern;                                                  \
        cudaStream_t st_ = (strm);                                         \
        cudaEvent_t e0_ = nullptr, e1_ = nullptr;                          \
        if ((ctx)->profiling) {                                            \
            cudaEventCreate(&e0_); cudaEventCreate(&e1_);                  \
            cudaEventRecord(e0_, st_);                                     \
        }                                                                  \
        kfn_<<<(grid), (block), 0, st_>>>(__VA_ARGS__);                    \
        if ((ctx)->profiling) {                                            \
            cudaEventRecord(e1_, st_);                                     \
            (ctx)->prof_events.push_back({cls, {e0_, e1_}});               \
        }                                                                  \
        (ctx)->launches++;                                                 \
        CU(cudaGetLastError());                                            \
    } while (0)
#define LAUNCH_C(cls, ctx, kern, grid, block, ...) LAUNCH_S(cls, ctx, (ctx)->compute, kern, grid, block, __VA_ARGS__)
#define LAUNCH(ctx, kern, grid, block, ...) LAUNCH_C(KV_PROF_OTHER, ctx, kern, grid, block, __VA_ARGS__)

// Mark [ptr, ptr+bytes) as the persisting L2 window of the compute stream (nullptr: none).
static void kv_l2_window(KvCtx *ctx, const void *ptr, size_t bytes)
{
    if (!ctx->l2_persist) return;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    if (ptr && bytes) {
        size_t win = std::min(bytes, ctx->l2_window_max);
        attr.accessPolicyWindow.base_ptr = const_cast<void *>(ptr);
        attr.accessPolicyWindow.num_bytes = win;
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)ctx->l2_persist / (double)win);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;
        attr.accessPolicyWindow.hitRatio = 0.f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    if (cudaStreamSetAttribute(ctx->compute, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

static inline unsigned kv_grid_for(const KvCtx *c, uint64_t n, int per_sm = 8)
{
    uint64_t blocks = (n + 255) / 256;
    uint64_t cap = (uint64_t)c->sm_count * per_sm;
    return (unsigned)std::max<uint64_t>(1, std::min(blocks, cap));
}

// ------------------------------------------------------------------ sketch

// a sketch whose tables span the HBM of several GPUs (see "spanning sketches" below)
struct KvSpan {
    int rank, world;
    uint64_t piece[KV_TABLES_DEV];       // bytes of table t that each rank holds (a multiple of the VMM granularity)
    uint64_t va_bytes;
    CUdeviceptr va;
    CUmemGenericAllocationHandle phys[KV_MAX_RANKS][KV_TABLES_DEV];   // one physical allocation per (rank, table): cuMemMap cannot map at an offset
    bool have[KV_MAX_RANKS];
    bool ready;
    // tiled-update exchange buffers: mine (cudaMalloc, exported over CUDA IPC) and the peers' (mapped)
    KvTileInfo ti;                       // fixed geometry: rb, run_base, cap, cursor = mine, slab = mine
    uint32_t runs;
    uint64_t chunk_pos;
    size_t smem;
    uint32_t direct_below;
    uint32_t *cursor[KV_MAX_RANKS];
    uint16_t *slab[KV_MAX_RANKS];
};

struct kv_sketch {
    int hasher, bits, ksize, n_tables, device;
    uint64_t sizes[KV_TABLES_DEV];    // buckets held here (the whole table, or this shard's bin range)
    uint64_t msizes[KV_TABLES_DEV];   // full table sizes (the primes)
    uint64_t lo[KV_TABLES_DEV];       // first bin held here
    int shard, n_shards;              // 0 of 1 for an ordinary sketch
    uint64_t nbytes[KV_TABLES_DEV];   // khmer byte length of each (local) table
    uint64_t toff[KV_TABLES_DEV];     // offset of each table in the flat allocation
    uint64_t flat_bytes;
    uint8_t *flat;
    uint32_t *state;                  // [occ bitmap of table 0 | ... | hot filter], 8/4-bit sketches only
    uint64_t soff[KV_TABLES_DEV];     // word offset of each table's occupancy bitmap
    uint64_t hot_off;                 // word offset of the hot bitmap
    uint64_t hot_base[KV_TABLES_DEV]; // first bit of each table inside the hot bitmap
    uint64_t state_words;
    bool track_unique, unique_valid;
    bool defer_unique;                // not tracked, but a single-chunk consume runs table 0's first-touch pass A so that kv_unique_last_batch need not
    bool state_stale;                 // tables were written behind the kernels' back: rebuild the hot bitmap before the next update
    uint64_t n_unique;                // host copy, updated at stats time
    unsigned long long *d_unique;     // device accumulator
    KvSpan *span;                     // non-NULL: the tables span the HBM of several GPUs (kv_sketch_create_span)
};

static uint64_t kv_table_bytes(int bits, uint64_t size)
{
    return bits == 8 ? size : (bits == 4 ? size / 2 + 1 : size / 8 + 1);
}

// device -> host copy of bytes [off, off + n) of table t; for spanning sketches one copy per piece (a
// single memcpy must not straddle physical allocations that live on different GPUs)
static cudaError_t kv_table_d2h(KvCtx *ctx, const kv_sketch *s, int t, uint64_t off, void *dst, uint64_t n);

static int kv_state_rebuild_locked(KvCtx *ctx, kv_sketch *s);
static void kv_span_free(kv_sketch *s);
static int kv_tile_rb_for(const KvCtx *ctx, const kv_sketch *s);
static int kv_tile_blk_for(const KvCtx *ctx, uint64_t runs);

static cudaError_t kv_table_d2h(KvCtx *ctx, const kv_sketch *s, int t, uint64_t off, void *dst, uint64_t n)
{
    const uint8_t *src = s->flat + s->toff[t];
    if (!s->span) return cudaMemcpyAsync(dst, src + off, n, cudaMemcpyDeviceToHost, ctx->compute);
    const uint64_t piece = s->span->piece[t];
    while (n) {
        const uint64_t room = piece - off % piece, m = std::min(n, room);
        cudaError_t e = cudaMemcpyAsync(dst, src + off, m, cudaMemcpyDeviceToHost, ctx->compute);
        if (e != cudaSuccess) return e;
        off += m; n -= m; dst = (uint8_t *)dst + m;
    }
    return cudaSuccess;
}

static KvView kv_view(const kv_sketch *s)
{
    KvView v;
    memset(&v, 0, sizeof v);
    v.n_tables = s->n_tables;
    v.bits = s->bits;
    v.hotf = s->state ? s->state + s->hot_off : nullptr;
    for (int t = 0; t < s->n_tables; t++) v.hot_base[t] = s->hot_base[t];
    for (int t = 0; t < s->n_tables; t++) {
        v.tab[t] = s->flat + s->toff[t];
        v.size[t] = s->sizes[t];
        v.msize[t] = s->msizes[t];
        v.lo[t] = s->lo[t];
        v.magic[t] = UINT64_MAX / s->msizes[t];
        v.occ[t] = s->state ? s->state + s->soff[t] : nullptr;
    }
    return v;
}

static bool is_prime_u64(uint64_t n)
{
    if (n < 2) return false;
    if (n < 4) return true;
    if (n % 2 == 0) return false;
    for (uint64_t i = 3; i * i <= n; i += 2)
        if (n % i == 0) return false;
    return true;
}

extern "C" int kv_primes_below(uint64_t x, int n, uint64_t *out)
{
    if (n < 1 || !out) return kv_fail(KV_EINVAL, "kv_primes_below: bad arguments");
    // khmer: one table "near 1" has size 1 (kevlar/tests/test_simlike.py:69 builds Nodetable(31, 1, 1);
    // the fixture term-high-abund/reference.sct is such a sketch)
    if (x == 1 && n == 1) { out[0] = 1; return KV_OK; }
    if (x < 3) return kv_fail(KV_EINVAL, "cannot find %d primes below %llu", n, (unsigned long long)x);
    uint64_t i = x - 1;
    if (i % 2 == 0) i--;
    int found = 0;
    while (found < n && i > 0) {
        if (is_prime_u64(i)) out[found++] = i;
        if (i == 1) break;
        i -= 2;
    }
    if (found != n) return kv_fail(KV_EINVAL, "cannot find %d primes below %llu", n, (unsigned long long)x);
    return KV_OK;
}

extern "C" int kv_device_count(int *n)
{
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { cudaGetLastError(); c = 0; }
    if (n) *n = c;
    return KV_OK;
}

// bins [lo, hi) of a table of `size` buckets that shard `shard` of `n_shards` holds: contiguous,
// multiples of 8 so that every shard's bytes are whole bytes for all counter widths
static void kv_shard_range(uint64_t size, int shard, int n_shards, uint64_t *lo, uint64_t *hi)
{
    uint64_t per = ((size + n_shards - 1) / n_shards + 7) & ~(uint64_t)7;
    *lo = std::min<uint64_t>(size, per * (uint64_t)shard);
    *hi = std::min<uint64_t>(size, per * (uint64_t)(shard + 1));
}

static int kv_sketch_alloc(int hasher, int bits, int ksize, int n_tables, const uint64_t *sizes, int device,
                           bool zero, kv_sketch **out, int shard = 0, int n_shards = 1)
{
    if (!out || !sizes) return kv_fail(KV_EINVAL, "null argument");
    if (hasher != KV_HASH_MURMUR && hasher != KV_HASH_TWOBIT) return kv_fail(KV_EINVAL, "unknown hasher %d", hasher);
    if (bits != 8 && bits != 4 && bits != 1) return kv_fail(KV_EINVAL, "counter width must be 8, 4 or 1 bits");
    if (n_tables < 1 || n_tables > KV_TABLES_DEV)
        return kv_fail(KV_EINVAL, "n_tables must be between 1 and %d", KV_TABLES_DEV);
    int kmax = hasher == KV_HASH_TWOBIT ? KV_MAX_KSIZE_TWOBIT : KV_MAX_KSIZE_MURMUR;
    if (ksize < 1 || ksize > kmax) return kv_fail(KV_EINVAL, "k-mer size %d not supported (1..%d for this hasher)", ksize, kmax);
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(device));
    kv_sketch *s = new kv_sketch();
    memset(s, 0, sizeof *s);
    s->hasher = hasher; s->bits = bits; s->ksize = ksize; s->n_tables = n_tables; s->device = device;
    s->shard = shard; s->n_shards = n_shards;
    uint64_t off = 0, soff = 0, hotbits = 0;
    for (int t = 0; t < n_tables; t++) {
        if (sizes[t] < 1 || sizes[t] >= (1ull << 62)) { delete s; return kv_fail(KV_EINVAL, "bad table size"); }
        uint64_t lo = 0, hi = sizes[t];
        if (n_shards > 1) kv_shard_range(sizes[t], shard, n_shards, &lo, &hi);
        s->msizes[t] = sizes[t];
        s->lo[t] = lo;
        s->sizes[t] = hi - lo;     // 0 if this shard holds nothing of the table (tiny tables, many shards)
        // this shard's bytes inside khmer's table layout; whoever holds the table's end also holds
        // the trailing byte of the nibble / bit layouts
        const uint64_t lo_bytes = lo * (uint64_t)bits / 8;
        if (hi <= lo) s->nbytes[t] = 0;
        else if (hi == sizes[t]) s->nbytes[t] = kv_table_bytes(bits, sizes[t]) - lo_bytes;
        else s->nbytes[t] = (hi - lo) * (uint64_t)bits / 8;
        s->toff[t] = off;
        off += (s->nbytes[t] + 256) & ~(uint64_t)255;
        s->soff[t] = soff;
        soff += ((s->sizes[t] + 31) / 32 + 64) & ~(uint64_t)63;
        s->hot_base[t] = hotbits;
        hotbits += ((s->sizes[t] >> 3) + 1 + 2047) & ~(uint64_t)2047;
    }
    s->hot_off = soff;
    s->state_words = bits == 1 ? 0 : soff + hotbits / 32;
    s->flat_bytes = off;
    cudaError_t e = cudaMalloc((void **)&s->flat, off);
    if (e != cudaSuccess) {
        delete s;
        cudaGetLastError();
        return kv_fail(KV_ENOMEM, "cannot allocate %llu bytes of HBM for the sketch: %s", (unsigned long long)off,
                       cudaGetErrorString(e));
    }
    if (zero) CU(cudaMemsetAsync(s->flat, 0, off, ctx->compute));
    if (s->state_words) {
        e = cudaMalloc((void **)&s->state, s->state_words * 4);
        if (e != cudaSuccess) {
            cudaFree(s->flat);
            delete s;
            cudaGetLastError();
            return kv_fail(KV_ENOMEM, "cannot allocate the bucket-state array: %s", cudaGetErrorString(e));
        }
        CU(cudaMemsetAsync(s->state, 0, s->state_words * 4, ctx->compute));
    }
    CU(cudaMalloc((void **)&s->d_unique, sizeof(unsigned long long)));
    CU(cudaMemsetAsync(s->d_unique, 0, sizeof(unsigned long long), ctx->compute));
    s->track_unique = n_shards == 1;   // the order-dependent statistic needs all buckets of a k-mer in one place
    s->unique_valid = n_shards == 1;
    *out = s;
    return KV_OK;
}

extern "C" int kv_sketch_create(int hasher, int bits, int ksize, int n_tables, const uint64_t *sizes, int device,
                                kv_sketch **out)
{
    return kv_sketch_alloc(hasher, bits, ksize, n_tables, sizes, device, true, out);
}

extern "C" int kv_sketch_create_shard(int hasher, int bits, int ksize, int n_tables, const uint64_t *sizes, int shard,
                                      int n_shards, int device, kv_sketch **out)
{
    if (n_shards < 1 || shard < 0 || shard >= n_shards) return kv_fail(KV_EINVAL, "shard %d of %d", shard, n_shards);
    return kv_sketch_alloc(hasher, bits, ksize, n_tables, sizes, device, true, out, shard, n_shards);
}

extern "C" int kv_sketch_shard_info(const kv_sketch *s, int *shard, int *n_shards, uint64_t *lo, uint64_t *count)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    if (shard) *shard = s->shard;
    if (n_shards) *n_shards = s->n_shards;
    for (int t = 0; t < s->n_tables; t++) {
        if (lo) lo[t] = s->lo[t];
        if (count) count[t] = s->sizes[t];
    }
    return KV_OK;
}

extern "C" int kv_sketch_destroy(kv_sketch *s)
{
    if (!s) return KV_OK;
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(ctx->compute));
    CU(cudaStreamSynchronize(ctx->merge));
    if (ctx->last_hashed.sketch == s) ctx->last_hashed.sketch = nullptr;
    if (s->span) kv_span_free(s);
    if (s->flat) cudaFree(s->flat);
    if (s->state) cudaFree(s->state);
    if (s->d_unique) cudaFree(s->d_unique);
    delete s;
    return KV_OK;
}

extern "C" int kv_sketch_clear(kv_sketch *s)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    if (s->span) {   // every rank clears the pieces it holds (collective by convention)
        for (int t = 0; t < s->n_tables; t++)
            CU(cudaMemsetAsync(s->flat + s->toff[t] + (uint64_t)s->span->rank * s->span->piece[t], 0, s->span->piece[t], ctx->compute));
        return KV_OK;
    }
    CU(cudaMemsetAsync(s->flat, 0, s->flat_bytes, ctx->compute));
    if (s->state) CU(cudaMemsetAsync(s->state, 0, s->state_words * 4, ctx->compute));
    CU(cudaMemsetAsync(s->d_unique, 0, sizeof(unsigned long long), ctx->compute));
    s->state_stale = false;   // all counters zero, hot bitmap zero
    s->unique_valid = true;
    s->n_unique = 0;
    return KV_OK;
}

extern "C" int kv_sketch_info(const kv_sketch *s, int *hasher, int *bits, int *ksize, int *n_tables, uint64_t *sizes,
                              int *device)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    if (hasher) *hasher = s->hasher;
    if (bits) *bits = s->bits;
    if (ksize) *ksize = s->ksize;
    if (n_tables) *n_tables = s->n_tables;
    if (device) *device = s->device;
    if (sizes) for (int t = 0; t < s->n_tables; t++) sizes[t] = s->msizes[t];
    return KV_OK;
}

extern "C" int kv_sketch_set_unique_tracking(kv_sketch *s, int on)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    s->track_unique = on == 1;
    s->defer_unique = on == 2;
    return KV_OK;
}

extern "C" int kv_sketch_table(kv_sketch *s, int t, void **dev_ptr, uint64_t *nbytes)
{
    if (!s || t < 0 || t >= s->n_tables) return kv_fail(KV_EINVAL, "bad table index");
    if (dev_ptr) *dev_ptr = s->flat + s->toff[t];
    if (nbytes) *nbytes = s->nbytes[t];
    return KV_OK;
}

extern "C" int kv_sketch_flat(kv_sketch *s, void **dev_ptr, uint64_t *nbytes)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    if (dev_ptr) *dev_ptr = s->flat;
    if (nbytes) *nbytes = s->flat_bytes;
    return KV_OK;
}

extern "C" int kv_sketch_read_table(kv_sketch *s, int t, uint8_t *host_out, uint64_t nbytes)
{
    if (!s || t < 0 || t >= s->n_tables || !host_out) return kv_fail(KV_EINVAL, "bad arguments");
    if (nbytes != s->nbytes[t]) return kv_fail(KV_EINVAL, "table %d holds %llu bytes", t, (unsigned long long)s->nbytes[t]);
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    CU(kv_table_d2h(ctx, s, t, 0, host_out, nbytes));
    CU(cudaStreamSynchronize(ctx->compute));
    return KV_OK;
}

extern "C" int kv_sketch_write_table(kv_sketch *s, int t, const uint8_t *host_in, uint64_t nbytes)
{
    if (!s || t < 0 || t >= s->n_tables || !host_in) return kv_fail(KV_EINVAL, "bad arguments");
    if (nbytes != s->nbytes[t]) return kv_fail(KV_EINVAL, "table %d holds %llu bytes", t, (unsigned long long)s->nbytes[t]);
    if (s->span) return kv_fail(KV_EINVAL, "raw table writes are not supported on spanning sketches");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(s->flat + s->toff[t], host_in, nbytes, cudaMemcpyHostToDevice, ctx->compute));
    s->state_stale = true;
    CU(cudaStreamSynchronize(ctx->compute));
    s->unique_valid = false;
    return KV_OK;
}

// bucket states from the counters (after anything that wrote the tables behind the kernels' back)
static int kv_state_rebuild_locked(KvCtx *ctx, kv_sketch *s)
{
    if (!s->state) return KV_OK;
    KvView v = kv_view(s);
    CU(cudaMemsetAsync(s->state, 0, s->state_words * 4, ctx->compute));
    for (int t = 0; t < s->n_tables; t++) {
        if (s->bits == 8) LAUNCH(ctx, kv_state_rebuild_kernel<8>, kv_grid_for(ctx, s->sizes[t], 16), 256, v, t);
        else LAUNCH(ctx, kv_state_rebuild_kernel<4>, kv_grid_for(ctx, s->sizes[t], 16), 256, v, t);
    }
    return KV_OK;
}

static int kv_occupied_locked(KvCtx *ctx, kv_sketch *s, uint64_t *out)
{
    CU(cudaMemsetAsync(ctx->counters + 3, 0, sizeof(unsigned long long), ctx->compute));
    LAUNCH(ctx, kv_occupied_kernel, kv_grid_for(ctx, s->sizes[0]), 256, kv_view(s), ctx->counters + 3);
    CU(cudaMemcpyAsync(ctx->h_counters + 3, ctx->counters + 3, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    *out = ctx->h_counters[3];
    return KV_OK;
}

extern "C" int kv_sketch_stats(kv_sketch *s, uint64_t *n_occupied, uint64_t *n_unique, int *n_unique_valid)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    if (n_occupied) KV_TRY(kv_occupied_locked(ctx, s, n_occupied));
    if (n_unique || n_unique_valid) {
        CU(cudaMemcpyAsync(ctx->h_counters + 1, s->d_unique, 8, cudaMemcpyDeviceToHost, ctx->compute));
        CU(cudaStreamSynchronize(ctx->compute));
        s->n_unique = ctx->h_counters[1];
        if (n_unique) *n_unique = s->n_unique;
        if (n_unique_valid) *n_unique_valid = s->unique_valid ? 1 : 0;
    }
    return KV_OK;
}

// ------------------------------------------------------------------ spanning sketches (SURVEY 8e plan B)
//
// A sketch larger than one GPU: every table lives in ONE virtual address range that is mapped on every
// rank, piece r of each table backed by physical HBM of rank r (CUDA virtual memory management; the
// pieces are shared between the processes as POSIX file descriptors).  Kernels see an ordinary KvView --
// a counter load or atomic lands in whichever GPU's HBM holds the page, over NVLink when it is a peer's --
// so get / novel / save / stats need no sharded variants.  Updates go through the tiled path with the
// exchange fused into the apply kernel: every rank files the updates of ITS reads under (table, region)
// in its own slabs, and the apply kernel of the rank that holds a region pulls that region's slab from
// every rank over NVLink (2 bytes per update) and applies it in shared memory.

struct KvDrv {
    bool ok = false;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemExport)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImport)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*MemGetGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
};

static KvDrv g_drv;

static int kv_drv_load()
{
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (g_drv.ok) return KV_OK;
    struct { const char *name; void **fn; } want[] = {
        {"cuMemCreate", (void **)&g_drv.MemCreate}, {"cuMemRelease", (void **)&g_drv.MemRelease},
        {"cuMemExportToShareableHandle", (void **)&g_drv.MemExport}, {"cuMemImportFromShareableHandle", (void **)&g_drv.MemImport},
        {"cuMemAddressReserve", (void **)&g_drv.MemAddressReserve}, {"cuMemAddressFree", (void **)&g_drv.MemAddressFree},
        {"cuMemMap", (void **)&g_drv.MemMap}, {"cuMemUnmap", (void **)&g_drv.MemUnmap},
        {"cuMemSetAccess", (void **)&g_drv.MemSetAccess}, {"cuMemGetAllocationGranularity", (void **)&g_drv.MemGetGranularity},
    };
    for (auto &w : want) {
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint(w.name, w.fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !*w.fn) {
            cudaGetLastError();
            return kv_fail(KV_ECUDA, "CUDA driver entry point %s is not available", w.name);
        }
    }
    g_drv.ok = true;
    return KV_OK;
}

#define CUD(expr)                                                                                         \
    do {                                                                                                  \
        CUresult r_ = (expr);                                                                             \
        if (r_ != CUDA_SUCCESS) return kv_fail(KV_ECUDA, "%s failed with CUresult %d (%s:%d)", #expr, (int)r_, __FILE__, __LINE__); \
    } while (0)


static CUmemAllocationProp kv_span_prop(int device)
{
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

// map rank r's physical allocations: piece r of every table
static int kv_span_map(kv_sketch *s, int r)
{
    KvSpan *sp = s->span;
    for (int t = 0; t < s->n_tables; t++)
        CUD(g_drv.MemMap(sp->va + s->toff[t] + (uint64_t)r * sp->piece[t], sp->piece[t], 0, sp->phys[r][t], 0));
    return KV_OK;
}

extern "C" int kv_sketch_create_span(int hasher, int bits, int ksize, int n_tables, const uint64_t *sizes, int rank, int world,
                                     int device, uint64_t chunk_positions, kv_sketch **out, int *mem_fds_out,
                                     uint8_t xchg_handle_out[128])
{
    if (!out || !sizes || !mem_fds_out || !xchg_handle_out) return kv_fail(KV_EINVAL, "null argument");
    if (hasher != KV_HASH_MURMUR && hasher != KV_HASH_TWOBIT) return kv_fail(KV_EINVAL, "unknown hasher %d", hasher);
    if (bits != 8 && bits != 4 && bits != 1) return kv_fail(KV_EINVAL, "counter width must be 8, 4 or 1 bits");
    if (n_tables < 1 || n_tables > KV_TABLES_DEV) return kv_fail(KV_EINVAL, "n_tables must be between 1 and %d", KV_TABLES_DEV);
    if (world < 1 || world > KV_MAX_RANKS || rank < 0 || rank >= world) return kv_fail(KV_EINVAL, "rank %d of %d", rank, world);
    int kmax = hasher == KV_HASH_TWOBIT ? KV_MAX_KSIZE_TWOBIT : KV_MAX_KSIZE_MURMUR;
    if (ksize < 1 || ksize > kmax) return kv_fail(KV_EINVAL, "k-mer size %d not supported (1..%d for this hasher)", ksize, kmax);
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    KV_TRY(kv_drv_load());
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(device));
    CU(cudaFree(0));   // make sure the primary context is current for the driver calls
    CUmemAllocationProp prop = kv_span_prop(device);
    size_t gran = 0;
    CUD(g_drv.MemGetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    if (gran < (2u << 20)) gran = 2u << 20;
    kv_sketch *s = new kv_sketch();
    memset(s, 0, sizeof *s);
    KvSpan *sp = new KvSpan();
    memset(sp, 0, sizeof *sp);
    s->span = sp;
    s->hasher = hasher; s->bits = bits; s->ksize = ksize; s->n_tables = n_tables; s->device = device;
    s->shard = 0; s->n_shards = 1;
    sp->rank = rank; sp->world = world;
    uint64_t va_off = 0;
    for (int t = 0; t < n_tables; t++) {
        if (sizes[t] < 1 || sizes[t] >= (1ull << 62)) { delete sp; delete s; return kv_fail(KV_EINVAL, "bad table size"); }
        s->sizes[t] = s->msizes[t] = sizes[t];
        s->lo[t] = 0;
        s->nbytes[t] = kv_table_bytes(bits, sizes[t]);
        const uint64_t per = (s->nbytes[t] + world - 1) / world;
        sp->piece[t] = (per + gran - 1) / gran * gran;
        s->toff[t] = va_off;
        va_off += sp->piece[t] * (uint64_t)world;
    }
    sp->va_bytes = va_off;
    s->flat_bytes = va_off;
    for (int t = 0; t < n_tables; t++) {
        CUresult r = g_drv.MemCreate(&sp->phys[rank][t], sp->piece[t], &prop, 0);
        if (r != CUDA_SUCCESS)
            return kv_fail(r == CUDA_ERROR_OUT_OF_MEMORY ? KV_ENOMEM : KV_ECUDA, "cuMemCreate(%llu bytes) failed with CUresult %d",
                           (unsigned long long)sp->piece[t], (int)r);
    }
    sp->have[rank] = true;
    CUD(g_drv.MemAddressReserve(&sp->va, sp->va_bytes, gran, 0, 0));
    s->flat = (uint8_t *)sp->va;
    KV_TRY(kv_span_map(s, rank));
    {   // my pieces are accessible right away (zero-filled below); the peers' after kv_sketch_span_ready
        CUmemAccessDesc acc;
        memset(&acc, 0, sizeof acc);
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        for (int t = 0; t < n_tables; t++)
            CUD(g_drv.MemSetAccess(sp->va + s->toff[t] + (uint64_t)rank * sp->piece[t], sp->piece[t], &acc, 1));
    }
    for (int t = 0; t < n_tables; t++)
        CU(cudaMemsetAsync(s->flat + s->toff[t] + (uint64_t)rank * sp->piece[t], 0, sp->piece[t], ctx->compute));
    for (int t = 0; t < n_tables; t++) {
        int fd = -1;
        CUD(g_drv.MemExport(&fd, sp->phys[rank][t], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
        mem_fds_out[t] = fd;
    }
    // fixed geometry of the tiled update (same on every rank)
    const int rb = kv_tile_rb_for(ctx, s);
    uint64_t runs = 0, min_regions = UINT64_MAX;
    for (int t = 0; t < n_tables; t++) {
        sp->ti.run_base[t] = (uint32_t)runs;
        const uint64_t nr = ((sizes[t] - 1) >> rb) + 1;
        runs += nr;
        min_regions = std::min(min_regions, nr);
    }
    if (runs >= (1ull << 30)) return kv_fail(KV_EINVAL, "too many regions");
    sp->ti.run_base[n_tables] = sp->ti.run_base[KV_TABLES_DEV] = (uint32_t)runs;
    sp->runs = (uint32_t)runs;
    sp->chunk_pos = std::max<uint64_t>(KV_TILE, (chunk_positions ? chunk_positions : ctx->tile_chunk_bases) / KV_TILE * KV_TILE);
    const double mean = (double)sp->chunk_pos / (double)min_regions;
    const int blk_log2 = std::max(3, kv_tile_blk_for(ctx, runs));
    const uint64_t blk = 1ull << blk_log2;
    uint64_t cap = (uint64_t)(mean * 1.0625 + 8.0 * sqrt(mean) + 64.0);
    cap = (cap + blk - 1) / blk * blk;
    sp->ti.rb = rb; sp->ti.cap = (uint32_t)cap; sp->ti.blk_log2 = blk_log2;
    sp->smem = bits == 1 ? ((size_t)1 << rb) / 8 : ((size_t)1 << rb) * 2;
    const uint64_t region_bytes = ((uint64_t)1 << rb) * (uint64_t)bits / 8;
    sp->direct_below = ctx->tile_direct_below >= 0 ? (uint32_t)ctx->tile_direct_below : (uint32_t)std::max<uint64_t>(1, region_bytes / 64);
    CU(cudaMalloc((void **)&sp->cursor[rank], runs * 4 + 16));
    CU(cudaMalloc((void **)&sp->slab[rank], runs * cap * 2));
    sp->ti.cursor = sp->cursor[rank];
    sp->ti.slab = sp->slab[rank];
    sp->ti.ovf_any = (unsigned *)(sp->cursor[rank] + runs);
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, sp->cursor[rank]));
    memcpy(xchg_handle_out, &h, 64);
    CU(cudaIpcGetMemHandle(&h, sp->slab[rank]));
    memcpy(xchg_handle_out + 64, &h, 64);
    CU(cudaMalloc((void **)&s->d_unique, sizeof(unsigned long long)));
    CU(cudaMemsetAsync(s->d_unique, 0, sizeof(unsigned long long), ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    s->track_unique = false;   // the order-dependent statistic is defined for one stream of reads
    s->unique_valid = false;
    *out = s;
    return KV_OK;
}

extern "C" int kv_sketch_span_attach(kv_sketch *s, int peer_rank, const int *mem_fds, const uint8_t xchg_handle[128])
{
    if (!s || !s->span || !xchg_handle || !mem_fds) return kv_fail(KV_EINVAL, "not a spanning sketch");
    KvSpan *sp = s->span;
    if (peer_rank < 0 || peer_rank >= sp->world || peer_rank == sp->rank || sp->have[peer_rank])
        return kv_fail(KV_EINVAL, "bad or repeated peer rank %d", peer_rank);
    CU(cudaSetDevice(s->device));
    for (int t = 0; t < s->n_tables; t++)
        CUD(g_drv.MemImport(&sp->phys[peer_rank][t], (void *)(uintptr_t)mem_fds[t], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    sp->have[peer_rank] = true;
    KV_TRY(kv_span_map(s, peer_rank));
    cudaIpcMemHandle_t h;
    memcpy(&h, xchg_handle, 64);
    CU(cudaIpcOpenMemHandle((void **)&sp->cursor[peer_rank], h, cudaIpcMemLazyEnablePeerAccess));
    memcpy(&h, xchg_handle + 64, 64);
    CU(cudaIpcOpenMemHandle((void **)&sp->slab[peer_rank], h, cudaIpcMemLazyEnablePeerAccess));
    return KV_OK;
}

extern "C" int kv_sketch_span_ready(kv_sketch *s)
{
    if (!s || !s->span) return kv_fail(KV_EINVAL, "not a spanning sketch");
    KvSpan *sp = s->span;
    for (int r = 0; r < sp->world; r++)
        if (!sp->have[r]) return kv_fail(KV_EINVAL, "rank %d is not attached yet", r);
    CU(cudaSetDevice(s->device));
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = s->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CUD(g_drv.MemSetAccess(sp->va, sp->va_bytes, &acc, 1));
    sp->ready = true;
    return KV_OK;
}

extern "C" int kv_sketch_span_info(const kv_sketch *s, int *rank, int *world, uint64_t *piece_bytes, uint64_t *chunk_positions)
{
    if (!s || !s->span) return kv_fail(KV_EINVAL, "not a spanning sketch");
    if (rank) *rank = s->span->rank;
    if (world) *world = s->span->world;
    if (chunk_positions) *chunk_positions = s->span->chunk_pos;
    if (piece_bytes) for (int t = 0; t < s->n_tables; t++) piece_bytes[t] = s->span->piece[t];
    return KV_OK;
}

// unmap + release everything that is not plain cudaMalloc memory (called by kv_sketch_destroy)
static void kv_span_free(kv_sketch *s)
{
    KvSpan *sp = s->span;
    if (!sp) return;
    for (int r = 0; r < sp->world; r++) {
        if (r != sp->rank) {
            if (sp->cursor[r]) cudaIpcCloseMemHandle(sp->cursor[r]);
            if (sp->slab[r]) cudaIpcCloseMemHandle(sp->slab[r]);
        }
        if (sp->have[r] && g_drv.ok) {
            for (int t = 0; t < s->n_tables; t++) {
                g_drv.MemUnmap(sp->va + s->toff[t] + (uint64_t)r * sp->piece[t], sp->piece[t]);
                g_drv.MemRelease(sp->phys[r][t]);
            }
        }
    }
    if (sp->va && g_drv.ok) g_drv.MemAddressFree(sp->va, sp->va_bytes);
    if (sp->cursor[sp->rank]) cudaFree(sp->cursor[sp->rank]);
    if (sp->slab[sp->rank]) cudaFree(sp->slab[sp->rank]);
    delete sp;
    s->span = nullptr;
    s->flat = nullptr;
}

// ------------------------------------------------------------------ OXLI v4 I/O

extern "C" int kv_sketch_save(kv_sketch *s, const char *path)
{
    if (!s || !path) return kv_fail(KV_EINVAL, "null argument");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    uint64_t occ = 0;
    KV_TRY(kv_occupied_locked(ctx, s, &occ));
    int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) return kv_fail(KV_EIO, "cannot open %s for writing", path);
    // header: "OXLI", version, type, [use_bigcount], k, n_tables, n_occupied; then per table its size and its bytes
    uint8_t head[32];
    size_t hn = 0;
    auto put = [&](const void *p, size_t n) { memcpy(head + hn, p, n); hn += n; };
    const uint8_t version = 4, type = s->bits == 8 ? 1 : (s->bits == 4 ? 7 : 2), big = 0, nt = (uint8_t)s->n_tables;
    const uint32_t k = (uint32_t)s->ksize;
    put("OXLI", 4); put(&version, 1); put(&type, 1);
    if (s->bits == 8) put(&big, 1);
    put(&k, 4); put(&nt, 1); put(&occ, 8);
    uint64_t pos = 0;
    int rc = KV_OK;
    auto put_small = [&](const void *p, size_t n) {
        if (rc == KV_OK && pwrite(fd, p, n, (off_t)pos) != (ssize_t)n) rc = kv_fail(KV_EIO, "short write to %s", path);
        pos += n;
    };
    put_small(head, hn);
    uint8_t *stage = nullptr;
    if (rc == KV_OK) rc = kv_io_stage(ctx, &stage);
    const size_t CH = KV_IO_STAGE / 2;
    KvFileJob job[2];
    int half = 0;
    for (int t = 0; t < s->n_tables && rc == KV_OK; t++) {
        put_small(&s->sizes[t], 8);
        for (uint64_t o = 0; o < s->nbytes[t] && rc == KV_OK; o += CH, half ^= 1) {
            const size_t n = (size_t)std::min<uint64_t>(CH, s->nbytes[t] - o);
            uint8_t *buf = stage + (size_t)half * CH;
            if (!job[half].join()) { rc = kv_fail(KV_EIO, "write error on %s", path); break; }   // this half is free again
            if (kv_table_d2h(ctx, s, t, o, buf, n) != cudaSuccess ||
                cudaStreamSynchronize(ctx->compute) != cudaSuccess) { rc = kv_fail(KV_ECUDA, "D2H copy failed"); break; }
            job[half].start(fd, buf, n, pos + o, true);
        }
        pos += s->nbytes[t];
    }
    for (int h = 0; h < 2; h++)
        if (!job[h].join() && rc == KV_OK) rc = kv_fail(KV_EIO, "write error on %s", path);
    if (s->bits == 8) { const uint64_t nbig = 0; put_small(&nbig, 8); }
    if (close(fd) != 0 && rc == KV_OK) rc = kv_fail(KV_EIO, "write error on %s", path);
    return rc;
}

// One piece of an OXLI v4 file written cooperatively by the shards of a sketch (the caller orders the
// calls: header by shard 0, then per table its size field by shard 0 and the shards' bytes in shard
// order, finally the trailer by shard 0).  piece: 0 header (creates the file), 1 size field of table t,
// 2 this shard's bytes of table t, 3 trailer.
extern "C" int kv_sketch_save_part(kv_sketch *s, const char *path, int piece, int t, uint64_t n_occupied)
{
    if (!s || !path) return kv_fail(KV_EINVAL, "null argument");
    if ((piece == 1 || piece == 2) && (t < 0 || t >= s->n_tables)) return kv_fail(KV_EINVAL, "bad table index");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    FILE *f = fopen(path, piece == 0 ? "wb" : "ab");
    if (!f) return kv_fail(KV_EIO, "cannot open %s for writing", path);
    int rc = KV_OK;
    if (piece == 0) {
        uint8_t version = 4, type = s->bits == 8 ? 1 : (s->bits == 4 ? 7 : 2);
        fwrite("OXLI", 1, 4, f);
        fwrite(&version, 1, 1, f);
        fwrite(&type, 1, 1, f);
        if (s->bits == 8) { uint8_t big = 0; fwrite(&big, 1, 1, f); }
        uint32_t k = (uint32_t)s->ksize;
        uint8_t nt = (uint8_t)s->n_tables;
        fwrite(&k, 4, 1, f);
        fwrite(&nt, 1, 1, f);
        fwrite(&n_occupied, 8, 1, f);
    } else if (piece == 1) {
        fwrite(&s->msizes[t], 8, 1, f);
    } else if (piece == 2) {
        uint8_t *stage = nullptr;
        rc = kv_io_stage(ctx, &stage);
        for (uint64_t o = 0; o < s->nbytes[t] && rc == KV_OK; o += KV_IO_STAGE) {
            size_t n = (size_t)std::min<uint64_t>(KV_IO_STAGE, s->nbytes[t] - o);
            if (cudaMemcpyAsync(stage, s->flat + s->toff[t] + o, n, cudaMemcpyDeviceToHost, ctx->compute) != cudaSuccess ||
                cudaStreamSynchronize(ctx->compute) != cudaSuccess) { rc = kv_fail(KV_ECUDA, "D2H copy failed"); break; }
            if (fwrite(stage, 1, n, f) != n) rc = kv_fail(KV_EIO, "short write to %s", path);
        }
    } else if (piece == 3) {
        if (s->bits == 8) { uint64_t nbig = 0; fwrite(&nbig, 8, 1, f); }
    } else
        rc = kv_fail(KV_EINVAL, "unknown piece %d", piece);
    if (rc == KV_OK && ferror(f)) rc = kv_fail(KV_EIO, "write error on %s", path);
    fclose(f);
    return rc;
}

extern "C" int kv_sketch_load(const char *path, int hasher, int expect_bits, int device, kv_sketch **out)
{
    if (!path || !out) return kv_fail(KV_EINVAL, "null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return kv_fail(KV_EIO, "cannot open %s", path);
    uint8_t head[6];
    if (fread(head, 1, 6, f) != 6 || memcmp(head, "OXLI", 4)) { fclose(f); return kv_fail(KV_EIO, "%s: not an OXLI sketch file", path); }
    if (head[4] != 4) { fclose(f); return kv_fail(KV_EIO, "%s: unsupported OXLI version %d", path, head[4]); }
    int bits = head[5] == 1 ? 8 : head[5] == 7 ? 4 : head[5] == 2 ? 1 : 0;
    if (!bits || (expect_bits && bits != expect_bits)) { fclose(f); return kv_fail(KV_EIO, "%s: unexpected table type %d", path, head[5]); }
    uint8_t big = 0, nt = 0;
    uint32_t k = 0;
    uint64_t occ = 0;
    bool ok = true;
    if (bits == 8) ok = fread(&big, 1, 1, f) == 1;
    ok = ok && fread(&k, 4, 1, f) == 1 && fread(&nt, 1, 1, f) == 1 && fread(&occ, 8, 1, f) == 1;
    if (!ok || nt < 1 || nt > KV_TABLES_DEV) { fclose(f); return kv_fail(KV_EIO, "%s: truncated or unsupported header", path); }
    // first pass: table sizes (they are interleaved with the data)
    uint64_t sizes[KV_TABLES_DEV];
    long data_pos[KV_TABLES_DEV];
    for (int t = 0; t < nt; t++) {
        if (fread(&sizes[t], 8, 1, f) != 1) { fclose(f); return kv_fail(KV_EIO, "%s: truncated file", path); }
        data_pos[t] = ftell(f);
        if (fseek(f, (long)kv_table_bytes(bits, sizes[t]), SEEK_CUR)) { fclose(f); return kv_fail(KV_EIO, "%s: truncated file", path); }
    }
    kv_sketch *s = nullptr;
    int rc = kv_sketch_alloc(hasher, bits, (int)k, nt, sizes, device, true, &s);
    if (rc != KV_OK) { fclose(f); return rc; }
    KvCtx *ctx;
    kv_ctx_get(device, &ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    const size_t CH = KV_IO_STAGE / 2;
    uint8_t *stage = nullptr;
    rc = kv_io_stage(ctx, &stage);
    {   // chunk i+1 is read from the file (several threads) while chunk i travels to the device
        struct Chunk { int t; uint64_t o; size_t n; };
        std::vector<Chunk> chunks;
        for (int t = 0; t < nt; t++)
            for (uint64_t o = 0; o < s->nbytes[t]; o += CH) chunks.push_back({t, o, (size_t)std::min<uint64_t>(CH, s->nbytes[t] - o)});
        KvFileJob job[2];
        cudaEvent_t sent[2] = {nullptr, nullptr};
        const int fd = fileno(f);
        for (size_t i = 0; i <= chunks.size() && rc == KV_OK; i++) {
            if (i < chunks.size()) {
                const int h = (int)(i & 1);
                if (sent[h] && cudaEventSynchronize(sent[h]) != cudaSuccess) { rc = kv_fail(KV_ECUDA, "H2D copy failed"); break; }
                job[h].start(fd, stage + (size_t)h * CH, chunks[i].n, (uint64_t)data_pos[chunks[i].t] + chunks[i].o, false);
            }
            if (i > 0) {
                const int h = (int)((i - 1) & 1);
                const Chunk &c = chunks[i - 1];
                if (!job[h].join()) { rc = kv_fail(KV_EIO, "%s: truncated file", path); break; }
                if (!sent[h] && cudaEventCreateWithFlags(&sent[h], cudaEventDisableTiming) != cudaSuccess) { rc = kv_fail(KV_ECUDA, "cannot create an event"); break; }
                if (cudaMemcpyAsync(s->flat + s->toff[c.t] + c.o, stage + (size_t)h * CH, c.n, cudaMemcpyHostToDevice, ctx->compute) != cudaSuccess ||
                    cudaEventRecord(sent[h], ctx->compute) != cudaSuccess) { rc = kv_fail(KV_ECUDA, "H2D copy failed"); break; }
            }
        }
        job[0].join(); job[1].join();
        if (cudaStreamSynchronize(ctx->compute) != cudaSuccess && rc == KV_OK) rc = kv_fail(KV_ECUDA, "H2D copy failed");
        for (int h = 0; h < 2; h++)
            if (sent[h]) cudaEventDestroy(sent[h]);
    }
    if (rc == KV_OK && bits == 8 && big) {
        // khmer's use_bigcount: counts above 255 live in a map behind the tables; kevlar never sets it
        // (Counttable has no bigcount) -- refuse rather than load saturated 255s in their place
        uint64_t n_big = 0;
        if (fread(&n_big, 8, 1, f) == 1 && n_big)
            rc = kv_fail(KV_EIO, "%s: the file carries %llu bigcount entries (khmer use_bigcount), which this sketch type does not hold",
                         path, (unsigned long long)n_big);
    }
    fclose(f);
    if (rc != KV_OK) { cudaFree(s->flat); cudaFree(s->state); cudaFree(s->d_unique); delete s; return rc; }
    s->state_stale = true;
    (void)occ;   // recomputed from table 0 whenever it is asked for
    *out = s;
    return KV_OK;
}

// ------------------------------------------------------------------ batch staging

struct KvBatch {
    const uint8_t *d_bases;
    const uint64_t *d_offsets;
    uint64_t total, n_reads, n_tiles;
    KvSlot *slot;   // non-null when staged from host memory
};

// Make the batch visible on the device: copy host buffers into a staging slot on the copy
// stream (compute waits on the event), or use device pointers in place.  Also builds the
// tile -> first read index.
static int kv_stage(KvCtx *ctx, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads, int where,
                    uint64_t total_hint, KvBatch *b)
{
    if (n_reads >= 0xffffffffull) return kv_fail(KV_EINVAL, "at most 2^32-2 reads per batch");
    b->n_reads = n_reads;
    b->slot = nullptr;
    uint64_t total = total_hint;
    if (where == KV_MEM_HOST) {
        total = offsets[n_reads];
        for (uint64_t r = 0; r < n_reads; r++)   // cheap sanity: offsets must be non-decreasing
            if (offsets[r + 1] < offsets[r]) return kv_fail(KV_EINVAL, "read offsets must be non-decreasing");
        if (offsets[0] != 0) return kv_fail(KV_EINVAL, "offsets[0] must be 0");
        KvSlot *s = &ctx->slot[ctx->next_slot];
        ctx->next_slot ^= 1;
        if (s->used) CU(cudaEventSynchronize(s->done));
        KV_TRY(kv_buf_ensure(s->bases, total + 16));
        KV_TRY(kv_buf_ensure(s->offsets, (n_reads + 1) * 8));
        if (total) CU(cudaMemcpyAsync(s->bases.p, bases, total, cudaMemcpyHostToDevice, ctx->copy));
        CU(cudaMemcpyAsync(s->offsets.p, offsets, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->copy));
        CU(cudaEventRecord(s->copied, ctx->copy));
        CU(cudaStreamWaitEvent(ctx->compute, s->copied, 0));
        s->used = true;
        b->slot = s;
        b->d_bases = (const uint8_t *)s->bases.p;
        b->d_offsets = (const uint64_t *)s->offsets.p;
    } else if (where == KV_MEM_DEVICE) {
        if (((uintptr_t)bases & 3) || ((uintptr_t)offsets & 7)) return kv_fail(KV_EINVAL, "device batch pointers must be 4/8-byte aligned");
        // the batch is caller-owned input, not produced on the compute stream: fetch the total on
        // the copy stream so queued kernels keep running
        CU(cudaMemcpyAsync(ctx->h_counters + 4, offsets + n_reads, 8, cudaMemcpyDeviceToHost, ctx->copy));
        CU(cudaStreamSynchronize(ctx->copy));
        total = ctx->h_counters[4];
        b->d_bases = bases;
        b->d_offsets = offsets;
    } else
        return kv_fail(KV_EINVAL, "where must be KV_MEM_HOST or KV_MEM_DEVICE");
    b->total = total;
    b->n_tiles = (total + KV_TILE - 1) / KV_TILE;
    if (total && n_reads) {
        KV_TRY(kv_buf_ensure(ctx->tile_first, (b->n_tiles + 1) * 4));
        LAUNCH(ctx, kv_tile_index_kernel, (unsigned)((b->n_tiles + 1 + 255) / 256), 256, b->d_offsets, n_reads, b->n_tiles,
               (uint32_t *)ctx->tile_first.p);
    }
    return KV_OK;
}

static inline void kv_stage_done(KvCtx *ctx, KvBatch *b)
{
    if (b->slot) cudaEventRecord(b->slot->done, ctx->compute);
}

template <int HASHER, bool SCATTER>
static int kv_launch_hash2(KvCtx *ctx, const KvHashParams &p, unsigned n_tiles)
{
    if (HASHER == KV_HASH_TWOBIT) { LAUNCH_C(KV_PROF_HASH, ctx, (kv_hash_kernel<KV_HASH_TWOBIT, 4, SCATTER>), n_tiles, KV_THREADS, p); return KV_OK; }
    int kw = 4 * ((p.k + 15) / 16);
    switch (kw) {
    case 4: LAUNCH_C(KV_PROF_HASH, ctx, (kv_hash_kernel<KV_HASH_MURMUR, 4, SCATTER>), n_tiles, KV_THREADS, p); break;
    case 8: LAUNCH_C(KV_PROF_HASH, ctx, (kv_hash_kernel<KV_HASH_MURMUR, 8, SCATTER>), n_tiles, KV_THREADS, p); break;
    case 12: LAUNCH_C(KV_PROF_HASH, ctx, (kv_hash_kernel<KV_HASH_MURMUR, 12, SCATTER>), n_tiles, KV_THREADS, p); break;
    default: LAUNCH_C(KV_PROF_HASH, ctx, (kv_hash_kernel<KV_HASH_MURMUR, 16, SCATTER>), n_tiles, KV_THREADS, p); break;
    }
    return KV_OK;
}

template <int HASHER>
static int kv_launch_hash(KvCtx *ctx, const KvHashParams &p, unsigned n_tiles)
{
    return p.scatter ? kv_launch_hash2<HASHER, true>(ctx, p, n_tiles) : kv_launch_hash2<HASHER, false>(ctx, p, n_tiles);
}

static int kv_band_interval(int num_bands, int band, uint64_t *lo, uint64_t *hi)
{
    if (num_bands <= 0 || band < 0 || band >= num_bands)
        return kv_fail(KV_EINVAL, "Band number must be less than number of bands");
    uint64_t size = UINT64_MAX / (uint64_t)num_bands;
    *lo = size * (uint64_t)band;
    *hi = size * (uint64_t)(band + 1);
    if (band == num_bands - 1) *hi = UINT64_MAX;
    return KV_OK;
}

template <int BITS>
static int kv_launch_increment(KvCtx *ctx, const KvView &v, uint64_t flat_bytes, const uint64_t *d_hashes,
                               const uint32_t *d_valid, uint64_t n)
{
    // grid-stride kernel: exactly one wave of resident CTAs (48 registers -> 5 CTAs of 256 threads
    // per SM; the generic 8 per SM would run as 1.6 waves with a long tail)
    static int per_sm = 0;
    if (!per_sm) {
        if (const char *e = getenv("KV_INC_CTAS")) per_sm = atoi(e);
        if (per_sm <= 0 && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kv_increment_kernel<BITS, true, false>, 256, 0) != cudaSuccess)
            per_sm = 0;
        if (per_sm <= 0) per_sm = 4;
    }
    unsigned grid = kv_grid_for(ctx, n, per_sm);
    const uint64_t stride = (n + 31) / 32 + 1;
    uint32_t *added = nullptr;
    unsigned *dirty = nullptr;
    if (BITS != 1) {
        KV_TRY(kv_buf_ensure(ctx->added, stride * 4 * KV_TABLES_DEV));
        added = (uint32_t *)ctx->added.p;
        if (ctx->dirty_next == 64) {
            CU(cudaMemsetAsync(ctx->dirty, 0, 64 * sizeof(unsigned), ctx->compute));
            ctx->dirty_next = 0;
        }
        dirty = ctx->dirty + ctx->dirty_next++;
    }
#define KV_INC(CLS_, VALID_, EXACT_)                                                                                  \
    LAUNCH_C(CLS_, ctx, (kv_increment_kernel<BITS, VALID_, EXACT_>), grid, 256, v, d_hashes, d_valid, n, added, stride, dirty, \
             ctx->counters + 5)
    kv_l2_window(ctx, v.tab[0], flat_bytes);
    if (d_valid) KV_INC(KV_PROF_INCREMENT, true, false); else KV_INC(KV_PROF_INCREMENT, false, false);
    if (BITS != 1) {
        // fix-up pair: both exit at once unless the speculative pass saw a counter overflow
        LAUNCH_C(KV_PROF_FIXUP, ctx, kv_rollback_kernel<BITS>, grid, 256, v, d_hashes, n, added, stride, dirty);
        if (d_valid) KV_INC(KV_PROF_FIXUP, true, true); else KV_INC(KV_PROF_FIXUP, false, true);
    }
#undef KV_INC
    kv_l2_window(ctx, nullptr, 0);
    return KV_OK;
}

// Region-partitioned update of one chunk (K3b): per-CTA histogram rows -> scan -> scatter -> apply
// (+ fix-up pair).  hist and scatter MUST use the same grid: CTA b owns the same slice of
// positions in both.
template <typename K, typename... Args>
static int kv_launch_smem(KvCtx *ctx, int cls, K kern, unsigned grid, unsigned block, size_t smem, Args... args)
{
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->profiling) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, ctx->compute); }
    kern<<<grid, block, smem, ctx->compute>>>(args...);
    if (ctx->profiling) { cudaEventRecord(e1, ctx->compute); ctx->prof_events.push_back({cls, {e0, e1}}); }
    ctx->launches++;
    CU(cudaGetLastError());
    return KV_OK;
}

template <int BITS>
static int kv_launch_partitioned(KvCtx *ctx, const KvView &v, const KvPartInfo &pi, const uint64_t *d_hashes,
                                 const uint32_t *d_valid, uint64_t n)
{
    const int P = (int)pi.pbase[v.n_tables];
    const uint64_t max_items = n * (uint64_t)v.n_tables;
    if (max_items >= 0xffffffffull) return kv_fail(KV_EINVAL, "internal: partitioned chunk too large");
    const unsigned grid = kv_grid_for(ctx, n);                       // <= 8 CTAs per SM
    const uint64_t slice = (((n + grid - 1) / grid) + 31) & ~(uint64_t)31;   // positions per CTA
    KV_TRY(kv_buf_ensure(ctx->part_items, max_items * 4));
    KV_TRY(kv_buf_ensure(ctx->part_small, ((size_t)P * grid + 2 * (size_t)P + 16) * 4));
    KV_TRY(kv_buf_ensure(ctx->added, (max_items / 32 + 64) * 4));
    uint32_t *rows = (uint32_t *)ctx->part_small.p, *runsum = rows + (size_t)P * grid, *runbase = runsum + P,
             *meta = runbase + P;
    uint32_t *items = (uint32_t *)ctx->part_items.p, *added = (uint32_t *)ctx->added.p;
    if (ctx->dirty_next == 64) {
        CU(cudaMemsetAsync(ctx->dirty, 0, 64 * sizeof(unsigned), ctx->compute));
        ctx->dirty_next = 0;
    }
    unsigned *dirty = ctx->dirty + ctx->dirty_next++;
    KV_TRY(kv_launch_smem(ctx, KV_PROF_PARTITION, kv_part_hist_kernel, grid, 256, (size_t)P * 4, v, pi, d_hashes, d_valid, n,
                          slice, rows));
    LAUNCH_C(KV_PROF_PARTITION, ctx, kv_part_rowscan_kernel, (unsigned)P, 256, rows, (int)grid, runsum);
    LAUNCH_C(KV_PROF_PARTITION, ctx, kv_part_scan_kernel, 1, 32, pi, v.n_tables, runsum, runbase, meta);
    KV_TRY(kv_launch_smem(ctx, KV_PROF_PARTITION, kv_part_scatter_kernel, grid, 256, (size_t)P * 4, v, pi, d_hashes, d_valid,
                          n, slice, rows, runbase, items));
    static int apply_per_sm = 0;   // one wave of resident CTAs, like kv_launch_increment
    if (!apply_per_sm && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&apply_per_sm, kv_part_apply_kernel<BITS, 0>, 256, 0) != cudaSuccess ||
                          apply_per_sm <= 0))
        apply_per_sm = 4;
    const unsigned agrid = kv_grid_for(ctx, max_items, apply_per_sm);
    LAUNCH_C(KV_PROF_INCREMENT, ctx, (kv_part_apply_kernel<BITS, 0>), agrid, 256, v, items, meta, added, dirty, ctx->counters + 5);
    LAUNCH_C(KV_PROF_FIXUP, ctx, (kv_part_apply_kernel<BITS, 1>), agrid, 256, v, items, meta, added, dirty, ctx->counters + 5);
    LAUNCH_C(KV_PROF_FIXUP, ctx, (kv_part_apply_kernel<BITS, 2>), agrid, 256, v, items, meta, added, dirty, ctx->counters + 5);
    return KV_OK;
}

// ---- exact n_unique_kmers (K5; kv_kernels.cuh has the algorithm)

struct KvFreshPre {
    bool fused0;        // table 0's pass A runs inside the hash kernel
    uint32_t tag0;
    uint64_t range;     // buckets covered by first[]
};

// a fresh epoch tag for one pass over first[] (15 passes per memset)
static int kv_first_tag(KvCtx *ctx, uint32_t *tag)
{
    if (ctx->first_epoch == 0) {
        CU(cudaMemsetAsync(ctx->first.p, 0xff, ctx->first.cap, ctx->compute));
        ctx->first_epoch = 15;
    }
    *tag = (uint32_t)(--ctx->first_epoch) << KV_POS_BITS;
    return KV_OK;
}

// Before the chunk is hashed: which buckets are empty right now (= at chunk start), the first[] scratch,
// and -- when table 0 fits first[] -- the tag under which the hash kernel runs table 0's pass A.
static int kv_fresh_prepare(KvCtx *ctx, const kv_sketch *s, const KvView &v, bool may_fuse, KvFreshPre *pre, bool rebuild_occ = true,
                            const uint32_t *done_tag0 = nullptr)
{
    uint64_t maxsize = 0;
    for (int t = 0; t < s->n_tables; t++) maxsize = std::max(maxsize, s->sizes[t]);
    const uint64_t range = std::max<uint64_t>(1, std::min(maxsize, ctx->first_range));
    if (range * 4 > ctx->first.cap) {
        if (kv_buf_ensure(ctx->first, range * 4) != KV_OK)
            return kv_fail(KV_ENOMEM, "exact n_unique_kmers tracking needs %llu bytes of scratch HBM; switch it off with "
                           "kv_sketch_set_unique_tracking(sketch, 0)", (unsigned long long)(range * 4));
        ctx->first_epoch = 0;   // (re)allocated: memset before the first pass
    }
    if (s->bits != 1 && rebuild_occ) {   // one streaming pass over the counters, all tables in one launch
        uint64_t words = 0;
        for (int t = 0; t < s->n_tables; t++) words = std::max(words, (s->sizes[t] + 31) / 32);
        dim3 grid(kv_grid_for(ctx, words, 16), (unsigned)s->n_tables);
        if (s->bits == 8) LAUNCH_C(KV_PROF_UNIQUE, ctx, kv_occ_rebuild_all_kernel<8>, grid, 256, v);
        else LAUNCH_C(KV_PROF_UNIQUE, ctx, kv_occ_rebuild_all_kernel<4>, grid, 256, v);
    }
    pre->range = range;
    pre->fused0 = may_fuse && s->sizes[0] <= range && s->sizes[0] > 0;
    pre->tag0 = 0;
    if (pre->fused0) {
        if (done_tag0) pre->tag0 = *done_tag0;   // table 0's pass A already ran under this tag (deferred n_unique)
        else KV_TRY(kv_first_tag(ctx, &pre->tag0));
    }
    return KV_OK;
}

// The chunk's contribution to n_unique_kmers; must run before the chunk's increments.  Leaves the bitmap
// of new positions in ctx->fresh.  dist_counts != NULL: also histogram dist_counts.get(h) over them.
static int kv_count_fresh(KvCtx *ctx, const kv_sketch *s, const KvView &v, const KvFreshPre &pre, const uint64_t *d_hashes,
                          const uint32_t *d_valid, uint64_t n, const kv_sketch *dist_counts = nullptr,
                          unsigned long long *d_hist = nullptr, unsigned long long *d_unique = nullptr, bool own_occupancy = true)
{
    if (!d_unique) d_unique = s->d_unique;
    if (n > (1ull << KV_POS_BITS)) return kv_fail(KV_EINVAL, "internal: n_unique chunk larger than 2^%d positions", KV_POS_BITS);
    const uint64_t n_words = (n + 31) / 32;
    const uint64_t n_segs = (n + (1u << KV_SEG_LOG2) - 1) >> KV_SEG_LOG2, list_len = n_segs << KV_SEG_LOG2;
    KV_TRY(kv_buf_ensure(ctx->fresh, n_words * 4));
    KV_TRY(kv_buf_ensure(ctx->list_h, list_len * 8));
    KV_TRY(kv_buf_ensure(ctx->list_p, list_len * 4));
    KV_TRY(kv_buf_ensure(ctx->seg_cnt, n_segs * 4));
    uint32_t *first = (uint32_t *)ctx->first.p, *fresh = (uint32_t *)ctx->fresh.p, *seg_cnt = (uint32_t *)ctx->seg_cnt.p;
    uint64_t *list_h = (uint64_t *)ctx->list_h.p;
    uint32_t *list_p = (uint32_t *)ctx->list_p.p;
    const unsigned grid = kv_grid_for(ctx, n);
    const unsigned sgrid = (unsigned)std::min<uint64_t>(n_segs, (uint64_t)ctx->sm_count * 8);
    // The unrolled kernels (4 positions per thread, first[] loaded before the occupancy bit is known) pay when most
    // buckets are empty at chunk start -- a sample counted into its own sketch.  With another rank's occupancy as
    // the occupied set (kv_unique_batch / kv_unique_last_batch) most of those early loads are wasted, and next to a
    // merge on the merge lane their registers cost residency: C2 at N = 2, step 10.27 ms unrolled, 9.23 ms plain.
    const bool wide = own_occupancy && !ctx->forked;
#define KV_COMPACT_LAUNCH(FUSED_, U_)                                                                                             \
    LAUNCH_C(KV_PROF_UNIQUE, ctx, (kv_first_compact_kernel<FUSED_, U_>), sgrid, 256, v, (const uint32_t *)first, d_hashes, d_valid, n, \
             fresh, list_h, list_p, seg_cnt, d_unique)
    if (pre.fused0) { if (wide) KV_COMPACT_LAUNCH(true, 4); else KV_COMPACT_LAUNCH(true, 1); }
    else { if (wide) KV_COMPACT_LAUNCH(false, 4); else KV_COMPACT_LAUNCH(false, 1); }
#undef KV_COMPACT_LAUNCH
    for (int t = pre.fused0 ? 1 : 0; t < s->n_tables; t++)
        for (uint64_t lo = 0; lo < s->sizes[t]; lo += pre.range) {
            const uint64_t nb = std::min(pre.range, s->sizes[t] - lo);
            uint32_t tag;
            KV_TRY(kv_first_tag(ctx, &tag));
            if (wide) {
                LAUNCH_C(KV_PROF_UNIQUE, ctx, kv_first_min_list_kernel<4>, sgrid, 256, v, t, first, tag, (const uint64_t *)list_h,
                         (const uint32_t *)list_p, (const uint32_t *)seg_cnt, n_segs, lo, nb);
                LAUNCH_C(KV_PROF_UNIQUE, ctx, kv_first_own_list_kernel<4>, sgrid, 256, v, t, (const uint32_t *)first, tag,
                         (const uint64_t *)list_h, (const uint32_t *)list_p, (const uint32_t *)seg_cnt, n_segs, lo, nb, fresh,
                         d_unique);
            } else {
                LAUNCH_C(KV_PROF_UNIQUE, ctx, kv_first_min_list_kernel<1>, sgrid, 256, v, t, first, tag, (const uint64_t *)list_h,
                         (const uint32_t *)list_p, (const uint32_t *)seg_cnt, n_segs, lo, nb);
                LAUNCH_C(KV_PROF_UNIQUE, ctx, kv_first_own_list_kernel<1>, sgrid, 256, v, t, (const uint32_t *)first, tag,
                         (const uint64_t *)list_h, (const uint32_t *)list_p, (const uint32_t *)seg_cnt, n_segs, lo, nb, fresh,
                         d_unique);
            }
        }
    if (dist_counts)
        LAUNCH_C(KV_PROF_UNIQUE, ctx, kv_abund_dist_kernel, grid, 256, kv_view(dist_counts), d_hashes, (const uint32_t *)fresh, n,
                 d_hist);
    return KV_OK;
}

// apply one chunk of hashes (device, n < 2^32) to the sketch: exact-unique bookkeeping first
// (it must see the buckets as they were before this chunk), then the saturating increments
static int kv_apply_hashes(KvCtx *ctx, kv_sketch *s, const uint64_t *d_hashes, const uint32_t *d_valid, uint64_t n,
                           const kv_sketch *dist_counts = nullptr, unsigned long long *d_hist = nullptr,
                           const KvFreshPre *fused = nullptr)
{
    if (!n) return KV_OK;
    KvView v = kv_view(s);
    if (s->span) {   // no hot bitmap, no n_unique: the exact update for every hash (add(), mask building ... -- small inputs)
        if (dist_counts) return kv_fail(KV_EINVAL, "abundance distribution is not defined on spanning sketches");
        LAUNCH_C(KV_PROF_INCREMENT, ctx, kv_add_exact_kernel, kv_grid_for(ctx, n), 256, v, d_hashes, d_valid, n);
        return KV_OK;
    }
    if (s->state_stale) {
        KV_TRY(kv_state_rebuild_locked(ctx, s));
        s->state_stale = false;
    }
    if (s->track_unique || dist_counts) {
        KvFreshPre pre;
        if (fused) pre = *fused;   // the caller ran kv_fresh_prepare before hashing (table 0's pass A is done)
        else KV_TRY(kv_fresh_prepare(ctx, s, v, false, &pre));
        KV_TRY(kv_count_fresh(ctx, s, v, pre, d_hashes, d_valid, n, dist_counts, d_hist));
    }
    if (!s->track_unique) s->unique_valid = false;
    // abundance_distribution counts a k-mer into the tracking sketch only when it was new there
    if (dist_counts) d_valid = (const uint32_t *)ctx->fresh.p;
    // large counter sketches: region-partitioned updates (tables must index with 32 bits, <= 4 tables)
    bool partitioned = s->bits != 1 && s->flat_bytes >= ctx->part_min_bytes && s->flat_bytes <= ctx->part_max_bytes &&
                       s->n_tables <= 4 && ctx->update_path != 1;
    if (ctx->update_path == 2) partitioned = s->bits != 1 && s->n_tables <= 4;
    KvPartInfo pi;
    memset(&pi, 0, sizeof pi);
    if (partitioned) {
        pi.rb = ctx->part_region_log2 + (s->bits == 4 ? 1 : 0);
        uint64_t runs = 0;
        for (int t = 0; t < s->n_tables && partitioned; t++) {
            if (s->sizes[t] >= 0xffffffffull) partitioned = false;
            pi.pbase[t] = (uint32_t)runs;
            runs += ((s->sizes[t] - 1) >> pi.rb) + 1;
        }
        pi.pbase[s->n_tables] = (uint32_t)runs;
        if (runs > KV_PART_MAX) partitioned = false;
    }
    if (partitioned) {
        if (s->bits == 8) return kv_launch_partitioned<8>(ctx, v, pi, d_hashes, d_valid, n);
        return kv_launch_partitioned<4>(ctx, v, pi, d_hashes, d_valid, n);
    }
    if (s->bits == 8) return kv_launch_increment<8>(ctx, v, s->flat_bytes, d_hashes, d_valid, n);
    if (s->bits == 4) return kv_launch_increment<4>(ctx, v, s->flat_bytes, d_hashes, d_valid, n);
    return kv_launch_increment<1>(ctx, v, s->flat_bytes, d_hashes, d_valid, n);
}

// ---- K3c: tiled update path

// Region size: 2^rb buckets, the configured value for large tables; small tables get smaller regions so that
// every table still has thousands of runs (few runs = the cursor atomics of a whole chunk pile onto few words).
static int kv_tile_rb_for(const KvCtx *ctx, const kv_sketch *s)
{
    if (s->bits == 1) return 16;
    uint64_t smallest = UINT64_MAX;
    for (int t = 0; t < s->n_tables; t++)
        if (s->sizes[t]) smallest = std::min(smallest, s->sizes[t]);
    int rb = ctx->tile_rb;
    while (rb > 8 && (smallest >> rb) < 8192) rb--;
    return rb;
}

// Slab interleave block: the write frontier (runs x block x 2 bytes) should stay well inside L2, or the
// 2-byte stores reach HBM as partial sectors.
static int kv_tile_blk_for(const KvCtx *ctx, uint64_t runs)
{
    if (ctx->tile_block_log2 < 0) return ctx->tile_block_log2;
    int blk = ctx->tile_block_log2;
    while (blk > 3 && ((runs << blk) * 2) > (24ull << 20)) blk--;
    return blk;
}

struct KvTilePlan {
    bool on;
    KvTileInfo ti;
    uint32_t runs, direct_below;
    size_t smem;
    uint64_t chunk_tiles;
};

// Should batches of up to `batch_pos` positions update `s` through the tiled path, and with what geometry?
static int kv_tile_plan(KvCtx *ctx, const kv_sketch *s, uint64_t batch_pos, KvTilePlan *pl)
{
    memset(pl, 0, sizeof *pl);
    if (s->span) {   // fixed at creation: every rank must use the same geometry and the exported buffers
        pl->on = true;
        pl->ti = s->span->ti;
        pl->runs = s->span->runs;
        pl->smem = s->span->smem;
        pl->direct_below = s->span->direct_below;
        pl->chunk_tiles = s->span->chunk_pos / KV_TILE;
        return KV_OK;
    }
    bool on = ctx->update_path == 3 || (ctx->update_path == 0 && s->flat_bytes >= ctx->tile_min_bytes);
    if (!on) return KV_OK;
    const int rb = kv_tile_rb_for(ctx, s);
    uint64_t runs = 0, min_regions = UINT64_MAX;
    for (int t = 0; t < s->n_tables; t++) {
        pl->ti.run_base[t] = (uint32_t)runs;
        const uint64_t nr = s->sizes[t] ? ((s->sizes[t] - 1) >> rb) + 1 : 0;
        runs += nr;
        if (nr) min_regions = std::min(min_regions, nr);
    }
    if (!runs || runs >= (1ull << 30)) return KV_OK;   // nothing held here / absurdly many regions: in-place updates
    pl->ti.run_base[s->n_tables] = (uint32_t)runs;
    pl->ti.run_base[KV_TABLES_DEV] = (uint32_t)runs;   // kv_slab_index reads the total there
    uint64_t chunk_pos = std::min<uint64_t>(ctx->tile_chunk_bases, (batch_pos + KV_TILE - 1) / KV_TILE * KV_TILE);
    if (s->track_unique) chunk_pos = std::min<uint64_t>(chunk_pos, 1ull << KV_POS_BITS);   // n_unique positions are 28-bit
    // slots per run: mean + 6 % + 8 sigma + slack, for offsets spread by the hash; anything beyond (skewed
    // inputs) is updated in place by the producer
    const double mean = (double)chunk_pos / (double)min_regions;
    uint64_t cap = (uint64_t)(mean * 1.0625 + 8.0 * sqrt(mean) + 64.0);
    const int blk_log2 = kv_tile_blk_for(ctx, runs);
    const uint64_t blk = blk_log2 >= 0 ? (1ull << blk_log2) : 16;
    cap = (cap + blk - 1) / blk * blk;
    if (cap >= 0xffffffffull) return KV_OK;
    KV_TRY(kv_buf_ensure(ctx->tile_cursor, runs * 4 + 16));
    if (kv_buf_ensure(ctx->tile_slab, runs * cap * 2) != KV_OK) { g_err.clear(); return KV_OK; }   // no room for slabs: in-place updates
    pl->on = true;
    pl->ti.rb = rb;
    pl->ti.cap = (uint32_t)cap;
    pl->ti.cursor = (uint32_t *)ctx->tile_cursor.p;
    pl->ti.slab = (uint16_t *)ctx->tile_slab.p;
    pl->ti.blk_log2 = blk_log2;
    pl->ti.ovf_any = (unsigned *)((uint32_t *)ctx->tile_cursor.p + runs);   // cleared together with the cursors
    if (s->track_unique) {
        pl->ti.ovf_stride = chunk_pos / 32 + 1;
        KV_TRY(kv_buf_ensure(ctx->tile_ovf, pl->ti.ovf_stride * 4 * (uint64_t)s->n_tables));
        pl->ti.ovf = (uint32_t *)ctx->tile_ovf.p;
    }
    pl->runs = (uint32_t)runs;
    pl->smem = s->bits == 1 ? ((size_t)1 << rb) / 8 : ((size_t)1 << rb) * 2;
    const uint64_t region_bytes = ((uint64_t)1 << rb) * (uint64_t)s->bits / 8;
    pl->direct_below = ctx->tile_direct_below >= 0 ? (uint32_t)ctx->tile_direct_below : (uint32_t)std::max<uint64_t>(1, region_bytes / 64);
    pl->chunk_tiles = chunk_pos / KV_TILE;
    return KV_OK;
}

template <int BITS, bool SPAN>
static int kv_launch_tile_apply2(KvCtx *ctx, const KvView &v, const KvTilePlan &pl, const KvTileSources &src)
{
    static bool configured = false;
    if (!configured) {
        CU(cudaFuncSetAttribute(kv_tile_apply_kernel<BITS, SPAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
        configured = true;
    }
    unsigned grid = pl.runs;
    if (SPAN) {   // resident grid striding over this rank's regions
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kv_tile_apply_kernel<BITS, SPAN>, KV_TILE_THREADS, pl.smem) != cudaSuccess || per_sm < 1)
            per_sm = 1;
        grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(src.own_total, (uint64_t)ctx->sm_count * per_sm));
    }
    return kv_launch_smem(ctx, KV_PROF_INCREMENT, kv_tile_apply_kernel<BITS, SPAN>, grid, KV_TILE_THREADS, pl.smem, v, pl.ti, pl.direct_below, src);
}

static int kv_launch_tile_apply(KvCtx *ctx, const kv_sketch *s, const KvView &v, const KvTilePlan &pl)
{
    KvTileSources src;
    memset(&src, 0, sizeof src);
    src.n = 1;
    if (s->span) {
        src.n = s->span->world;
        src.rank = s->span->rank;
        for (int r = 0; r < s->span->world; r++) { src.cursor[r] = s->span->cursor[r]; src.slab[r] = s->span->slab[r]; }
        const uint64_t region_bytes = std::max<uint64_t>(1, (((uint64_t)1 << pl.ti.rb) * (uint64_t)s->bits) / 8);
        for (int t = 0; t < s->n_tables; t++) {
            src.piece[t] = s->span->piece[t];
            // regions of table t inside this rank's piece (pieces are multiples of 2 MB, regions powers of two <= 64 KB)
            const uint64_t n_regions = pl.ti.run_base[t + 1] - pl.ti.run_base[t];
            const uint64_t lo = std::min<uint64_t>(n_regions, (uint64_t)s->span->rank * s->span->piece[t] / region_bytes);
            const uint64_t hi = std::min<uint64_t>(n_regions, (uint64_t)(s->span->rank + 1) * s->span->piece[t] / region_bytes);
            src.own_lo[t] = (uint32_t)lo;
            src.own_n[t] = (uint32_t)(hi - lo);
            src.own_total += src.own_n[t];
        }
        if (s->bits == 8) return kv_launch_tile_apply2<8, true>(ctx, v, pl, src);
        if (s->bits == 4) return kv_launch_tile_apply2<4, true>(ctx, v, pl, src);
        return kv_launch_tile_apply2<1, true>(ctx, v, pl, src);
    }
    if (s->bits == 8) return kv_launch_tile_apply2<8, false>(ctx, v, pl, src);
    if (s->bits == 4) return kv_launch_tile_apply2<4, false>(ctx, v, pl, src);
    return kv_launch_tile_apply2<1, false>(ctx, v, pl, src);
}

static int kv_check_mask(const kv_sketch *s, const kv_sketch *mask)
{
    if (!mask) return KV_OK;
    if (mask->device != s->device) return kv_fail(KV_EINVAL, "mask lives on another device");
    if (mask->ksize != s->ksize || mask->hasher != s->hasher)
        return kv_fail(KV_EINVAL, "mask must use the same k-mer size and hash function as the sketch");
    return KV_OK;
}

struct kv_peer_sync;
static int kv_peer_barrier_locked(KvCtx *ctx, kv_peer_sync *ps);

// span_sync != NULL: `s` is a spanning sketch and the call is COLLECTIVE -- every rank runs exactly
// n_chunks chunks (empty ones when its batch is shorter), with two device-side rank barriers per chunk
static int kv_consume_impl(kv_sketch *s, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                           int where, int num_bands, int band, const kv_sketch *mask, int mask_threshold,
                           int consume_masked, uint64_t *n_kmers_out, kv_peer_sync *span_sync, uint64_t n_chunks)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    if (n_kmers_out) *n_kmers_out = 0;
    if (s->span && !span_sync) return kv_fail(KV_EINVAL, "spanning sketches are updated with kv_consume_batch_span (collective)");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->last_hashed.sketch = nullptr;   // whatever this call does, the scratch no longer describes "the last batch"
    if (n_reads == 0 && !span_sync) return KV_OK;
    if (n_reads && (!bases || !offsets)) return kv_fail(KV_EINVAL, "null batch pointers");
    KV_TRY(kv_check_mask(s, mask));
    uint64_t lo = 0, hi = 0;
    if (num_bands > 0) KV_TRY(kv_band_interval(num_bands, band, &lo, &hi));
    CU(cudaSetDevice(s->device));
    KvBatch b;
    memset(&b, 0, sizeof b);
    if (n_reads) KV_TRY(kv_stage(ctx, bases, offsets, n_reads, where, 0, &b));
    if (b.total == 0 && !span_sync) { kv_stage_done(ctx, &b); return KV_OK; }
    CU(cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long), ctx->compute));

    KvTilePlan plan;
    KV_TRY(kv_tile_plan(ctx, s, b.n_tiles * KV_TILE, &plan));
    // positions are 32-bit inside a chunk, and the partitioned path indexes n_tables items per position
    const uint64_t chunk_limit = plan.on ? plan.chunk_tiles * KV_TILE
                                         : std::min<uint64_t>(ctx->chunk_bases, (0xfffffff0ull / (uint64_t)s->n_tables) / KV_TILE * KV_TILE);
    const uint64_t chunk_tiles = chunk_limit / KV_TILE;
    const uint64_t chunk_pos = std::min<uint64_t>(chunk_limit, b.n_tiles * KV_TILE);
    const bool need_hashes = !plan.on || s->track_unique;
    KV_TRY(kv_hash_scratch(ctx, chunk_pos, need_hashes));
    const KvView sv = kv_view(s);
    const uint64_t own_chunks = (b.n_tiles + chunk_tiles - 1) / chunk_tiles;
    if (span_sync && own_chunks > n_chunks)
        return kv_fail(KV_EINVAL, "this rank's batch needs %llu chunks, the collective call announced %llu",
                       (unsigned long long)own_chunks, (unsigned long long)n_chunks);
    const uint64_t total_chunks = span_sync ? n_chunks : own_chunks;
    bool deferred_pass_a = false;
    uint32_t deferred_tag0 = 0;
    for (uint64_t ci = 0; ci < total_chunks; ci++) {
        const uint64_t t0 = ci * chunk_tiles;
        if (t0 >= b.n_tiles) {   // collective call, nothing left here: take part in the exchange with empty slabs
            CU(cudaMemsetAsync(plan.ti.cursor, 0, (size_t)plan.runs * 4 + 4, ctx->compute));
            KV_TRY(kv_peer_barrier_locked(ctx, span_sync));
            KV_TRY(kv_launch_tile_apply(ctx, s, sv, plan));
            KV_TRY(kv_peer_barrier_locked(ctx, span_sync));
            continue;
        }
        uint64_t nt = std::min(chunk_tiles, b.n_tiles - t0);
        uint64_t npos = std::min<uint64_t>(nt * KV_TILE, b.total - t0 * KV_TILE);
        KvHashParams p;
        memset(&p, 0, sizeof p);
        p.bases = b.d_bases; p.offsets = b.d_offsets; p.tile_first = (const uint32_t *)ctx->tile_first.p;
        p.total = b.total; p.tile0 = t0; p.k = s->ksize;
        p.banded = num_bands > 0; p.band_lo = lo; p.band_hi = hi;
        if (mask) { p.use_mask = 1; p.mask = kv_view(mask); p.mask_threshold = mask_threshold; p.consume_masked = consume_masked != 0; }
        p.strict = 0;
        p.hashes = need_hashes ? (uint64_t *)ctx->hashes.p : nullptr; p.valid = (uint32_t *)ctx->valid.p; p.n_valid = ctx->counters;
        KvFreshPre pre;
        if (s->track_unique) {
            if (s->state_stale && !plan.on) {   // (the in-place path rebuilds its hot bitmap anyway; do it before occ is derived)
                KV_TRY(kv_state_rebuild_locked(ctx, s));
                s->state_stale = false;
            }
            KV_TRY(kv_fresh_prepare(ctx, s, sv, ctx->unique_fuse0, &pre));
            if (pre.fused0) { p.track0 = 1; p.first0 = (uint32_t *)ctx->first.p; p.tag0 = pre.tag0; p.sk = sv; }
        } else if (s->defer_unique && total_chunks == 1 && !plan.on && ctx->unique_fuse0) {
            // deferred n_unique (kv_unique_last_batch will follow): table 0's pass A rides along in the hash kernel now.
            // Which buckets the OTHER ranks occupy is not known yet, but the owner of a bucket -- the first position of
            // this batch that touches it -- does not depend on that; the later passes only ask whether the bucket counts.
            if (s->state_stale) {
                KV_TRY(kv_state_rebuild_locked(ctx, s));
                s->state_stale = false;
            }
            KV_TRY(kv_fresh_prepare(ctx, s, sv, true, &pre));
            if (pre.fused0) {
                p.track0 = 1; p.first0 = (uint32_t *)ctx->first.p; p.tag0 = pre.tag0; p.sk = sv;
                deferred_pass_a = true; deferred_tag0 = pre.tag0;
            }
        }
        if (plan.on) {
            p.scatter = 1; p.ti = plan.ti; p.sk = sv;
            CU(cudaMemsetAsync(plan.ti.cursor, 0, (size_t)plan.runs * 4 + 4, ctx->compute));
            if (plan.ti.ovf) CU(cudaMemsetAsync(plan.ti.ovf, 0, plan.ti.ovf_stride * 4 * (uint64_t)s->n_tables, ctx->compute));
        }
        if (s->hasher == KV_HASH_TWOBIT) KV_TRY(kv_launch_hash<KV_HASH_TWOBIT>(ctx, p, (unsigned)nt));
        else KV_TRY(kv_launch_hash<KV_HASH_MURMUR>(ctx, p, (unsigned)nt));
        if (!plan.on) {
            KV_TRY(kv_apply_hashes(ctx, s, p.hashes, p.valid, npos, nullptr, nullptr, s->track_unique ? &pre : nullptr));
            continue;
        }
        // tiled path: the producer has filed the updates; the exact-n_unique passes must still see the
        // tables as they were before this chunk, then one CTA per region applies its slab
        if (s->track_unique) {
            KV_TRY(kv_count_fresh(ctx, s, sv, pre, p.hashes, p.valid, npos));
            LAUNCH_C(KV_PROF_FIXUP, ctx, kv_tile_overflow_kernel, kv_grid_for(ctx, npos / 32 + 1), 256, sv, plan.ti, p.hashes, npos);
        } else
            s->unique_valid = false;
        if (span_sync) KV_TRY(kv_peer_barrier_locked(ctx, span_sync));   // every rank has filed this chunk's updates
        KV_TRY(kv_launch_tile_apply(ctx, s, sv, plan));
        if (span_sync) KV_TRY(kv_peer_barrier_locked(ctx, span_sync));   // every rank has read my slabs: they may be refilled
        s->state_stale = true;   // the hot bitmap of the in-place path is not maintained here
    }
    kv_stage_done(ctx, &b);
    if (total_chunks == 1 && need_hashes && !span_sync) {   // hashes + valid bits of the whole batch stay in the scratch
        ctx->last_hashed.sketch = s;
        ctx->last_hashed.npos = std::min<uint64_t>(chunk_tiles * KV_TILE, b.total);
        ctx->last_hashed.pass_a = deferred_pass_a;
        ctx->last_hashed.tag0 = deferred_tag0;
    }
    if (n_kmers_out) {
        CU(cudaMemcpyAsync(ctx->h_counters, ctx->counters, 8, cudaMemcpyDeviceToHost, ctx->compute));
        CU(cudaStreamSynchronize(ctx->compute));
        *n_kmers_out = ctx->h_counters[0];
    }
    return KV_OK;
}

extern "C" int kv_consume_batch(kv_sketch *s, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                                int where, int num_bands, int band, const kv_sketch *mask, int mask_threshold,
                                int consume_masked, uint64_t *n_kmers_out)
{
    return kv_consume_impl(s, bases, offsets, n_reads, where, num_bands, band, mask, mask_threshold, consume_masked, n_kmers_out,
                           nullptr, 0);
}

extern "C" int kv_consume_batch_span(kv_sketch *s, kv_peer_sync *ps, const uint8_t *bases, const uint64_t *offsets,
                                     uint64_t n_reads, int where, uint64_t n_chunks, int num_bands, int band,
                                     const kv_sketch *mask, int mask_threshold, int consume_masked, uint64_t *n_kmers_out)
{
    if (!s || !s->span || !s->span->ready) return kv_fail(KV_EINVAL, "not a (ready) spanning sketch");
    if (!ps) return kv_fail(KV_EINVAL, "a spanning update needs the device-side rank barrier (kv_peer_sync)");
    return kv_consume_impl(s, bases, offsets, n_reads, where, num_bands, band, mask, mask_threshold, consume_masked, n_kmers_out,
                           ps, n_chunks);
}

// ------------------------------------------------------------------ novel

template <int HASHER, bool FAST>
static int kv_launch_novel2(KvCtx *ctx, const KvNovelParams &p)
{
    unsigned n_tiles = (unsigned)p.n_tiles;
    if (HASHER == KV_HASH_TWOBIT) { LAUNCH_C(KV_PROF_NOVEL, ctx, (kv_novel_kernel<KV_HASH_TWOBIT, 4, FAST>), n_tiles, KV_THREADS, p); return KV_OK; }
    int kw = 4 * ((p.k + 15) / 16);
    switch (kw) {
    case 4: LAUNCH_C(KV_PROF_NOVEL, ctx, (kv_novel_kernel<KV_HASH_MURMUR, 4, FAST>), n_tiles, KV_THREADS, p); break;
    case 8: LAUNCH_C(KV_PROF_NOVEL, ctx, (kv_novel_kernel<KV_HASH_MURMUR, 8, FAST>), n_tiles, KV_THREADS, p); break;
    case 12: LAUNCH_C(KV_PROF_NOVEL, ctx, (kv_novel_kernel<KV_HASH_MURMUR, 12, FAST>), n_tiles, KV_THREADS, p); break;
    default: LAUNCH_C(KV_PROF_NOVEL, ctx, (kv_novel_kernel<KV_HASH_MURMUR, 16, FAST>), n_tiles, KV_THREADS, p); break;
    }
    return KV_OK;
}

// the order-free evaluation applies when nothing depends on WHICH test failed first: no abundance
// screen (kevlar/novel.py:36-43) and abundances looked up here rather than handed in
template <int HASHER>
static int kv_launch_novel(KvCtx *ctx, const KvNovelParams &p)
{
    bool fast = p.screen <= 0 && !getenv("KV_NOVEL_REFERENCE_ORDER");
    for (int i = 0; i < p.n_case + p.n_ctrl; i++)
        if (p.pre[i]) fast = false;
    return fast ? kv_launch_novel2<HASHER, true>(ctx, p) : kv_launch_novel2<HASHER, false>(ctx, p);
}

static int kv_novel_impl(const kv_sketch *const *cases, int n_case, const kv_sketch *const *ctrls, int n_ctrl,
                         const uint8_t *const *pre, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                         int where, int case_min, int ctrl_max, int screen, int num_bands, int64_t band_minus_1,
                         kv_hit *hits, uint64_t max_hits, uint64_t *n_hits, uint8_t *read_flags, uint32_t *discard_pos)
{
    if (n_hits) *n_hits = 0;
    if (n_case < 1 || !cases || !cases[0]) return kv_fail(KV_EINVAL, "need at least one case sketch");
    if (n_ctrl < 0 || n_case + n_ctrl > KV_MAX_SAMPLES) return kv_fail(KV_EINVAL, "at most %d sketches per scan", KV_MAX_SAMPLES);
    if (!n_hits || !read_flags || (max_hits && !hits)) return kv_fail(KV_EINVAL, "null output pointer");
    if (screen > 0 && !discard_pos) return kv_fail(KV_EINVAL, "discard_pos is required when the abundance screen is on");
    const kv_sketch *c0 = cases[0];
    KvNovelParams p;
    memset(&p, 0, sizeof p);
    for (int i = 0; i < n_case + n_ctrl; i++) {
        const kv_sketch *s = i < n_case ? cases[i] : ctrls[i - n_case];
        if (!s) return kv_fail(KV_EINVAL, "null sketch in sample list");
        if (s->device != c0->device) return kv_fail(KV_EINVAL, "all sketches of a scan must live on one device");
        if (s->ksize != c0->ksize || s->hasher != c0->hasher)
            return kv_fail(KV_EINVAL, "all sketches of a scan must share k-mer size and hash function");
        p.sk[i] = kv_view(s);
        p.pre[i] = pre ? pre[i] : nullptr;
    }
    if (n_reads == 0) return KV_OK;
    if (!bases || !offsets) return kv_fail(KV_EINVAL, "null batch pointers");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(c0->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(c0->device));
    KvBatch b;
    KV_TRY(kv_stage(ctx, bases, offsets, n_reads, where, 0, &b));

    KV_TRY(kv_buf_ensure(ctx->flags, n_reads * 4));
    CU(cudaMemsetAsync(ctx->flags.p, 0, n_reads * 4, ctx->compute));
    if (screen > 0) {
        KV_TRY(kv_buf_ensure(ctx->discard, n_reads * 4));
        CU(cudaMemsetAsync(ctx->discard.p, 0xff, n_reads * 4, ctx->compute));
    }
    KV_TRY(kv_buf_ensure(ctx->hits, std::max<uint64_t>(max_hits, 1) * sizeof(kv_hit)));
    CU(cudaMemsetAsync(ctx->counters + 2, 0, sizeof(unsigned long long), ctx->compute));

    p.bases = b.d_bases; p.offsets = b.d_offsets; p.tile_first = (const uint32_t *)ctx->tile_first.p;
    p.total = b.total; p.n_tiles = b.n_tiles; p.k = c0->ksize;
    p.n_case = n_case; p.n_ctrl = n_ctrl; p.case_min = case_min; p.ctrl_max = ctrl_max; p.screen = screen;
    p.banded = num_bands > 0; p.band_mask = num_bands > 0 ? (uint64_t)(num_bands - 1) : 0; p.band_minus_1 = band_minus_1;
    p.hits = (kv_hit *)ctx->hits.p; p.max_hits = max_hits; p.n_hits = ctx->counters + 2;
    p.read_flags = (uint32_t *)ctx->flags.p; p.discard_pos = screen > 0 ? (uint32_t *)ctx->discard.p : nullptr;
    // reads shorter than k are skipped (kevlar/novel.py:134); flagged on the device so that host- and
    // device-resident batches behave alike
    LAUNCH(ctx, kv_short_reads_kernel, kv_grid_for(ctx, n_reads), 256, b.d_offsets, n_reads, c0->ksize, (uint32_t *)ctx->flags.p);
    if (b.total) {
        if (c0->hasher == KV_HASH_TWOBIT) KV_TRY(kv_launch_novel<KV_HASH_TWOBIT>(ctx, p));
        else KV_TRY(kv_launch_novel<KV_HASH_MURMUR>(ctx, p));
    }
    kv_stage_done(ctx, &b);

    // results back to the host: the hit count, and the reads with a flag or a discard position as a
    // compact list (the per-read arrays stay on the device unless nearly every read is noted)
    const uint64_t note_cap = std::max<uint64_t>(4096, n_reads / 16);
    KV_TRY(kv_buf_ensure(ctx->notes, note_cap * sizeof(KvReadNote)));
    CU(cudaMemsetAsync(ctx->counters + 6, 0, sizeof(unsigned long long), ctx->compute));
    LAUNCH(ctx, kv_read_notes_kernel, kv_grid_for(ctx, n_reads), 256, (const uint32_t *)ctx->flags.p,
           screen > 0 ? (const uint32_t *)ctx->discard.p : (const uint32_t *)nullptr, n_reads, (KvReadNote *)ctx->notes.p,
           (unsigned long long)note_cap, ctx->counters + 6);
    CU(cudaMemcpyAsync(ctx->h_counters + 2, ctx->counters + 2, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaMemcpyAsync(ctx->h_counters + 6, ctx->counters + 6, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    uint64_t found = ctx->h_counters[2];
    const uint64_t n_notes = ctx->h_counters[6];
    if (found > max_hits) { *n_hits = found; return kv_fail(KV_EOVERFLOW, "hit buffer too small: %llu hits, room for %llu", (unsigned long long)found, (unsigned long long)max_hits); }
    std::vector<kv_hit> hh(found);
    if (found) CU(cudaMemcpyAsync(hh.data(), ctx->hits.p, found * sizeof(kv_hit), cudaMemcpyDeviceToHost, ctx->compute));
    // read-level flags (kevlar/novel.py:134-139,152-154); hdisc = discard position of the reads that have one
    memset(read_flags, 0, n_reads);
    if (screen > 0) for (uint64_t r = 0; r < n_reads; r++) discard_pos[r] = 0xffffffffu;
    auto note = [&](uint64_t r, uint32_t f, uint32_t d) {
        uint8_t fl = (uint8_t)(f & KV_READ_SKIPPED);
        if (screen > 0) {
            if (!(fl & KV_READ_SKIPPED) && d != 0xffffffffu) fl |= KV_READ_DISCARDED;
            discard_pos[r] = (fl & KV_READ_SKIPPED) ? 0xffffffffu : d;
        }
        read_flags[r] = fl;
    };
    if (n_notes <= note_cap) {
        std::vector<KvReadNote> notes(n_notes);
        if (n_notes) CU(cudaMemcpyAsync(notes.data(), ctx->notes.p, n_notes * sizeof(KvReadNote), cudaMemcpyDeviceToHost, ctx->compute));
        CU(cudaStreamSynchronize(ctx->compute));
        for (const KvReadNote &nt : notes) note(nt.read, nt.flags, nt.discard);
    } else {   // nearly every read is noted: fetch the per-read arrays
        std::vector<uint32_t> hflags(n_reads), hd;
        CU(cudaMemcpyAsync(hflags.data(), ctx->flags.p, n_reads * 4, cudaMemcpyDeviceToHost, ctx->compute));
        if (screen > 0) {
            hd.resize(n_reads);
            CU(cudaMemcpyAsync(hd.data(), ctx->discard.p, n_reads * 4, cudaMemcpyDeviceToHost, ctx->compute));
        }
        CU(cudaStreamSynchronize(ctx->compute));
        for (uint64_t r = 0; r < n_reads; r++) note(r, hflags[r], screen > 0 ? hd[r] : 0xffffffffu);
    }
    // a discarded read keeps its raw discard position for the hit filter below even when it is also skipped
    auto raw_discard = [&](uint32_t r) -> uint32_t { return screen > 0 ? discard_pos[r] : 0xffffffffu; };
    uint64_t kept = 0;
    for (uint64_t i = 0; i < found; i++) {
        const kv_hit &h = hh[i];
        if (read_flags[h.read] & KV_READ_SKIPPED) continue;
        const uint32_t d = raw_discard(h.read);
        if (d != 0xffffffffu && h.offset > d) continue;
        hh[kept++] = h;
    }
    std::sort(hh.begin(), hh.begin() + kept, [](const kv_hit &a, const kv_hit &b2) {
        return a.read != b2.read ? a.read < b2.read : a.offset < b2.offset;
    });
    if (kept) memcpy(hits, hh.data(), kept * sizeof(kv_hit));
    *n_hits = kept;
    return KV_OK;
}

extern "C" int kv_novel_batch(const kv_sketch *const *cases, int n_case, const kv_sketch *const *ctrls, int n_ctrl,
                              const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads, int where, int case_min,
                              int ctrl_max, int screen, int num_bands, int64_t band_minus_1, kv_hit *hits,
                              uint64_t max_hits, uint64_t *n_hits, uint8_t *read_flags, uint32_t *discard_pos)
{
    for (int i = 0; i < n_case + n_ctrl && i < KV_MAX_SAMPLES; i++) {
        const kv_sketch *s = i < n_case ? (cases ? cases[i] : nullptr) : (ctrls ? ctrls[i - n_case] : nullptr);
        if (s && s->n_shards > 1)
            return kv_fail(KV_EINVAL, "sharded sketches are scanned with kv_get_hashes_dev + kv_novel_from_counts");
    }
    return kv_novel_impl(cases, n_case, ctrls, n_ctrl, nullptr, bases, offsets, n_reads, where, case_min, ctrl_max, screen,
                         num_bands, band_minus_1, hits, max_hits, n_hits, read_flags, discard_pos);
}

extern "C" int kv_novel_from_counts(const kv_sketch *like, int n_case, int n_ctrl, const uint8_t *const *dev_counts,
                                    const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads, int where,
                                    int case_min, int ctrl_max, int screen, int num_bands, int64_t band_minus_1,
                                    kv_hit *hits, uint64_t max_hits, uint64_t *n_hits, uint8_t *read_flags,
                                    uint32_t *discard_pos)
{
    if (!like || !dev_counts) return kv_fail(KV_EINVAL, "null argument");
    if (n_case < 1 || n_ctrl < 0 || n_case + n_ctrl > KV_MAX_SAMPLES) return kv_fail(KV_EINVAL, "bad sample counts");
    const kv_sketch *all[KV_MAX_SAMPLES];
    for (int i = 0; i < n_case + n_ctrl; i++) {
        if (!dev_counts[i]) return kv_fail(KV_EINVAL, "null count array");
        all[i] = like;   // only k, the hasher and the device are taken from it: every abundance comes from dev_counts
    }
    return kv_novel_impl(all, n_case, all + n_case, n_ctrl, dev_counts, bases, offsets, n_reads, where, case_min, ctrl_max,
                         screen, num_bands, band_minus_1, hits, max_hits, n_hits, read_flags, discard_pos);
}

// ------------------------------------------------------------------ device-resident hash streams (sharded sketches)

extern "C" int kv_hash_batch_dev(int hasher, int ksize, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                                 int where, int num_bands, int band, int device, uint64_t *dev_hashes,
                                 uint32_t *dev_valid, uint64_t capacity, uint64_t *n_positions, uint64_t *n_kmers)
{
    if (n_positions) *n_positions = 0;
    if (n_kmers) *n_kmers = 0;
    if (!n_reads) return KV_OK;
    if (!bases || !offsets || !dev_hashes || !dev_valid) return kv_fail(KV_EINVAL, "null argument");
    int kmax = hasher == KV_HASH_TWOBIT ? KV_MAX_KSIZE_TWOBIT : KV_MAX_KSIZE_MURMUR;
    if (ksize < 1 || ksize > kmax) return kv_fail(KV_EINVAL, "k-mer size %d not supported", ksize);
    uint64_t lo = 0, hi = 0;
    if (num_bands > 0) KV_TRY(kv_band_interval(num_bands, band, &lo, &hi));
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(device));
    KvBatch b;
    KV_TRY(kv_stage(ctx, bases, offsets, n_reads, where, 0, &b));
    if (b.n_tiles * KV_TILE > capacity) {
        kv_stage_done(ctx, &b);
        return kv_fail(KV_EOVERFLOW, "hash buffer holds %llu positions, the batch needs %llu", (unsigned long long)capacity,
                       (unsigned long long)(b.n_tiles * KV_TILE));
    }
    CU(cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long), ctx->compute));
    if (b.total) {
        KvHashParams p;
        memset(&p, 0, sizeof p);
        p.bases = b.d_bases; p.offsets = b.d_offsets; p.tile_first = (const uint32_t *)ctx->tile_first.p;
        p.total = b.total; p.tile0 = 0; p.k = ksize;
        p.banded = num_bands > 0; p.band_lo = lo; p.band_hi = hi;
        p.hashes = dev_hashes; p.valid = dev_valid; p.n_valid = ctx->counters;
        if (hasher == KV_HASH_TWOBIT) KV_TRY(kv_launch_hash<KV_HASH_TWOBIT>(ctx, p, (unsigned)b.n_tiles));
        else KV_TRY(kv_launch_hash<KV_HASH_MURMUR>(ctx, p, (unsigned)b.n_tiles));
    }
    kv_stage_done(ctx, &b);
    CU(cudaMemcpyAsync(ctx->h_counters, ctx->counters, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    if (n_positions) *n_positions = b.total;
    if (n_kmers) *n_kmers = ctx->h_counters[0];
    return KV_OK;
}

extern "C" int kv_add_hashes_dev(kv_sketch *s, const uint64_t *dev_hashes, const uint32_t *dev_valid, uint64_t n)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    if (!n) return KV_OK;
    if (!dev_hashes || !dev_valid) return kv_fail(KV_EINVAL, "null argument");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    const uint64_t step = std::min<uint64_t>(ctx->chunk_bases, (0xfffffff0ull / (uint64_t)s->n_tables) / KV_TILE * KV_TILE);
    for (uint64_t o = 0; o < n; o += step)   // step is a multiple of 32: valid words stay aligned
        KV_TRY(kv_apply_hashes(ctx, s, dev_hashes + o, dev_valid + o / 32, std::min(step, n - o)));
    CU(cudaStreamSynchronize(ctx->compute));
    return KV_OK;
}

extern "C" int kv_get_hashes_dev(const kv_sketch *s, const uint64_t *dev_hashes, const uint32_t *dev_valid, uint64_t n,
                                 uint8_t *dev_counts)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    if (!n) return KV_OK;
    if (!dev_hashes || !dev_counts) return kv_fail(KV_EINVAL, "null argument");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    LAUNCH(ctx, kv_get_kernel, kv_grid_for(ctx, n), 256, kv_view(s), dev_hashes, dev_valid, n, dev_counts);
    CU(cudaStreamSynchronize(ctx->compute));
    return KV_OK;
}

// ------------------------------------------------------------------ point / list operations

extern "C" int kv_hash_kmers(int hasher, int ksize, const uint8_t *kmers, uint64_t n, int device, uint64_t *hashes_out,
                             uint8_t *ok_out)
{
    if (!n) return KV_OK;
    if (!kmers || !hashes_out || !ok_out) return kv_fail(KV_EINVAL, "null argument");
    int kmax = hasher == KV_HASH_TWOBIT ? KV_MAX_KSIZE_TWOBIT : KV_MAX_KSIZE_MURMUR;
    if (ksize < 1 || ksize > kmax) return kv_fail(KV_EINVAL, "k-mer size %d not supported", ksize);
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    std::vector<uint64_t> offs(n + 1);
    for (uint64_t i = 0; i <= n; i++) offs[i] = i * (uint64_t)ksize;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(device));
    KvBatch b;
    KV_TRY(kv_stage(ctx, kmers, offs.data(), n, KV_MEM_HOST, 0, &b));
    CU(cudaStreamSynchronize(ctx->copy));   // offs is a local
    const uint64_t npos = b.n_tiles * KV_TILE;
    KV_TRY(kv_hash_scratch(ctx, npos));
    KV_TRY(kv_buf_ensure(ctx->misc, n * 9));
    KvHashParams p;
    memset(&p, 0, sizeof p);
    p.bases = b.d_bases; p.offsets = b.d_offsets; p.tile_first = (const uint32_t *)ctx->tile_first.p;
    p.total = b.total; p.tile0 = 0; p.k = ksize; p.strict = 1;
    p.hashes = (uint64_t *)ctx->hashes.p; p.valid = (uint32_t *)ctx->valid.p; p.n_valid = ctx->counters;
    if (hasher == KV_HASH_TWOBIT) KV_TRY(kv_launch_hash<KV_HASH_TWOBIT>(ctx, p, (unsigned)b.n_tiles));
    else KV_TRY(kv_launch_hash<KV_HASH_MURMUR>(ctx, p, (unsigned)b.n_tiles));
    uint64_t *d_out = (uint64_t *)ctx->misc.p;
    uint8_t *d_ok = (uint8_t *)ctx->misc.p + n * 8;
    LAUNCH(ctx, kv_gather_kernel, kv_grid_for(ctx, n), 256, p.hashes, p.valid, n, (uint64_t)ksize, d_out, d_ok);
    kv_stage_done(ctx, &b);
    CU(cudaMemcpyAsync(hashes_out, d_out, n * 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaMemcpyAsync(ok_out, d_ok, n, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    return KV_OK;
}

extern "C" int kv_get_hashes(const kv_sketch *s, const uint64_t *hashes, uint64_t n, uint8_t *counts_out)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    if (!n) return KV_OK;
    if (!hashes || !counts_out) return kv_fail(KV_EINVAL, "null argument");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    KV_TRY(kv_buf_ensure(ctx->misc, n * 9));
    uint64_t *d_h = (uint64_t *)ctx->misc.p;
    uint8_t *d_c = (uint8_t *)ctx->misc.p + n * 8;
    CU(cudaMemcpyAsync(d_h, hashes, n * 8, cudaMemcpyHostToDevice, ctx->compute));
    LAUNCH(ctx, kv_get_kernel, kv_grid_for(ctx, n), 256, kv_view(s), d_h, (const uint32_t *)nullptr, n, d_c);
    CU(cudaMemcpyAsync(counts_out, d_c, n, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    return KV_OK;
}

extern "C" int kv_add_hashes(kv_sketch *s, const uint64_t *hashes, uint64_t n)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    if (!n) return KV_OK;
    if (!hashes) return kv_fail(KV_EINVAL, "null argument");
    if (n >= 0xffffffffull) return kv_fail(KV_EINVAL, "at most 2^32-2 hashes per call");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    KV_TRY(kv_buf_ensure(ctx->misc, n * 8));
    CU(cudaMemcpyAsync(ctx->misc.p, hashes, n * 8, cudaMemcpyHostToDevice, ctx->compute));
    const uint64_t step = std::min<uint64_t>(ctx->chunk_bases, 1ull << 26);
    for (uint64_t o = 0; o < n; o += step)
        KV_TRY(kv_apply_hashes(ctx, s, (const uint64_t *)ctx->misc.p + o, nullptr, std::min(step, n - o)));
    CU(cudaStreamSynchronize(ctx->compute));
    return KV_OK;
}

extern "C" int kv_kmer_counts_batch(const kv_sketch *s, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                                    int where, uint64_t *hashes_out, uint8_t *counts_out, uint8_t *valid_out)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    if (!n_reads) return KV_OK;
    if (!bases || !offsets) return kv_fail(KV_EINVAL, "null batch pointers");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    KvBatch b;
    KV_TRY(kv_stage(ctx, bases, offsets, n_reads, where, 0, &b));
    if (!b.total) { kv_stage_done(ctx, &b); return KV_OK; }
    const uint64_t npos = b.n_tiles * KV_TILE;
    KV_TRY(kv_hash_scratch(ctx, npos));
    KV_TRY(kv_buf_ensure(ctx->misc, b.total * 2));
    KvHashParams p;
    memset(&p, 0, sizeof p);
    p.bases = b.d_bases; p.offsets = b.d_offsets; p.tile_first = (const uint32_t *)ctx->tile_first.p;
    p.total = b.total; p.tile0 = 0; p.k = s->ksize; p.strict = 1;
    p.hashes = (uint64_t *)ctx->hashes.p; p.valid = (uint32_t *)ctx->valid.p; p.n_valid = ctx->counters;
    if (s->hasher == KV_HASH_TWOBIT) KV_TRY(kv_launch_hash<KV_HASH_TWOBIT>(ctx, p, (unsigned)b.n_tiles));
    else KV_TRY(kv_launch_hash<KV_HASH_MURMUR>(ctx, p, (unsigned)b.n_tiles));
    uint8_t *d_counts = (uint8_t *)ctx->misc.p, *d_valid8 = (uint8_t *)ctx->misc.p + b.total;
    unsigned grid = kv_grid_for(ctx, b.total);
    LAUNCH(ctx, kv_get_kernel, grid, 256, kv_view(s), p.hashes, p.valid, b.total, d_counts);
    LAUNCH(ctx, kv_expand_bits_kernel, grid, 256, p.valid, b.total, d_valid8);
    kv_stage_done(ctx, &b);
    if (hashes_out) CU(cudaMemcpyAsync(hashes_out, p.hashes, b.total * 8, cudaMemcpyDeviceToHost, ctx->compute));
    if (counts_out) CU(cudaMemcpyAsync(counts_out, d_counts, b.total, cudaMemcpyDeviceToHost, ctx->compute));
    if (valid_out) CU(cudaMemcpyAsync(valid_out, d_valid8, b.total, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    return KV_OK;
}

// ------------------------------------------------------------------ abundance distribution

extern "C" int kv_abund_dist_batch(const kv_sketch *counts, kv_sketch *tracking, const uint8_t *bases,
                                   const uint64_t *offsets, uint64_t n_reads, int where, uint64_t *dist_out)
{
    if (!counts || !tracking || !dist_out) return kv_fail(KV_EINVAL, "null argument");
    if (counts == tracking) return kv_fail(KV_EINVAL, "the tracking sketch must not be the counts sketch");
    if (tracking->device != counts->device) return kv_fail(KV_EINVAL, "tracking sketch lives on another device");
    if (tracking->ksize != counts->ksize || tracking->hasher != counts->hasher)
        return kv_fail(KV_EINVAL, "tracking sketch must use the same k-mer size and hash function as the counts");
    if (counts->n_shards > 1 || tracking->n_shards > 1)
        return kv_fail(KV_EINVAL, "abundance distribution is not defined on bin-range shards");
    memset(dist_out, 0, 256 * sizeof(uint64_t));
    if (n_reads == 0) return KV_OK;
    if (!bases || !offsets) return kv_fail(KV_EINVAL, "null batch pointers");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(counts->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(counts->device));
    KvBatch b;
    KV_TRY(kv_stage(ctx, bases, offsets, n_reads, where, 0, &b));
    if (b.total == 0) { kv_stage_done(ctx, &b); return KV_OK; }
    KV_TRY(kv_buf_ensure(ctx->hist, 256 * sizeof(unsigned long long)));
    unsigned long long *d_hist = (unsigned long long *)ctx->hist.p;
    CU(cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned long long), ctx->compute));
    CU(cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long), ctx->compute));
    // same chunking as kv_consume_batch: the tracking updates of chunk i are in place before the
    // first-touch passes of chunk i+1
    const uint64_t chunk_limit = std::min<uint64_t>(ctx->chunk_bases, (0xfffffff0ull / (uint64_t)tracking->n_tables) / KV_TILE * KV_TILE);
    const uint64_t chunk_tiles = chunk_limit / KV_TILE;
    const uint64_t chunk_pos = std::min<uint64_t>(chunk_limit, b.n_tiles * KV_TILE);
    KV_TRY(kv_hash_scratch(ctx, chunk_pos));
    for (uint64_t t0 = 0; t0 < b.n_tiles; t0 += chunk_tiles) {
        uint64_t nt = std::min(chunk_tiles, b.n_tiles - t0);
        uint64_t npos = std::min<uint64_t>(nt * KV_TILE, b.total - t0 * KV_TILE);
        KvHashParams p;
        memset(&p, 0, sizeof p);
        p.bases = b.d_bases; p.offsets = b.d_offsets; p.tile_first = (const uint32_t *)ctx->tile_first.p;
        p.total = b.total; p.tile0 = t0; p.k = counts->ksize;
        p.hashes = (uint64_t *)ctx->hashes.p; p.valid = (uint32_t *)ctx->valid.p; p.n_valid = ctx->counters;
        if (counts->hasher == KV_HASH_TWOBIT) KV_TRY(kv_launch_hash<KV_HASH_TWOBIT>(ctx, p, (unsigned)nt));
        else KV_TRY(kv_launch_hash<KV_HASH_MURMUR>(ctx, p, (unsigned)nt));
        KV_TRY(kv_apply_hashes(ctx, tracking, p.hashes, p.valid, npos, counts, d_hist));
    }
    kv_stage_done(ctx, &b);
    CU(cudaMemcpyAsync(dist_out, d_hist, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    return KV_OK;
}

// ------------------------------------------------------------------ n_unique_kmers across ranks
//
// khmer's n_unique_kmers is defined on ONE stream of reads.  With the reads sharded contiguously over R ranks
// (rank-major order = file order) an occurrence on rank r is new  <=>  it is the first on rank r to touch a
// bucket that was empty at the start AND that no lower rank touches.  So, once every rank has counted its
// shard into a zeroed partial sketch: rank r ORs the occupancy bitmaps of the partial sketches of ranks < r
// (kv_sketch_occupancy, exchanged by the caller), and kv_unique_batch re-runs the first-touch passes over its
// reads with THAT as the occupied set; the sum over the ranks is the reference's number.

extern "C" int kv_sketch_occupancy(kv_sketch *s, uint32_t **dev_words_out, uint64_t *n_words_out)
{
    if (!s || !dev_words_out || !n_words_out) return kv_fail(KV_EINVAL, "null argument");
    if (s->span || s->n_shards > 1) return kv_fail(KV_EINVAL, "occupancy bitmaps are kept for ordinary sketches only");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    const KvView v = kv_view(s);
    if (s->bits != 1) {
        uint64_t words = 0;
        for (int t = 0; t < s->n_tables; t++) words = std::max(words, (s->sizes[t] + 31) / 32);
        dim3 grid(kv_grid_for(ctx, words, 16), (unsigned)s->n_tables);
        if (s->bits == 8) LAUNCH_C(KV_PROF_UNIQUE, ctx, kv_occ_rebuild_all_kernel<8>, grid, 256, v);
        else LAUNCH_C(KV_PROF_UNIQUE, ctx, kv_occ_rebuild_all_kernel<4>, grid, 256, v);
    }
    for (int t = 0; t < s->n_tables; t++) {
        // a bit table IS its occupancy bitmap (same bit order); its allocation is padded to 256 bytes
        dev_words_out[t] = s->bits == 1 ? (uint32_t *)(s->flat + s->toff[t]) : v.occ[t];
        n_words_out[t] = (s->sizes[t] + 31) / 32;
    }
    return KV_OK;
}

extern "C" int kv_sketch_set_unique(kv_sketch *s, uint64_t n_unique)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    unsigned long long v = n_unique;
    CU(cudaMemcpyAsync(s->d_unique, &v, 8, cudaMemcpyHostToDevice, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    s->n_unique = n_unique;
    s->unique_valid = true;
    return KV_OK;
}

extern "C" int kv_sketch_set_unique_dev(kv_sketch *s, const uint64_t *dev_n_unique)
{
    if (!s || !dev_n_unique) return kv_fail(KV_EINVAL, "null argument");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(s->d_unique, dev_n_unique, 8, cudaMemcpyDeviceToDevice, ctx->compute));
    s->unique_valid = true;
    return KV_OK;
}

extern "C" int kv_unique_batch(const kv_sketch *like, uint32_t *const *dev_occupied, const uint8_t *bases, const uint64_t *offsets,
                               uint64_t n_reads, int where, int num_bands, int band, const kv_sketch *mask,
                               int mask_threshold, int consume_masked, uint64_t *n_unique_out, uint64_t *dev_n_unique_out)
{
    if (!like || !dev_occupied || (!n_unique_out && !dev_n_unique_out)) return kv_fail(KV_EINVAL, "null argument");
    if (n_unique_out) *n_unique_out = 0;
    if (n_reads == 0) {
        if (dev_n_unique_out) CU(cudaMemset(dev_n_unique_out, 0, 8));
        return KV_OK;
    }
    if (!bases || !offsets) return kv_fail(KV_EINVAL, "null batch pointers");
    KV_TRY(kv_check_mask(like, mask));
    uint64_t lo = 0, hi = 0;
    if (num_bands > 0) KV_TRY(kv_band_interval(num_bands, band, &lo, &hi));
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(like->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(like->device));
    KvBatch b;
    KV_TRY(kv_stage(ctx, bases, offsets, n_reads, where, 0, &b));
    if (b.total == 0) {
        kv_stage_done(ctx, &b);
        if (dev_n_unique_out) CU(cudaMemsetAsync(dev_n_unique_out, 0, 8, ctx->compute));
        return KV_OK;
    }
    // the view the first-touch kernels see: `like`'s geometry, the caller's bitmaps as the occupied set
    KvView v = kv_view(like);
    v.bits = 8;   // (kv_bucket_empty reads occ[] for counters and the table itself for bit tables: always occ[] here)
    for (int t = 0; t < like->n_tables; t++) {
        if (!dev_occupied[t]) return kv_fail(KV_EINVAL, "null occupancy bitmap for table %d", t);
        v.occ[t] = dev_occupied[t];
    }
    unsigned long long *d_unique = ctx->counters + 7;
    CU(cudaMemsetAsync(d_unique, 0, sizeof(unsigned long long), ctx->compute));
    CU(cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long), ctx->compute));
    const uint64_t chunk_limit = std::min<uint64_t>(ctx->chunk_bases, 1ull << KV_POS_BITS);
    const uint64_t chunk_tiles = chunk_limit / KV_TILE;
    const uint64_t chunk_pos = std::min<uint64_t>(chunk_limit, b.n_tiles * KV_TILE);
    KV_TRY(kv_hash_scratch(ctx, chunk_pos));
    for (uint64_t t0 = 0; t0 < b.n_tiles; t0 += chunk_tiles) {
        const uint64_t nt = std::min(chunk_tiles, b.n_tiles - t0);
        const uint64_t npos = std::min<uint64_t>(nt * KV_TILE, b.total - t0 * KV_TILE);
        KvHashParams p;
        memset(&p, 0, sizeof p);
        p.bases = b.d_bases; p.offsets = b.d_offsets; p.tile_first = (const uint32_t *)ctx->tile_first.p;
        p.total = b.total; p.tile0 = t0; p.k = like->ksize;
        p.banded = num_bands > 0; p.band_lo = lo; p.band_hi = hi;
        if (mask) { p.use_mask = 1; p.mask = kv_view(mask); p.mask_threshold = mask_threshold; p.consume_masked = consume_masked != 0; }
        p.hashes = (uint64_t *)ctx->hashes.p; p.valid = (uint32_t *)ctx->valid.p; p.n_valid = ctx->counters;
        KvFreshPre pre;
        KV_TRY(kv_fresh_prepare(ctx, like, v, ctx->unique_fuse0, &pre, false));
        if (pre.fused0) { p.track0 = 1; p.first0 = (uint32_t *)ctx->first.p; p.tag0 = pre.tag0; p.sk = v; }
        if (like->hasher == KV_HASH_TWOBIT) KV_TRY(kv_launch_hash<KV_HASH_TWOBIT>(ctx, p, (unsigned)nt));
        else KV_TRY(kv_launch_hash<KV_HASH_MURMUR>(ctx, p, (unsigned)nt));
        KV_TRY(kv_count_fresh(ctx, like, v, pre, p.hashes, p.valid, npos, nullptr, nullptr, d_unique, false));
        if (t0 + chunk_tiles < b.n_tiles) {   // later chunks of this rank must see this chunk's buckets as occupied
            const uint64_t n_segs = (npos + (1u << KV_SEG_LOG2) - 1) >> KV_SEG_LOG2;
            const unsigned sgrid = (unsigned)std::min<uint64_t>(n_segs, (uint64_t)ctx->sm_count * 8);
            LAUNCH_C(KV_PROF_UNIQUE, ctx, kv_occ_mark_list_kernel, sgrid, 256, v, (const uint64_t *)ctx->list_h.p,
                     (const uint32_t *)ctx->seg_cnt.p, n_segs);
        }
    }
    kv_stage_done(ctx, &b);
    if (dev_n_unique_out) CU(cudaMemcpyAsync(dev_n_unique_out, d_unique, 8, cudaMemcpyDeviceToDevice, ctx->compute));
    if (n_unique_out) {
        CU(cudaMemcpyAsync(ctx->h_counters + 7, d_unique, 8, cudaMemcpyDeviceToHost, ctx->compute));
        CU(cudaStreamSynchronize(ctx->compute));
        *n_unique_out = ctx->h_counters[7];
    }
    return KV_OK;
}

// The same passes over the batch that kv_consume_batch counted into `like` LAST on this device, when its hashes
// and valid bits are still in the device scratch (one chunk, nothing hashed since): no second copy of the reads
// to the device, no second MurmurHash -- table 0's pass A runs as a plain pass over the stored hashes.
// KV_ESTATE when the scratch holds something else; the caller then falls back to kv_unique_batch.
extern "C" int kv_unique_last_batch(const kv_sketch *like, uint32_t *const *dev_occupied, uint64_t *n_unique_out,
                                    uint64_t *dev_n_unique_out)
{
    if (!like || !dev_occupied || (!n_unique_out && !dev_n_unique_out)) return kv_fail(KV_EINVAL, "null argument");
    if (n_unique_out) *n_unique_out = 0;
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(like->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(like->device));
    if (ctx->last_hashed.sketch != like || !ctx->last_hashed.npos)
        return kv_fail(KV_ESTATE, "the hashes of the last batch counted into this sketch are no longer in the device scratch");
    const uint64_t npos = ctx->last_hashed.npos;
    if (npos > (1ull << KV_POS_BITS)) return kv_fail(KV_ESTATE, "batch too long for one first-touch chunk");
    KvView v = kv_view(like);
    v.bits = 8;   // always occ[] (see kv_unique_batch)
    for (int t = 0; t < like->n_tables; t++) {
        if (!dev_occupied[t]) return kv_fail(KV_EINVAL, "null occupancy bitmap for table %d", t);
        v.occ[t] = dev_occupied[t];
    }
    unsigned long long *d_unique = ctx->counters + 7;
    CU(cudaMemsetAsync(d_unique, 0, sizeof(unsigned long long), ctx->compute));
    const uint64_t *d_hashes = (const uint64_t *)ctx->hashes.p;
    const uint32_t *d_valid = (const uint32_t *)ctx->valid.p;
    KvFreshPre pre;
    const bool pass_a_done = ctx->last_hashed.pass_a;
    KV_TRY(kv_fresh_prepare(ctx, like, v, ctx->unique_fuse0, &pre, false, pass_a_done ? &ctx->last_hashed.tag0 : nullptr));
    if (pre.fused0 && !pass_a_done)
        LAUNCH_C(KV_PROF_UNIQUE, ctx, kv_first_min0_kernel, kv_grid_for(ctx, npos), 256, v, (uint32_t *)ctx->first.p, pre.tag0,
                 d_hashes, d_valid, npos);
    ctx->last_hashed.sketch = nullptr;   // first[] is about to be reused for the other tables: the shortcut works once
    KV_TRY(kv_count_fresh(ctx, like, v, pre, d_hashes, d_valid, npos, nullptr, nullptr, d_unique, false));
    if (dev_n_unique_out) CU(cudaMemcpyAsync(dev_n_unique_out, d_unique, 8, cudaMemcpyDeviceToDevice, ctx->compute));
    if (n_unique_out) {
        CU(cudaMemcpyAsync(ctx->h_counters + 7, d_unique, 8, cudaMemcpyDeviceToHost, ctx->compute));
        CU(cudaStreamSynchronize(ctx->compute));
        *n_unique_out = ctx->h_counters[7];
    }
    return KV_OK;
}

// ------------------------------------------------------------------ multi-GPU merge

extern "C" int kv_sketch_widen(kv_sketch *s, void *dev_out, uint64_t *n_elems, int *elem_bytes)
{
    if (!s) return kv_fail(KV_EINVAL, "null sketch");
    uint64_t n = s->bits == 4 ? s->flat_bytes * 2 : s->flat_bytes;
    if (n_elems) *n_elems = n;
    if (elem_bytes) *elem_bytes = s->bits == 8 ? 2 : 1;
    if (!dev_out) return KV_OK;   // size query
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    LAUNCH_C(KV_PROF_MERGE, ctx, kv_widen_kernel, kv_grid_for(ctx, s->flat_bytes, 16), 256, s->flat, s->flat_bytes, s->bits, dev_out);
    CU(cudaStreamSynchronize(ctx->compute));
    return KV_OK;
}

extern "C" int kv_sketch_narrow(kv_sketch *s, const void *dev_in)
{
    if (!s || !dev_in) return kv_fail(KV_EINVAL, "null argument");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    LAUNCH_C(KV_PROF_MERGE, ctx, kv_narrow_kernel, kv_grid_for(ctx, s->flat_bytes, 16), 256, s->flat, s->flat_bytes, s->bits, dev_in);
    s->state_stale = true;
    CU(cudaStreamSynchronize(ctx->compute));
    s->unique_valid = false;
    return KV_OK;
}

static int kv_check_range(const kv_sketch *s, uint64_t &lo, uint64_t &hi)
{
    if (hi == 0) hi = s->flat_bytes;
    if (lo > hi || hi > s->flat_bytes || (lo & 255) || (hi & 255))
        return kv_fail(KV_EINVAL, "byte range must be 256-byte aligned and inside the table storage");
    return KV_OK;
}

static int kv_merge_peers_impl(kv_sketch *s, void *const *peer_flat, int n_peers, uint64_t byte_lo, uint64_t byte_hi, bool push)
{
    if (!s || (n_peers && !peer_flat)) return kv_fail(KV_EINVAL, "null argument");
    if (n_peers < 0 || n_peers > KV_MAX_RANKS - 1) return kv_fail(KV_EINVAL, "at most %d peers per merge", KV_MAX_RANKS - 1);
    KV_TRY(kv_check_range(s, byte_lo, byte_hi));
    if (!n_peers || byte_lo == byte_hi) return KV_OK;
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    KvPeers peers;
    memset(&peers, 0, sizeof peers);
    peers.n = n_peers;
    for (int i = 0; i < n_peers; i++) peers.peer[i] = (uint4 *)((uint8_t *)peer_flat[i] + byte_lo);
    uint64_t n_vec = (byte_hi - byte_lo) / 16;
    // on the merge lane the kernel shares the SMs with the counting of the next sample: a few CTAs per SM are
    // enough to keep the links busy (8 loads in flight per thread) and leave the rest of the machine alone
    cudaStream_t lane = ctx->forked ? ctx->merge : ctx->compute;
    const int vecs = n_peers > 7 ? 1 : std::min(4, std::max(1, 8 / n_peers));   // vectors per thread and iteration (U below)
    const unsigned grid = ctx->forked ? (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_vec + 255) / 256, (uint64_t)ctx->sm_count * ctx->merge_lane_ctas))
                                      : kv_grid_for(ctx, (n_vec + vecs - 1) / vecs, 16);
    uint4 *mine = (uint4 *)(s->flat + byte_lo);
#define KV_MERGE_CASE(NP_, EXACT_)                                                                                               \
    if (push) LAUNCH_S(KV_PROF_MERGE, ctx, lane, (kv_merge_peers_kernel<true, NP_, (NP_ <= 2 ? 4 : NP_ <= 4 ? 2 : 1), EXACT_>), grid, 256, mine, n_vec, s->bits, peers); \
    else LAUNCH_S(KV_PROF_MERGE, ctx, lane, (kv_merge_peers_kernel<false, NP_, (NP_ <= 2 ? 4 : NP_ <= 4 ? 2 : 1), EXACT_>), grid, 256, mine, n_vec, s->bits, peers);     \
    break;
    switch (n_peers) {
    case 1: KV_MERGE_CASE(1, true)
    case 2: KV_MERGE_CASE(2, true)
    case 3: KV_MERGE_CASE(3, true)
    case 4: KV_MERGE_CASE(4, true)
    case 5: KV_MERGE_CASE(5, true)
    case 6: KV_MERGE_CASE(6, true)
    case 7: KV_MERGE_CASE(7, true)
    default: KV_MERGE_CASE(8, false)
    }
#undef KV_MERGE_CASE
    s->state_stale = true;
    s->unique_valid = false;
    return KV_OK;
}

extern "C" int kv_sketch_merge_peers(kv_sketch *s, const void *const *peer_flat, int n_peers, uint64_t byte_lo,
                                     uint64_t byte_hi)
{
    return kv_merge_peers_impl(s, (void *const *)peer_flat, n_peers, byte_lo, byte_hi, false);
}

extern "C" int kv_sketch_allreduce_peers(kv_sketch *s, void *const *peer_flat, int n_peers, uint64_t byte_lo, uint64_t byte_hi)
{
    return kv_merge_peers_impl(s, peer_flat, n_peers, byte_lo, byte_hi, true);
}

extern "C" int kv_sketch_copy_from_peer(kv_sketch *s, const void *peer_flat, uint64_t byte_lo, uint64_t byte_hi)
{
    if (!s || !peer_flat) return kv_fail(KV_EINVAL, "null argument");
    KV_TRY(kv_check_range(s, byte_lo, byte_hi));
    if (byte_lo == byte_hi) return KV_OK;
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(s->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(s->flat + byte_lo, (const uint8_t *)peer_flat + byte_lo, byte_hi - byte_lo,
                       cudaMemcpyDeviceToDevice, ctx->compute));
    s->state_stale = true;
    s->unique_valid = false;
    return KV_OK;
}

extern "C" int kv_sketch_ipc_export(kv_sketch *s, uint8_t handle_out[64])
{
    if (!s || !handle_out) return kv_fail(KV_EINVAL, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(s->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->flat));
    memcpy(handle_out, &h, 64);
    return KV_OK;
}

extern "C" int kv_ipc_open(int device, const uint8_t handle[64], void **dev_ptr)
{
    if (!handle || !dev_ptr) return kv_fail(KV_EINVAL, "null argument");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return KV_OK;
}

extern "C" int kv_ipc_close(int device, void *dev_ptr)
{
    CU(cudaSetDevice(device));
    CU(cudaIpcCloseMemHandle(dev_ptr));
    return KV_OK;
}

// ------------------------------------------------------------------ device-side rank barrier

struct kv_peer_sync {
    int device, rank, world;
    uint32_t *flags;            // my array (KV_MAX_RANKS slots used), own 2 MB allocation so the IPC handle maps nothing else
    unsigned *timed_out;        // device flag raised by a barrier kernel that gave up
    void *peer[KV_MAX_RANKS];   // mapped peer arrays
    uint32_t epoch;
    unsigned long long timeout_ns;
    int lane;                   // 0: barriers on the compute stream; 1: on the merge lane (kv_peer_sync_set_lane)
};

extern "C" int kv_peer_sync_create(int device, int rank, int world, kv_peer_sync **out, uint8_t handle_out[64])
{
    if (!out || !handle_out) return kv_fail(KV_EINVAL, "null argument");
    if (world < 1 || world > KV_MAX_RANKS || rank < 0 || rank >= world)
        return kv_fail(KV_EINVAL, "device-side barriers support 1..%d ranks", KV_MAX_RANKS);
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(device));
    kv_peer_sync *ps = new kv_peer_sync();
    memset(ps, 0, sizeof *ps);
    ps->device = device; ps->rank = rank; ps->world = world;
    ps->timeout_ns = 30000ull * 1000000ull;
    if (const char *e = getenv("KV_PEER_TIMEOUT_MS")) ps->timeout_ns = strtoull(e, nullptr, 10) * 1000000ull;
    const size_t bytes = 2u << 20;
    if (cudaMalloc((void **)&ps->flags, bytes) != cudaSuccess) { delete ps; return kv_fail(KV_ENOMEM, "cannot allocate barrier flags"); }
    cudaMemset(ps->flags, 0, bytes);
    ps->timed_out = (unsigned *)(ps->flags + 1024);   // same allocation, never touched by peers
    CU(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ps->flags));
    memcpy(handle_out, &h, 64);
    *out = ps;
    return KV_OK;
}

extern "C" int kv_peer_sync_connect(kv_peer_sync *ps, int peer_rank, const uint8_t handle[64])
{
    if (!ps || !handle) return kv_fail(KV_EINVAL, "null argument");
    if (peer_rank < 0 || peer_rank >= ps->world || peer_rank == ps->rank) return kv_fail(KV_EINVAL, "bad peer rank");
    if (ps->peer[peer_rank]) return kv_fail(KV_EINVAL, "peer %d is already connected", peer_rank);
    CU(cudaSetDevice(ps->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(&ps->peer[peer_rank], h, cudaIpcMemLazyEnablePeerAccess));
    return KV_OK;
}

static int kv_peer_barrier_locked(KvCtx *ctx, kv_peer_sync *ps)
{
    KvPeerFlags f;
    memset(&f, 0, sizeof f);
    for (int p = 0; p < ps->world; p++) {
        if (p != ps->rank && !ps->peer[p]) return kv_fail(KV_EINVAL, "peer %d is not connected", p);
        f.peer[p] = (uint32_t *)ps->peer[p];
    }
    f.mine = ps->flags; f.rank = ps->rank; f.world = ps->world;
    ps->epoch++;
    LAUNCH_S(KV_PROF_MERGE, ctx, ps->lane ? ctx->merge : ctx->compute, kv_peer_barrier_kernel, 1, 32, f, ps->epoch, ps->timeout_ns,
             ps->timed_out);
    return KV_OK;
}

extern "C" int kv_peer_barrier(kv_peer_sync *ps)
{
    if (!ps) return kv_fail(KV_EINVAL, "null argument");
    KvPeerFlags f;
    memset(&f, 0, sizeof f);
    for (int p = 0; p < ps->world; p++) {
        if (p != ps->rank && !ps->peer[p]) return kv_fail(KV_EINVAL, "peer %d is not connected", p);
        f.peer[p] = (uint32_t *)ps->peer[p];
    }
    f.mine = ps->flags; f.rank = ps->rank; f.world = ps->world;
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(ps->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ps->device));
    ps->epoch++;
    LAUNCH_S(KV_PROF_MERGE, ctx, ps->lane ? ctx->merge : ctx->compute, kv_peer_barrier_kernel, 1, 32, f, ps->epoch, ps->timeout_ns,
             ps->timed_out);
    return KV_OK;
}

extern "C" int kv_peer_sync_set_lane(kv_peer_sync *ps, int lane)
{
    if (!ps || lane < 0 || lane > 1) return kv_fail(KV_EINVAL, "lane is 0 (compute stream) or 1 (merge lane)");
    ps->lane = lane;
    return KV_OK;
}

// Merge lane.  kv_merge_fork: whatever is enqueued on the compute stream so far (the sample just counted) must
// finish before anything enqueued on the merge lane from now on; until kv_merge_join the peer-to-peer merge
// kernels go to the merge lane, so the merge of sample i crosses NVLink while sample i+1 is being counted.
// Fork may be called repeatedly (once per sample); kv_merge_join makes the compute stream wait for the lane.
extern "C" int kv_merge_fork(int device)
{
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(device));
    CU(cudaEventRecord(ctx->ev_fork, ctx->compute));
    CU(cudaStreamWaitEvent(ctx->merge, ctx->ev_fork, 0));
    ctx->forked = true;
    return KV_OK;
}

extern "C" int kv_merge_join(int device)
{
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(device));
    if (!ctx->forked) return KV_OK;
    CU(cudaEventRecord(ctx->ev_join, ctx->merge));
    CU(cudaStreamWaitEvent(ctx->compute, ctx->ev_join, 0));
    ctx->forked = false;
    return KV_OK;
}

extern "C" int kv_peer_sync_status(kv_peer_sync *ps)
{
    if (!ps) return kv_fail(KV_EINVAL, "null argument");
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(ps->device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ps->device));
    unsigned flag = 0;
    CU(cudaStreamSynchronize(ctx->merge));
    CU(cudaMemcpyAsync(&flag, ps->timed_out, sizeof flag, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    if (flag) return kv_fail(KV_ECUDA, "a device-side barrier timed out waiting for a peer rank (KV_PEER_TIMEOUT_MS)");
    return KV_OK;
}

extern "C" int kv_peer_sync_destroy(kv_peer_sync *ps)
{
    if (!ps) return KV_OK;
    cudaSetDevice(ps->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < ps->world; p++)
        if (ps->peer[p]) cudaIpcCloseMemHandle(ps->peer[p]);
    cudaFree(ps->flags);
    delete ps;
    return KV_OK;
}

// ------------------------------------------------------------------ synthetic reads (measurement fixture)

extern "C" int kv_synth_reads(int device, const uint8_t *const *dev_haplotypes, const uint64_t *hap_lens, int n_haps,
                              uint64_t n_reads, uint64_t first_read, uint32_t read_len, double error_rate, uint64_t seed,
                              uint8_t *dev_bases, uint64_t *dev_offsets)
{
    if (!dev_haplotypes || !hap_lens || !dev_bases) return kv_fail(KV_EINVAL, "null argument");
    if (n_haps < 1 || n_haps > 8 || read_len < 1) return kv_fail(KV_EINVAL, "1..8 haplotypes, read_len >= 1");
    if (error_rate < 0 || error_rate >= 1) return kv_fail(KV_EINVAL, "error_rate must be in [0, 1)");
    KvSynthParams p;
    memset(&p, 0, sizeof p);
    for (int h = 0; h < n_haps; h++) {
        if (!dev_haplotypes[h] || hap_lens[h] <= read_len) return kv_fail(KV_EINVAL, "haplotype %d is shorter than a read", h);
        p.hap[h] = dev_haplotypes[h];
        p.hap_len[h] = hap_lens[h];
    }
    p.n_haps = n_haps; p.n_reads = n_reads; p.first_read = first_read; p.read_len = read_len;
    p.err_q32 = (uint32_t)(error_rate * 4294967296.0);
    p.seed = seed; p.out = dev_bases; p.offsets = dev_offsets;
    if (!n_reads) return KV_OK;
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(device));
    LAUNCH(ctx, kv_synth_reads_kernel, kv_grid_for(ctx, n_reads * read_len, 16), 256, p);
    CU(cudaStreamSynchronize(ctx->compute));
    return KV_OK;
}

// ------------------------------------------------------------------ plumbing

extern "C" int kv_stream(int device, void **cuda_stream)
{
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    if (cuda_stream) *cuda_stream = (void *)ctx->compute;
    return KV_OK;
}

extern "C" int kv_sync(int device)
{
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    CU(cudaSetDevice(device));
    CU(cudaStreamSynchronize(ctx->copy));
    CU(cudaStreamSynchronize(ctx->compute));
    return KV_OK;
}

extern "C" int kv_profile(int device, int enable, double *ms_out, uint64_t *n_out)
{
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(device));
    CU(cudaStreamSynchronize(ctx->compute));
    for (auto &ev : ctx->prof_events) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ev.second.first, ev.second.second) == cudaSuccess) {
            ctx->prof_ms[ev.first] += ms;
            ctx->prof_n[ev.first]++;
        }
        cudaEventDestroy(ev.second.first);
        cudaEventDestroy(ev.second.second);
    }
    ctx->prof_events.clear();
    for (int c = 0; c < KV_PROF_CLASSES; c++) {
        if (ms_out) ms_out[c] = ctx->prof_ms[c];
        if (n_out) n_out[c] = ctx->prof_n[c];
        if (enable != 2) { ctx->prof_ms[c] = 0; ctx->prof_n[c] = 0; }   // 2 = read without resetting
    }
    if (enable != 2) ctx->profiling = enable != 0;
    return KV_OK;
}

extern "C" int kv_redo_count(int device, uint64_t *n)
{
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(device));
    CU(cudaMemcpyAsync(ctx->h_counters + 5, ctx->counters + 5, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    if (n) *n = ctx->h_counters[5];
    return KV_OK;
}

extern "C" int kv_launch_count(int device, uint64_t *n)
{
    KvCtx *ctx;
    KV_TRY(kv_ctx_get(device, &ctx));
    if (n) *n = ctx->launches;
    return KV_OK;
}
