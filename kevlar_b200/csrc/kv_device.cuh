// kv_device.cuh -- device-side building blocks shared by the kvsketch kernels (sm_100a).
//
// Arithmetic follows SURVEY.md Appendix A (khmer's behaviour as pinned by the reference's
// golden sketches); every function is exercised bit-for-bit against oracle/kmer_oracle.c by
// tests/test_gpu_parity.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kvsketch.h"

#define KV_TABLES_DEV 8          // tables per sketch view carried in kernel parameters
#define KV_MAX_RANKS 16          // ranks of one device-side barrier (one NVSwitch domain)

// ------------------------------------------------------------------ sketch view

struct KvView {
    uint8_t *tab[KV_TABLES_DEV];   // khmer layout: u8[p] | nibbles (even bin = high) | bits (LSB first)
    uint64_t size[KV_TABLES_DEV];  // buckets of table t held HERE (all of them, or this shard's bin range)
    uint64_t msize[KV_TABLES_DEV]; // the table's full size (the prime): bin = hash % msize
    uint64_t magic[KV_TABLES_DEV]; // floor((2^64-1)/msize) for the Barrett reduction below
    uint64_t lo[KV_TABLES_DEV];    // first bin held here (0 unless the sketch is bin-range sharded)
    uint32_t *occ[KV_TABLES_DEV];  // 1 bit per bucket: counter != 0 (8/4-bit sketches; NULL for bit tables)
    uint32_t *hotf;                // "maybe hot" bitmap: one bit per 8 ADJACENT buckets of a table
    uint64_t hot_base[KV_TABLES_DEV];   // first bit of table t in hotf
    int n_tables;
    int bits;                      // 8, 4 or 1
};

// Side information kept in SEPARATE small arrays so that the update path never has to load a
// counter line before hitting it with an atomic (a plain load followed by an atomic on the same
// line costs ~3x the atomic alone on B200, profiles/r01_atomic_microbench_variants.csv):
//   occ[t]  occupancy bitmap, rebuilt from the counters by one streaming pass at the start of each
//           batch's n_unique bookkeeping ("was this bucket empty when the batch started?");
//   hotf    one bit per group of 8 adjacent buckets; a set bit means some bucket of the group MAY
//           hold a counter at or above KV_HOT (128 / 8), and the whole group takes the exact
//           compare-and-swap path.  1/8 bit per bucket keeps it cache-resident, false positives
//           only cost speed, and -- unlike a hashed filter -- it is as local as the buckets are,
//           which the region-partitioned update path relies on.
// Bits only ever get set (atomicOr), so a stale read is always on the safe side.
template <int BITS>
__device__ __forceinline__ unsigned kv_hot_threshold() { return BITS == 8 ? 128u : 8u; }

__device__ __forceinline__ uint64_t kv_hot_index(const KvView &v, int t, uint64_t bin)
{
    return v.hot_base[t] + (bin >> 3);
}

__device__ __forceinline__ bool kv_maybe_hot(const KvView &v, int t, uint64_t bin)
{
    uint64_t i = kv_hot_index(v, t, bin);
    return (*(volatile uint32_t *)(v.hotf + (i >> 5)) >> (i & 31)) & 1u;
}

__device__ __forceinline__ void kv_mark_hot(const KvView &v, int t, uint64_t bin)
{
    uint64_t i = kv_hot_index(v, t, bin);
    atomicOr(v.hotf + (i >> 5), 1u << (i & 31));
}

// is the bucket empty?  (occupancy bitmap for counters, the table itself for bit tables)
__device__ __forceinline__ bool kv_bucket_empty(const KvView &v, int t, uint64_t bin)
{
    if (v.bits == 1) return !((__ldg(v.tab[t] + (bin >> 3)) >> (bin & 7)) & 1u);
    return !((__ldg(v.occ[t] + (bin >> 5)) >> (bin & 31)) & 1u);
}

// h mod p, identical to C's `%` on uint64 (khmer: bin = hash % tablesize).
// q = mulhi(h, floor((2^64-1)/p)) underestimates floor(h/p) by at most 2, so at most two
// corrective subtractions; no 128-bit division on the device.
__device__ __forceinline__ uint64_t kv_mod(uint64_t h, uint64_t p, uint64_t magic)
{
    uint64_t q = __umul64hi(h, magic);
    uint64_t r = h - q * p;
    if (r >= p) r -= p;
    if (r >= p) r -= p;
    return r;
}

// Bucket of hash h in table t, as an index into the storage held here.  Returns false when the
// bucket belongs to another shard (bin-range sharded sketches, SURVEY 8e plan B); always true
// for an ordinary sketch.
__device__ __forceinline__ bool kv_bin(const KvView &v, int t, uint64_t h, uint64_t &bin)
{
    bin = kv_mod(h, v.msize[t], v.magic[t]) - v.lo[t];
    return bin < v.size[t];
}

// One counter byte for a lookup.  A lookup is a random 1-byte load; on sketches that do not fit L2 almost every one
// misses, and by default a miss brings a whole 128-byte line in from HBM (ncu on 4 GB sketches: 123 B of DRAM reads per
// load, DRAM 80 % busy in kv_novel_kernel).  The L2::64B qualifier asks for the smallest fill the hardware does.
#ifndef KV_COUNTER_FILL
#define KV_COUNTER_FILL 64   // 0: plain __ldg
#endif
__device__ __forceinline__ unsigned kv_ld_counter(const uint8_t *p)
{
#if KV_COUNTER_FILL == 64
    unsigned x;
    asm("ld.global.nc.L2::64B.u8 %0, [%1];" : "=r"(x) : "l"(p));
    return x;
#else
    return __ldg(p);
#endif
}

// Counter read for one table (khmer Storage::get_count inner step, App. A.4).
__device__ __forceinline__ unsigned kv_bucket_get(const KvView &v, int t, uint64_t bin)
{
    if (v.bits == 8) return kv_ld_counter(v.tab[t] + bin);
    if (v.bits == 4) {
        unsigned b = kv_ld_counter(v.tab[t] + (bin >> 1));
        return (b >> ((bin & 1) ? 0 : 4)) & 15u;
    }
    unsigned b = kv_ld_counter(v.tab[t] + (bin >> 3));
    return (b >> (bin & 7)) & 1u;
}

// min over tables (Counttable.get; kevlar/novel.py:38,48)
__device__ __forceinline__ unsigned kv_get(const KvView &v, uint64_t h)
{
    unsigned m = 0xffffffffu;
#pragma unroll 4
    for (int t = 0; t < v.n_tables; t++) {
        uint64_t bin;
        if (!kv_bin(v, t, h, bin)) continue;   // another shard answers for this table
        unsigned c = kv_bucket_get(v, t, bin);
        m = c < m ? c : m;
    }
    return m;   // 0xffffffff (255 as a byte) if no table of this k-mer lives here
}

// Saturating increment of one bucket.  CUDA has no 8-/4-bit atomics and a plain 32-bit
// atomicAdd would carry into the neighbouring counter at 255/15, so the byte/nibble is
// updated with a 32-bit compare-and-swap on the containing word; 1-bit tables use a
// fire-and-forget atomicOr.  Returns the bucket value seen before the update.
template <int BITS>
__device__ __forceinline__ void kv_word_addr(const KvView &v, int t, uint64_t bin, unsigned *&word, unsigned &shift)
{
    uint64_t byte = BITS == 8 ? bin : (BITS == 4 ? (bin >> 1) : (bin >> 3));
    uint8_t *p = v.tab[t] + byte;
    word = (unsigned *)((uintptr_t)p & ~(uintptr_t)3);
    shift = (unsigned)((uintptr_t)p & 3) * 8;
    if (BITS == 4) shift += (bin & 1) ? 0 : 4;
    if (BITS == 1) shift += (unsigned)(bin & 7);
}

// Exact saturating update of one bucket: atomic read of the containing word (an atomic, not a
// load -- see above), then 32-bit compare-and-swap until the byte/nibble has been bumped or is
// saturated.  Returns true if this call added one; *seen_old gets the counter value it replaced.
template <int BITS>
__device__ __forceinline__ bool kv_sat_inc_exact(unsigned *word, unsigned shift, unsigned &seen_old)
{
    const unsigned maxv = BITS == 8 ? 255u : 15u;
    unsigned old = atomicOr(word, 0u);
    while (((old >> shift) & maxv) != maxv) {
        unsigned assumed = old;
        old = atomicCAS(word, assumed, assumed + (1u << shift));
        if (old == assumed) { seen_old = (old >> shift) & maxv; return true; }
    }
    seen_old = maxv;
    return false;
}

// after an update that replaced the value `ob`: a counter that reaches the hot threshold routes
// its group of buckets to the exact path from now on
template <int BITS>
__device__ __forceinline__ void kv_state_publish(const KvView &v, int t, uint64_t bin, unsigned ob)
{
    if (ob + 1 >= kv_hot_threshold<BITS>()) kv_mark_hot(v, t, bin);
}

// ------------------------------------------------------------------ MurmurHash3

__device__ __forceinline__ uint64_t kv_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

__device__ __forceinline__ uint64_t kv_fmix64(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// Low 64 bits of MurmurHash3_x64_128(seed 0) over the first k bytes held little-endian in
// w[0..KW) (bytes >= k MUST be zero).  KW = 4*ceil(k/16) words so all loops unroll.
template <int KW>
__device__ __forceinline__ uint64_t kv_murmur_words(const uint32_t (&w)[KW], int k)
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = 0, h2 = 0;
    const int nblocks = k >> 4;
    uint64_t t1 = 0, t2 = 0;   // tail words
#pragma unroll
    for (int b = 0; b < KW / 4; b++) {
        uint64_t k1 = (uint64_t)w[4 * b] | ((uint64_t)w[4 * b + 1] << 32);
        uint64_t k2 = (uint64_t)w[4 * b + 2] | ((uint64_t)w[4 * b + 3] << 32);
        if (b < nblocks) {
            k1 *= c1; k1 = kv_rotl64(k1, 31); k1 *= c2; h1 ^= k1;
            h1 = kv_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
            k2 *= c2; k2 = kv_rotl64(k2, 33); k2 *= c1; h2 ^= k2;
            h2 = kv_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
        } else if (b == nblocks) {
            t1 = k1; t2 = k2;
        }
    }
    const int rem = k & 15;
    if (rem > 8) { t2 *= c2; t2 = kv_rotl64(t2, 33); t2 *= c1; h2 ^= t2; }
    if (rem > 0) { t1 *= c1; t1 = kv_rotl64(t1, 31); t1 *= c2; h1 ^= t1; }
    h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
    h1 += h2; h2 += h1;
    h1 = kv_fmix64(h1); h2 = kv_fmix64(h2);
    return h1 + h2;
}

// complement of 4 packed upper-case ASCII bases: A<->T differ by 0x15, C<->G by 0x04, and
// bit 1 is set exactly for C/G.
__device__ __forceinline__ uint32_t kv_comp4(uint32_t x)
{
    uint32_t cg = (x >> 1) & 0x01010101u;
    return x ^ 0x15151515u ^ (cg | (cg << 4));
}

// ---------------------------------------------------------------- tile in smem
//
// A CTA works on KV_TILE consecutive base positions of the batch.  The cleaned ASCII bytes of
// [tile_start - KV_FRONT, tile_start + KV_TILE + KV_BACK) sit in shared memory so that each
// thread can pull the forward window AND the window ending at its k-mer's last base (for the
// reverse complement) with aligned 32-bit LDS + funnel shifts.

#define KV_TILE 1024
#define KV_THREADS 256
#define KV_FRONT 64
#define KV_BACK 80
#define KV_SM_BYTES (KV_FRONT + KV_TILE + KV_BACK)  // 1168
#define KV_SM_WORDS (KV_SM_BYTES / 4)               // 292
#define KV_MAXB 320                                  // read boundaries cached per tile

struct KvTileSmem {
    uint32_t bytes[KV_SM_WORDS + 4];        // cleaned ASCII
    uint32_t packed[KV_SM_BYTES / 16 + 4];  // 2-bit codes, 16 bases per word, first base in the top bits
    uint32_t bad[KV_SM_BYTES / 32 + 4];     // bit i of word w: byte 32w+i is outside ACGT (strict mode)
    uint64_t ends[KV_MAXB];                 // end offsets of the reads that intersect the tile
    int nb;                                 // number of cached ends, or -1: search global memory
    uint64_t r0;                            // index of the first read intersecting the tile
};

// Upper-case acgt, map everything outside ACGT to 'A' (khmer Read::set_clean_seq, App. A.6).
// strict_bad gets one bit per byte (bit 0 of each byte lane) where the ORIGINAL byte is not
// one of upper-case ACGT (kevlar/novel.py:136: re.search('[^ACGT]')).
__device__ __forceinline__ uint32_t kv_clean4(uint32_t x, uint32_t &strict_bad)
{
    uint32_t ok_strict = __vcmpeq4(x, 0x41414141u) | __vcmpeq4(x, 0x43434343u) |
                         __vcmpeq4(x, 0x47474747u) | __vcmpeq4(x, 0x54545454u);
    strict_bad = ~ok_strict & 0x01010101u;
    uint32_t up = x & 0xDFDFDFDFu;
    // lower-case letters differ from upper-case only in bit 5; anything else that maps onto
    // ACGT after clearing bit 5 would be e.g. 'a'..'t' only, since A/C/G/T +0x20 are the sole preimages
    uint32_t ok = __vcmpeq4(up, 0x41414141u) | __vcmpeq4(up, 0x43434343u) |
                  __vcmpeq4(up, 0x47474747u) | __vcmpeq4(up, 0x54545454u);
    return (up & ok) | (0x41414141u & ~ok);
}

// 4 cleaned ASCII bases (little-endian word, first base in the low byte) -> 8 bits of 2-bit
// codes, first base in the top two bits.  khmer code: A=0, T=1, C=2, G=3 = (bit1, bit2) of the
// ASCII byte.
__device__ __forceinline__ uint32_t kv_pack4(uint32_t x)
{
    uint32_t t = (((x >> 1) & 0x01010101u) << 1) | ((x >> 2) & 0x01010101u);
    return (t * 0x40100401u) >> 24;
}

// Load + clean the tile.  `total` = number of bases in the batch.  All threads run the same
// number of iterations (the warp shuffles below need full warps).
template <bool NEED_PACKED, bool NEED_BAD>
__device__ __forceinline__ void kv_tile_load(KvTileSmem &sm, const uint8_t *__restrict__ bases, uint64_t tile_start,
                                             uint64_t total)
{
    const uint32_t *gw = (const uint32_t *)bases;   // batch buffers are at least 4-byte aligned
    const int64_t w0 = (int64_t)(tile_start / 4) - KV_FRONT / 4;
    const int64_t wfull = (int64_t)(total / 4);     // words that lie completely inside the batch
    constexpr int NW = KV_SM_WORDS + 4;
    for (int base = 0; base < NW; base += KV_THREADS) {
        const int i = base + threadIdx.x;
        const int64_t gi = w0 + i;
        uint32_t x = 0;
        if (i < KV_SM_WORDS && gi >= 0) {
            if (gi < wfull) x = __ldg(gw + gi);
            else if (gi == wfull)   // ragged tail: never read past the caller's buffer
                for (uint64_t b = 0; b < (total & 3); b++) x |= (uint32_t)__ldg(bases + 4 * gi + b) << (8 * b);
        }
        uint32_t bad;
        uint32_t c = kv_clean4(x, bad);
        if (i < NW) sm.bytes[i] = c;
        if (NEED_PACKED) {
            // each lane packs 4 bases into 8 bits; 4 consecutive lanes make one 16-base word
            uint32_t pk = kv_pack4(c) << (24 - 8 * (i & 3));
            pk |= __shfl_xor_sync(0xffffffffu, pk, 1);
            pk |= __shfl_xor_sync(0xffffffffu, pk, 2);
            if ((i & 3) == 0 && i < NW) sm.packed[i >> 2] = pk;
        }
        if (NEED_BAD) {
            // 4 bad-bits per lane; 8 consecutive lanes make one 32-byte mask word
            uint32_t nib = ((bad & 1u) | ((bad >> 7) & 2u) | ((bad >> 14) & 4u) | ((bad >> 21) & 8u));
            uint32_t bw = nib << (4 * (i & 7));
            bw |= __shfl_xor_sync(0xffffffffu, bw, 1);
            bw |= __shfl_xor_sync(0xffffffffu, bw, 2);
            bw |= __shfl_xor_sync(0xffffffffu, bw, 4);
            if ((i & 7) == 0 && i < NW) sm.bad[i >> 3] = bw;
        }
    }
}

// Cache the end offsets of the reads intersecting this tile (tile_first[] comes from
// kv_tile_index_kernel).
__device__ __forceinline__ void kv_tile_bounds(KvTileSmem &sm, const uint64_t *__restrict__ offsets,
                                               const uint32_t *__restrict__ tile_first, uint64_t tile)
{
    uint64_t r0 = tile_first[tile], r1 = tile_first[tile + 1];
    uint64_t nb = r1 - r0 + 1;
    if (threadIdx.x == 0) { sm.r0 = r0; sm.nb = nb <= KV_MAXB ? (int)nb : -1; }
    if (nb <= KV_MAXB)
        for (int j = threadIdx.x; j < (int)nb; j += KV_THREADS) sm.ends[j] = __ldg(offsets + r0 + 1 + j);
}

// Which read does base position g belong to?  Returns false if g starts no k-mer.
// read = index in batch, rstart/rend = its offsets.
__device__ __forceinline__ void kv_find_read(const KvTileSmem &sm, const uint64_t *__restrict__ offsets,
                                             const uint32_t *__restrict__ tile_first, uint64_t tile, uint64_t g,
                                             uint64_t &read, uint64_t &rstart, uint64_t &rend)
{
    if (sm.nb >= 0) {
        int lo = 0, hi = sm.nb - 1;   // smallest j with ends[j] > g (exists for g < total)
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (sm.ends[mid] > g) hi = mid; else lo = mid + 1;
        }
        read = sm.r0 + lo;
        rend = sm.ends[lo];
        rstart = lo ? sm.ends[lo - 1] : __ldg(offsets + sm.r0);
    } else {
        uint64_t lo = tile_first[tile], hi = tile_first[tile + 1];
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if (__ldg(offsets + mid + 1) > g) hi = mid; else lo = mid + 1;
        }
        read = lo;
        rend = __ldg(offsets + lo + 1);
        rstart = __ldg(offsets + lo);
    }
}

// true if any byte of the k-mer starting at smem byte index L is flagged in sm.bad
__device__ __forceinline__ bool kv_window_bad(const KvTileSmem &sm, int L, int k)
{
    int j = L >> 5, o = L & 31;
    uint32_t lo = __funnelshift_r(sm.bad[j], sm.bad[j + 1], o);
    uint32_t hi = __funnelshift_r(sm.bad[j + 1], sm.bad[j + 2], o);
    if (k <= 32) return (lo & (k == 32 ? 0xffffffffu : ((1u << k) - 1u))) != 0;
    return (lo | (hi & (k == 64 ? 0xffffffffu : ((1u << (k - 32)) - 1u)))) != 0;
}

// The tile positions that start a k-mer (inside one read; in strict mode also free of bytes
// outside ACGT), as an ascending list in shared memory.  With 100 bp reads and k = 31 that is 70 %
// of the positions, in runs that never leave a whole warp idle, so the kernels do their expensive
// per-k-mer work over this list instead of over all positions.  Built from the read boundaries:
// start from "every existing position", then each read end e clears [e-k+1, e) -- O(reads in the
// tile) work instead of a boundary search per position.  Returns the list length; ends with a
// __syncthreads().
struct KvTileList {
    uint32_t bits[KV_TILE / 32];        // bit l: position l of the tile starts a k-mer
    unsigned base[KV_TILE / 32 + 1];    // exclusive prefix sums of popc(bits[]), [32] = total
    uint16_t pos[KV_TILE];
};

__device__ __forceinline__ unsigned kv_tile_kmer_list(const KvTileSmem &sm, KvTileList &ls, const uint64_t *__restrict__ offsets,
                                                      const uint32_t *__restrict__ tile_first, uint64_t tile, uint64_t total,
                                                      int k, bool strict)
{
    const uint64_t tile_start = tile * KV_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (sm.nb >= 0) {
        if (threadIdx.x < KV_TILE / 32) {
            uint64_t first = tile_start + 32u * threadIdx.x;
            ls.bits[threadIdx.x] = first >= total ? 0u : (total - first >= 32 ? 0xffffffffu : ((1u << (total - first)) - 1u));
        }
        __syncthreads();
        for (int j = threadIdx.x; j < sm.nb; j += KV_THREADS) {
            const uint64_t e = sm.ends[j];
            uint64_t lo = e >= (uint64_t)(k - 1) ? e - (uint64_t)(k - 1) : 0, hi = e;
            if (lo < tile_start) lo = tile_start;
            if (hi > tile_start + KV_TILE) hi = tile_start + KV_TILE;
            if (lo >= hi) continue;
            const int a = (int)(lo - tile_start), b = (int)(hi - tile_start);   // clear bits [a, b)
            for (int w = a >> 5; w <= (b - 1) >> 5; w++) {
                uint32_t m = 0xffffffffu;
                if (w == (a >> 5)) m &= 0xffffffffu << (a & 31);
                if (w == ((b - 1) >> 5) && (b & 31)) m &= (1u << (b & 31)) - 1u;
                atomicAnd(&ls.bits[w], ~m);
            }
        }
    } else {   // more reads in the tile than the boundary cache holds: search per position
        for (int it = 0; it < KV_TILE / KV_THREADS; it++) {
            const int l = it * KV_THREADS + threadIdx.x;
            const uint64_t g = tile_start + l;
            bool ok = false;
            if (g < total) {
                uint64_t read, rs, re;
                kv_find_read(sm, offsets, tile_first, tile, g, read, rs, re);
                ok = g + k <= re;
            }
            unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) ls.bits[l >> 5] = bal;
        }
    }
    __syncthreads();
    if (strict) {
        for (int it = 0; it < KV_TILE / KV_THREADS; it++) {
            const int l = it * KV_THREADS + threadIdx.x;
            bool drop = ((ls.bits[l >> 5] >> (l & 31)) & 1u) && kv_window_bad(sm, l + KV_FRONT, k);
            unsigned bal = __ballot_sync(0xffffffffu, drop);
            if (lane == 0 && bal) ls.bits[l >> 5] &= ~bal;   // word l>>5 belongs to this warp in this iteration
        }
        __syncthreads();
    }
    if (warp == 0) {
        unsigned c = __popc(ls.bits[lane]), x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        ls.base[lane] = x - c;
        if (lane == 31) ls.base[32] = x;
    }
    __syncthreads();
    for (int it = 0; it < KV_TILE / KV_THREADS; it++) {
        const int l = it * KV_THREADS + threadIdx.x;
        const uint32_t w = ls.bits[l >> 5];
        if ((w >> lane) & 1u) ls.pos[ls.base[l >> 5] + __popc(w & ((1u << lane) - 1u))] = (uint16_t)l;
    }
    __syncthreads();
    return ls.base[32];
}

// Canonical MurmurHash of the k-mer at local position l (khmer _hash_murmur: forward XOR
// reverse complement, App. A.2).
template <int KW>
__device__ __forceinline__ uint64_t kv_tile_hash_murmur(const KvTileSmem &sm, int l, int k)
{
    uint32_t f[KW], r[KW];
    {   // forward: bytes [L, L+k)
        const int L = l + KV_FRONT;
        const int j = L >> 2, sh = (L & 3) * 8;
        uint32_t prev = sm.bytes[j];
#pragma unroll
        for (int i = 0; i < KW; i++) {
            uint32_t nxt = sm.bytes[j + i + 1];
            f[i] = __funnelshift_r(prev, nxt, sh);
            prev = nxt;
        }
    }
    {   // reverse complement: take the 4*KW bytes ENDING at the k-mer's last base, reverse them
        const int E = l + KV_FRONT + k - 4 * KW;   // >= 1 because KV_FRONT >= 4*KW_max
        const int j = E >> 2, sh = (E & 3) * 8;
        uint32_t v[KW];
        uint32_t prev = sm.bytes[j];
#pragma unroll
        for (int i = 0; i < KW; i++) {
            uint32_t nxt = sm.bytes[j + i + 1];
            v[i] = __funnelshift_r(prev, nxt, sh);
            prev = nxt;
        }
#pragma unroll
        for (int i = 0; i < KW; i++) r[i] = kv_comp4(__byte_perm(v[KW - 1 - i], 0, 0x0123));
    }
    // zero the bytes at and beyond k
    const int kw = k >> 2;
    const uint32_t part = (k & 3) ? ((1u << (8 * (k & 3))) - 1u) : 0u;
#pragma unroll
    for (int i = 0; i < KW; i++) {
        uint32_t m = i < kw ? 0xffffffffu : (i == kw ? part : 0u);
        f[i] &= m;
        r[i] &= m;
    }
    return kv_murmur_words<KW>(f, k) ^ kv_murmur_words<KW>(r, k);
}

// Canonical 2-bit hash of the k-mer at local position l (khmer _hash + uniqify_rc, App. A.3):
// a 2k-bit window cut out of the packed tile with two funnel shifts; the reverse complement
// is the pair-reversed complement of the same window -- no rolling dependency between lanes.
__device__ __forceinline__ uint64_t kv_tile_hash_twobit(const KvTileSmem &sm, int l, int k)
{
    const int L = l + KV_FRONT;
    const int j = L >> 4, o = 2 * (L & 15);
    uint32_t p0 = sm.packed[j], p1 = sm.packed[j + 1], p2 = sm.packed[j + 2];
    uint32_t a = __funnelshift_l(p1, p0, o);
    uint32_t b = __funnelshift_l(p2, p1, o);
    uint64_t v = ((uint64_t)a << 32) | b;            // 32 bases starting at l, first base on top
    uint64_t fwd = v >> (64 - 2 * k);
    uint64_t c = __brevll(v ^ 0x5555555555555555ULL); // complement (code^1), then reverse bit order
    c = ((c & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((c & 0x5555555555555555ULL) << 1);
    uint64_t rc = k == 32 ? c : (c & ((1ULL << (2 * k)) - 1ULL));
    return fwd < rc ? fwd : rc;
}

template <int HASHER, int KW>
__device__ __forceinline__ uint64_t kv_tile_hash(const KvTileSmem &sm, int l, int k)
{
    if (HASHER == KV_HASH_TWOBIT) return kv_tile_hash_twobit(sm, l, k);
    return kv_tile_hash_murmur<KW>(sm, l, k);
}
