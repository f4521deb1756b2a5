// kv_kernels.cuh -- the sm_100a kernels behind libkvsketch.so.
//
//   K0  kv_tile_index_kernel    which read does each 1024-base tile start in
//   K1/K2 kv_hash_kernel        clean + (2-bit pack) + canonical hash per base position,
//                               band and mask predicates fused; writes hashes + valid bits
//   K3  kv_increment_kernel     saturating 8/4/1-bit Count-Min / Bloom update (khmer Storage::add):
//                               speculative ATOM.ADD for cold buckets, exact CAS for hot ones;
//       kv_rollback_kernel      undo of a chunk whose speculative pass overflowed a counter
//   K3b kv_part_*_kernel        the same update for sketches larger than L2: updates partitioned
//                               by (table, 16 MB region), then applied region by region
//   K4  kv_novel_kernel         fused hash + case/control lookups + thresholds
//                               (kevlar/novel.py:21-53,123-169)
//   K5  kv_occ_rebuild / kv_first_min / kv_first_resolve / kv_popcount   exact n_unique_kmers;
//       kv_abund_dist_kernel    abundance histogram of first-seen k-mers (kevlar dist)
//       kv_occupied_kernel      n_occupied
//   K6  kv_widen / kv_narrow / kv_merge_peers kernels   multi-GPU saturating merge
//   +   kv_get_kernel, kv_gather_kernel, kv_expand_bits_kernel, kv_state_rebuild_kernel   helpers
#pragma once
#include <cuda_fp16.h>

#include "kv_device.cuh"

#ifndef KV_INC_MIN_CTAS
#define KV_INC_MIN_CTAS 6   // resident CTAs per SM the increment kernel is compiled for (register budget)
#endif

// ----------------------------------------------------------------------- K0

// tile_first[i] = min(first read r with offsets[r+1] > i*KV_TILE, n_reads-1), i in [0, n_tiles]
__global__ void kv_tile_index_kernel(const uint64_t *__restrict__ offsets, uint64_t n_reads, uint64_t n_tiles,
                                     uint32_t *__restrict__ tile_first)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_tiles) return;
    uint64_t pos = i * KV_TILE;
    uint64_t lo = 0, hi = n_reads - 1;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid + 1) > pos) hi = mid; else lo = mid + 1;
    }
    tile_first[i] = (uint32_t)lo;
}

// -------------------------------------------------------------------- K1/K2

// K3c bookkeeping shared by the hash kernel (producer) and kv_tile_apply_kernel (consumer): table t is
// cut into regions of 2^rb buckets; run = run_base[t] + (bin >> rb) owns slab[run * cap .. + cap) and
// the cursor that hands out its slots.
struct KvTileInfo {
    int rb;                              // log2(buckets per region), <= 16 so offsets fit 16 bits
    uint32_t run_base[KV_TABLES_DEV + 1];   // first run of table t; [n_tables] = number of runs
    uint32_t cap;                        // slots per run in this chunk (a multiple of the block size)
    uint32_t *cursor;                    // [runs] slots requested so far (may exceed cap)
    uint16_t *slab;                      // runs * cap offsets, laid out by kv_slab_index
    int blk_log2;                        // >= 0: block-interleaved slabs, 2^blk_log2 slots per block; < 0: run-major
    // overflow (a run asked for more than cap slots: skewed input).  Untracked sketches are updated in
    // place by the producer; with exact n_unique tracking the tables must stay untouched until the
    // first-touch passes have run, so the producer only marks (position, table) here and
    // kv_tile_overflow_kernel applies the marks afterwards.
    uint32_t *ovf;                       // NULL: update in place; else [n_tables][ovf_stride] bitmaps over chunk positions
    uint64_t ovf_stride;
    unsigned *ovf_any;
};

// Where slot `slot` of run `run` lives.  Hashing spreads the updates evenly, so all runs fill at the same
// pace: with block-interleaved slabs ([block][run][2^blk_log2 slots]) the write frontier of ALL runs is
// one contiguous stretch of runs * 2^blk_log2 * 2 bytes that moves through the array, instead of one
// open sector per run spread over the whole allocation (gigabytes, thousands of pages).
__device__ __forceinline__ size_t kv_slab_index(const KvTileInfo &ti, uint32_t run, uint32_t slot)
{
    if (ti.blk_log2 < 0) return (size_t)run * ti.cap + slot;
    const uint32_t n_runs = ti.run_base[KV_TABLES_DEV];   // mirrored there by the host: total number of runs
    return ((((size_t)(slot >> ti.blk_log2) * n_runs + run) << ti.blk_log2) | (slot & ((1u << ti.blk_log2) - 1u)));
}

// exact saturating update of one bucket in global memory, counter width chosen at run time (the
// overflow path of K3c and its sparse regions: rare, so no template instance per width)
__device__ __forceinline__ void kv_bucket_inc_exact(const KvView &v, int t, uint64_t bin)
{
    unsigned *w, sh, seen;
    if (v.bits == 8) { kv_word_addr<8>(v, t, bin, w, sh); kv_sat_inc_exact<8>(w, sh, seen); }
    else if (v.bits == 4) { kv_word_addr<4>(v, t, bin, w, sh); kv_sat_inc_exact<4>(w, sh, seen); }
    else { kv_word_addr<1>(v, t, bin, w, sh); atomicOr(w, 1u << sh); }
}

// file the updates of hash h for tables t0 .. t0+3 (K3c producer)
__device__ __forceinline__ void kv_scatter4(const KvTileInfo &ti, const KvView &sk, int t0, uint64_t h, uint64_t pos)
{
    uint64_t bin[4];
    uint32_t run[4], slot[4];
    bool own[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        own[j] = t0 + j < sk.n_tables && kv_bin(sk, t0 + j, h, bin[j]);
        if (own[j]) run[j] = ti.run_base[t0 + j] + (uint32_t)(bin[j] >> ti.rb);
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (own[j]) slot[j] = atomicAdd(ti.cursor + run[j], 1u);
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (own[j]) {
            if (slot[j] < ti.cap)
                ti.slab[kv_slab_index(ti, run[j], slot[j])] = (uint16_t)(bin[j] & ((1u << ti.rb) - 1u));
            else if (ti.ovf) {   // slab full (skewed input), tables frozen until the n_unique passes are done
                atomicOr(ti.ovf + (size_t)(t0 + j) * ti.ovf_stride + (pos >> 5), 1u << (pos & 31));
                *ti.ovf_any = 1u;
            } else
                kv_bucket_inc_exact(sk, t0 + j, bin[j]);   // slab full: update in place, exactly
        }
}

struct KvHashParams {
    const uint8_t *bases;
    const uint64_t *offsets;
    const uint32_t *tile_first;
    uint64_t total;        // bases in the batch
    uint64_t tile0;        // first tile of this chunk; outputs are indexed relative to tile0*KV_TILE
    int k;
    int banded;            // khmer range banding (App. A.7): keep lo <= h < hi
    uint64_t band_lo, band_hi;
    int use_mask;          // App. A.8
    int mask_threshold, consume_masked;
    KvView mask;
    int strict;            // 1: k-mers touching a byte outside ACGT are invalid (kv_hash_kmers ok[])
    uint64_t *hashes;      // [chunk positions] (may be NULL)
    uint32_t *valid;       // [chunk positions / 32] bit: position starts a k-mer that passed the filters
    unsigned long long *n_valid;   // running count of valid k-mers (khmer n_consumed)
    // K3c (tiled update path): the kernel also files every update of the sketch `sk` under its
    // (table, region) run, as a 16-bit bucket offset inside the region
    int scatter;
    KvTileInfo ti;
    KvView sk;             // the sketch being updated (scatter and/or track0)
    // K5, pass A of table 0 fused in: first0[bin0] = min(first0[bin0], tag0 | position) where the bucket is empty
    int track0;
    uint32_t *first0;
    uint32_t tag0;
};

#ifndef KV_HASH_MIN_CTAS
#define KV_HASH_MIN_CTAS 6      // resident CTAs per SM the plain hash kernel is compiled for
#endif
#ifndef KV_SCATTER_MIN_CTAS
#define KV_SCATTER_MIN_CTAS 4   // resident CTAs per SM the scatter variant is compiled for
#endif
template <int HASHER, int KW, bool SCATTER>
__global__ void __launch_bounds__(KV_THREADS, SCATTER ? KV_SCATTER_MIN_CTAS : KV_HASH_MIN_CTAS) kv_hash_kernel(const __grid_constant__ KvHashParams p)
{
    __shared__ KvTileSmem sm;
    __shared__ KvTileList ls;
    __shared__ uint32_t s_valid[KV_TILE / 32];
    const uint64_t tile = p.tile0 + blockIdx.x;
    const uint64_t tile_start = tile * KV_TILE;
    const uint64_t pos0 = p.tile0 * KV_TILE;
    kv_tile_load<HASHER == KV_HASH_TWOBIT, true>(sm, p.bases, tile_start, p.total);
    kv_tile_bounds(sm, p.offsets, p.tile_first, tile);
    if (threadIdx.x < KV_TILE / 32) s_valid[threadIdx.x] = 0;
    __syncthreads();
    const unsigned n_list = kv_tile_kmer_list(sm, ls, p.offsets, p.tile_first, tile, p.total, p.k, p.strict != 0);

    // positions that start no k-mer get hash 0 (the output stays deterministic)
    if (p.hashes)
        for (int it = 0; it < KV_TILE / KV_THREADS; it++) {
            const int l = it * KV_THREADS + threadIdx.x;
            if (tile_start + l < p.total && !((ls.bits[l >> 5] >> (l & 31)) & 1u)) p.hashes[tile_start + l - pos0] = 0;
        }

    // hash + band / mask predicates over the k-mer list
    unsigned n_ok = 0;
#pragma unroll 1
    for (unsigned i = threadIdx.x; i < n_list; i += KV_THREADS) {
        const int l = ls.pos[i];
        uint64_t h = kv_tile_hash<HASHER, KW>(sm, l, p.k);
        bool ok = true;
        if (p.banded) ok = h >= p.band_lo && h < p.band_hi;
        if (ok && p.use_mask) {
            int c = (int)kv_get(p.mask, h);
            ok = p.consume_masked ? (c >= p.mask_threshold) : (c <= p.mask_threshold);
        }
        if (p.hashes) p.hashes[tile_start + l - pos0] = h;
        if (ok) atomicOr(&s_valid[l >> 5], 1u << (l & 31));
        n_ok += ok;
        if (ok && p.track0) {
            uint64_t bin0;
            if (kv_bin(p.sk, 0, h, bin0) && kv_bucket_empty(p.sk, 0, bin0))
                atomicMin(p.first0 + bin0, p.tag0 | (uint32_t)(tile_start + l - pos0));
        }
        if (SCATTER && ok) {
            // K3c: one cursor atomic + one 2-byte store per table (kevlar always builds 4 tables; more are
            // handled four at a time)
            const uint64_t pos = tile_start + l - pos0;
            for (int t0 = 0; t0 < p.sk.n_tables; t0 += 4) kv_scatter4(p.ti, p.sk, t0, h, pos);
        }
    }
    __syncthreads();
    if (threadIdx.x < KV_TILE / 32 && tile_start + threadIdx.x * 32 < p.total)
        p.valid[(tile_start - pos0) / 32 + threadIdx.x] = s_valid[threadIdx.x];
    // one atomic per CTA for the k-mer count
    n_ok = __reduce_add_sync(0xffffffffu, n_ok);
    __shared__ unsigned s_total;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && n_ok) atomicAdd(&s_total, n_ok);
    __syncthreads();
    if (threadIdx.x == 0 && s_total) atomicAdd(p.n_valid, (unsigned long long)s_total);
}

// ----------------------------------------------------------------------- K3
//
// Saturating counter update, one thread per base position (grid-stride).  For every valid k-mer
// and each of its T tables:
//   cold bucket (state != hot): ONE speculative 32-bit ATOM.ADD on the containing word.  Its return
//       value tells what was replaced: 0 -> publish "occupied"; near the hot threshold -> publish
//       "hot"; the maximum -> the add carried into the neighbouring counter: raise the chunk's
//       dirty flag (only possible if >= 128 adds to one bucket were in flight together).
//   hot bucket: exact path (atomic read + compare-and-swap loop, stops at the maximum).
//   bit tables: fire-and-forget RED.OR.
// Every add that actually changed memory is recorded (one bit per position and table), so a
// dirty chunk can be undone arithmetically -- 32-bit adds commute, whatever transient carries
// happened -- and redone with the exact path by the two follow-up kernels below, which exit
// immediately when the flag is clear.
template <int BITS, bool HAS_VALID, bool EXACT>
__global__ void __launch_bounds__(256, KV_INC_MIN_CTAS) kv_increment_kernel(KvView v, const uint64_t *__restrict__ hashes,
                                                           const uint32_t *__restrict__ valid, uint64_t total,
                                                           uint32_t *__restrict__ added, uint64_t added_stride,
                                                           unsigned *dirty, unsigned long long *n_redone)
{
    if (EXACT && *dirty == 0) return;   // redo pass of a clean chunk
    if (EXACT && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(n_redone, 1ULL);
    const unsigned maxv = BITS == 8 ? 255u : 15u;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t total_pad = (total + 31) & ~(uint64_t)31;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_pad; g += stride) {
        bool live = g < total;
        if (HAS_VALID && live) live = (__ldg(valid + (g >> 5)) >> (g & 31)) & 1u;
        uint64_t h = 0;
        if (live) h = __ldcs(hashes + g);
        if (BITS != 1 && !EXACT && v.n_tables == 4) {
            // common shape (kevlar always builds 4 tables): addresses and states for all four
            // tables first, then the four speculative adds back to back (independent L2 round
            // trips), only then look at what they returned
            unsigned *w[4], sh[4], ob[4];
            uint64_t bin[4];
            bool did[4], hot[4], own[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                did[t] = false;
                ob[t] = 0;
                hot[t] = true;
                own[t] = live && kv_bin(v, t, h, bin[t]);
                if (own[t]) {
                    kv_word_addr<BITS>(v, t, bin[t], w[t], sh[t]);
                    hot[t] = kv_maybe_hot(v, t, bin[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < 4; t++)
                if (own[t] && !hot[t]) {
                    ob[t] = (atomicAdd(w[t], 1u << sh[t]) >> sh[t]) & maxv;
                    did[t] = true;
                }
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (own[t]) {
                    if (hot[t]) {
                        did[t] = kv_sat_inc_exact<BITS>(w[t], sh[t], ob[t]);   // the group is hot already: nothing to publish
                    } else if (ob[t] == maxv)
                        atomicOr(dirty, 1u);
                    else
                        kv_state_publish<BITS>(v, t, bin[t], ob[t]);
                }
                unsigned bal = __ballot_sync(0xffffffffu, did[t]);
                if ((threadIdx.x & 31) == 0) added[t * added_stride + (g >> 5)] = bal;
            }
            continue;
        }
#pragma unroll 4
        for (int t = 0; t < v.n_tables; t++) {
            bool did = false;
            uint64_t bin;
            if (live && kv_bin(v, t, h, bin)) {
                unsigned *w, sh;
                kv_word_addr<BITS>(v, t, bin, w, sh);
                if (BITS == 1) {
                    atomicOr(w, 1u << sh);
                } else {
                    unsigned ob;
                    if (EXACT || kv_maybe_hot(v, t, bin)) {
                        did = kv_sat_inc_exact<BITS>(w, sh, ob);
                        if (did && EXACT) kv_state_publish<BITS>(v, t, bin, ob);   // (a hot group needs no second mark)
                    } else {
                        ob = (atomicAdd(w, 1u << sh) >> sh) & maxv;
                        did = true;
                        if (ob == maxv) atomicOr(dirty, 1u);
                        else kv_state_publish<BITS>(v, t, bin, ob);
                    }
                }
            }
            if (BITS != 1 && !EXACT) {
                unsigned bal = __ballot_sync(0xffffffffu, did);
                if ((threadIdx.x & 31) == 0) added[t * added_stride + (g >> 5)] = bal;
            }
        }
    }
}

// Undo a dirty chunk: subtract exactly what the speculative pass added.
template <int BITS>
__global__ void __launch_bounds__(256) kv_rollback_kernel(KvView v, const uint64_t *__restrict__ hashes, uint64_t total,
                                                          const uint32_t *__restrict__ added, uint64_t added_stride,
                                                          const unsigned *dirty)
{
    if (*dirty == 0) return;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const uint64_t h = hashes[g];
        for (int t = 0; t < v.n_tables; t++) {
            if (!((added[t * added_stride + (g >> 5)] >> (g & 31)) & 1u)) continue;
            unsigned *w, sh;
            uint64_t bin;
            kv_bin(v, t, h, bin);   // a recorded add was an owned bucket
            kv_word_addr<BITS>(v, t, bin, w, sh);
            atomicAdd(w, 0u - (1u << sh));
        }
    }
}

// Recompute the hot bitmap from the counters (after load / merge / raw writes).
template <int BITS>
__global__ void kv_state_rebuild_kernel(KvView v, int t)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t bin = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; bin < v.size[t]; bin += stride) {
        unsigned c = BITS == 8 ? v.tab[t][bin] : ((v.tab[t][bin >> 1] >> ((bin & 1) ? 0 : 4)) & 15u);
        if (c >= kv_hot_threshold<BITS>()) kv_mark_hot(v, t, bin);
    }
}

__global__ void kv_get_kernel(KvView v, const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ valid,
                              uint64_t n, uint8_t *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        bool ok = valid ? ((valid[i >> 5] >> (i & 31)) & 1u) : true;
        out[i] = ok ? (uint8_t)kv_get(v, hashes[i]) : 0;
    }
}

// out[i] = hashes[i*step], ok[i] = valid bit of position i*step (kv_hash_kmers)
__global__ void kv_gather_kernel(const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ valid,
                                 uint64_t n, uint64_t step, uint64_t *__restrict__ out, uint8_t *__restrict__ ok)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t g = i * step;
        out[i] = hashes[g];
        ok[i] = (valid[g >> 5] >> (g & 31)) & 1u;
    }
}

// valid bits -> one byte per position
__global__ void kv_expand_bits_kernel(const uint32_t *__restrict__ valid, uint64_t n, uint8_t *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (valid[i >> 5] >> (i & 31)) & 1u;
}

// ------------------------------------------------------------- K3b: region-partitioned updates
//
// For sketches larger than L2 every counter update is a random DRAM sector read-modify-write
// that also misses the TLB (2 MB pages, 256 MB reach): 13-20 G updates/s however the update is
// written (profiles/r01_notes.md).  The partitioned path turns a chunk's updates into streams:
//   hist     count the chunk's (table, region) pairs            region = 2^rb buckets of one table
//   scan     exclusive prefix -> where each (table, region) run starts in the item array
//   scatter  write every update as a 32-bit bucket index into its run.  A CTA first counts its
//            tile in shared memory, then reserves room in each run with ONE global atomic per
//            (CTA, run), so its items land next to each other and L2 merges them into full sectors
//   apply    walk the item array front to back: at any moment all CTAs work inside the same
//            16 MB region, which therefore stays L2- and TLB-resident while it is hit; same
//            speculative / exact update logic, rollback and redo as the direct kernel.

#define KV_PART_MAX 4096        // (table, region) runs

struct KvPartInfo {
    int rb;                        // log2(buckets per region)
    uint32_t pbase[KV_TABLES_DEV + 1];   // first run of table t; pbase[n_tables] = number of runs
};

// CTA b owns positions [b*slice, (b+1)*slice) in BOTH hist and scatter.  rows[run * G + b] first
// holds how many items of `run` CTA b produces, after kv_part_rowscan_kernel the number produced
// by CTAs before b (exclusive prefix inside the run).
__global__ void __launch_bounds__(256) kv_part_hist_kernel(KvView v, KvPartInfo pi, const uint64_t *__restrict__ hashes,
                                                           const uint32_t *__restrict__ valid, uint64_t total, uint64_t slice,
                                                           uint32_t *__restrict__ rows)
{
    extern __shared__ uint32_t sm_cnt[];
    const int P = (int)pi.pbase[v.n_tables];
    for (int q = threadIdx.x; q < P; q += blockDim.x) sm_cnt[q] = 0;
    __syncthreads();
    const uint64_t lo = (uint64_t)blockIdx.x * slice, hi = lo + slice < total ? lo + slice : total;
    for (uint64_t g0 = lo; g0 < hi; g0 += blockDim.x) {
        const uint64_t g = g0 + threadIdx.x;
        const bool live = g < hi && (!valid || ((__ldg(valid + (g >> 5)) >> (g & 31)) & 1u));
        const uint64_t h = live ? __ldcs(hashes + g) : 0;
        if (live)
            for (int t = 0; t < v.n_tables; t++) {
                uint64_t bin;
                if (kv_bin(v, t, h, bin)) atomicAdd(&sm_cnt[pi.pbase[t] + (uint32_t)(bin >> pi.rb)], 1u);
            }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < P; q += blockDim.x) rows[(size_t)q * gridDim.x + blockIdx.x] = sm_cnt[q];
}

// one CTA per run: exclusive scan of its G per-CTA counts, run total to runsum[run]  (G <= 2048)
__global__ void __launch_bounds__(256) kv_part_rowscan_kernel(uint32_t *__restrict__ rows, int G, uint32_t *__restrict__ runsum)
{
    __shared__ uint32_t warp_tot[8];
    uint32_t *row = rows + (size_t)blockIdx.x * G;
    uint32_t val[8], mine = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int i = threadIdx.x * 8 + j;
        val[j] = i < G ? row[i] : 0u;
        mine += val[j];
    }
    uint32_t incl = mine;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < wid; w++) before += warp_tot[w];
    uint32_t run = before + incl - mine;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int i = threadIdx.x * 8 + j;
        if (i < G) row[i] = run;
        run += val[j];
    }
    if (threadIdx.x == 255) runsum[blockIdx.x] = before + incl;
}

// meta[0] = number of items, meta[1 + t] = index of the first item of table t (t <= n_tables)
__global__ void kv_part_scan_kernel(KvPartInfo pi, int n_tables, const uint32_t *__restrict__ runsum,
                                    uint32_t *__restrict__ runbase, uint32_t *__restrict__ meta)
{
    if (threadIdx.x || blockIdx.x) return;
    const int P = (int)pi.pbase[n_tables];
    uint32_t run = 0;
    int t = 0;
    for (int q = 0; q < P; q++) {
        while (t <= n_tables && (uint32_t)q == pi.pbase[t]) meta[1 + t++] = run;
        runbase[q] = run;
        run += runsum[q];
    }
    while (t <= n_tables) meta[1 + t++] = run;
    meta[0] = run;
}

// Each CTA re-walks its slice; a shared-memory cursor per run hands out absolute item slots, so
// the CTA's items of one run are written next to each other (L2 merges them into full sectors)
// and no global atomic is needed.
__global__ void __launch_bounds__(256) kv_part_scatter_kernel(KvView v, KvPartInfo pi, const uint64_t *__restrict__ hashes,
                                                              const uint32_t *__restrict__ valid, uint64_t total,
                                                              uint64_t slice, const uint32_t *__restrict__ rows,
                                                              const uint32_t *__restrict__ runbase,
                                                              uint32_t *__restrict__ items)
{
    extern __shared__ uint32_t sm_cur[];
    const int P = (int)pi.pbase[v.n_tables];
    for (int q = threadIdx.x; q < P; q += blockDim.x) sm_cur[q] = runbase[q] + rows[(size_t)q * gridDim.x + blockIdx.x];
    __syncthreads();
    const uint64_t lo = (uint64_t)blockIdx.x * slice, hi = lo + slice < total ? lo + slice : total;
    for (uint64_t g0 = lo; g0 < hi; g0 += blockDim.x) {
        const uint64_t g = g0 + threadIdx.x;
        const bool live = g < hi && (!valid || ((__ldg(valid + (g >> 5)) >> (g & 31)) & 1u));
        const uint64_t h = live ? __ldcs(hashes + g) : 0;
        if (live)
            for (int t = 0; t < v.n_tables; t++) {
                uint64_t bin;
                if (kv_bin(v, t, h, bin)) items[atomicAdd(&sm_cur[pi.pbase[t] + (uint32_t)(bin >> pi.rb)], 1u)] = (uint32_t)bin;
            }
    }
}

// Apply / undo / redo over the item array.  MODE 0: speculative update, records `added`;
// MODE 1: rollback of a dirty chunk; MODE 2: exact redo of a dirty chunk.
template <int BITS, int MODE>
__global__ void __launch_bounds__(256, KV_INC_MIN_CTAS) kv_part_apply_kernel(KvView v, const uint32_t *__restrict__ items,
                                                            const uint32_t *__restrict__ meta, uint32_t *__restrict__ added,
                                                            unsigned *dirty, unsigned long long *n_redone)
{
    if (MODE != 0 && *dirty == 0) return;
    if (MODE == 2 && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(n_redone, 1ULL);
    const unsigned maxv = BITS == 8 ? 255u : 15u;
    const uint64_t n = meta[0];
    // each CTA iteration covers 4 consecutive groups of 256 items; a thread owns one item in each
    // group, so four independent atomics are in flight per thread
    const uint64_t span = 4ull * blockDim.x;
    const uint64_t n_pad = (n + span - 1) / span * span;
    const uint64_t stride = (uint64_t)gridDim.x * span;
    for (uint64_t i0 = (uint64_t)blockIdx.x * span + threadIdx.x; i0 < n_pad; i0 += stride) {
        unsigned *w[4], sh[4], ob[4];
        uint64_t bin[4];
        int tt[4];
        bool live[4], hot[4], did[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t i = i0 + (uint64_t)j * blockDim.x;
            live[j] = i < n;
            did[j] = false;
            hot[j] = true;
            ob[j] = 0;
            if (live[j]) {
                int t = 0;
                while (t + 1 < v.n_tables && i >= meta[2 + t]) t++;
                tt[j] = t;
                bin[j] = __ldcs(items + i);
                kv_word_addr<BITS>(v, t, bin[j], w[j], sh[j]);
                if (MODE == 0) hot[j] = kv_maybe_hot(v, t, bin[j]);
            }
        }
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (live[j] && !hot[j]) {
                    ob[j] = (atomicAdd(w[j], 1u << sh[j]) >> sh[j]) & maxv;
                    did[j] = true;
                }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t i = i0 + (uint64_t)j * blockDim.x;
            if (live[j]) {
                if (MODE == 1) {
                    if ((added[i >> 5] >> (i & 31)) & 1u) atomicAdd(w[j], 0u - (1u << sh[j]));
                } else if (hot[j]) {
                    did[j] = kv_sat_inc_exact<BITS>(w[j], sh[j], ob[j]);
                    if (did[j] && MODE == 2) kv_state_publish<BITS>(v, tt[j], bin[j], ob[j]);   // MODE 0: the group is hot already
                } else if (ob[j] == maxv)
                    atomicOr(dirty, 1u);
                else
                    kv_state_publish<BITS>(v, tt[j], bin[j], ob[j]);
            }
            if (MODE == 0) {
                unsigned bal = __ballot_sync(0xffffffffu, did[j]);
                if ((threadIdx.x & 31) == 0 && i < n_pad) added[i >> 5] = bal;
            }
        }
    }
}

// ------------------------------------------------------------- K3c: tiled updates in shared memory
//
// For sketches that do not fit L2, a counter update in place is a random DRAM sector read-modify-write
// (~22 G/s on B200, row-activation bound, whatever the instruction).  K3c never does that: the hash
// kernel files each update under its (table, region) run -- region = 2^rb <= 65536 adjacent buckets --
// as a 16-bit offset, and here ONE CTA per run
//   streams its region from HBM into shared memory, widened to 16 bits per counter,
//   applies the run's offsets with shared-memory adds (two counters per 32-bit word; a lane cannot
//     overflow before the clamp because at most KV_TILE_BATCH offsets are applied between clamps),
//   clamps to the counter maximum and streams the region back.
// All HBM traffic is coalesced: the region (read + write) and 2 bytes per update.  Exact by
// construction -- the CTA owns the region, so there is no speculation, rollback or hot bitmap.
// Regions that received only a few offsets are cheaper to update in place (cnt * 64 B of sector
// traffic against 2 x region bytes): those go through the exact global path.  Bit tables (Bloom
// filters) keep their bytes and use shared-memory ORs.
// khmer add() on sketches that have no hot bitmap (spanning sketches): the exact update for every hash
__global__ void __launch_bounds__(256) kv_add_exact_kernel(const __grid_constant__ KvView v, const uint64_t *__restrict__ hashes,
                                                           const uint32_t *__restrict__ valid, uint64_t n)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        if (valid && !((valid[g >> 5] >> (g & 31)) & 1u)) continue;
        const uint64_t h = hashes[g];
        for (int t = 0; t < v.n_tables; t++) {
            uint64_t bin;
            if (kv_bin(v, t, h, bin)) kv_bucket_inc_exact(v, t, bin);
        }
    }
}

// the producer's overflow marks (tracked sketches), applied in place once the n_unique passes are done
__global__ void __launch_bounds__(256) kv_tile_overflow_kernel(const __grid_constant__ KvView v, const __grid_constant__ KvTileInfo ti,
                                                               const uint64_t *__restrict__ hashes, uint64_t n_pos)
{
    if (*ti.ovf_any == 0) return;
    const uint64_t n_words = (n_pos + 31) / 32;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (int t = 0; t < v.n_tables; t++)
        for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
            uint32_t m = ti.ovf[(size_t)t * ti.ovf_stride + w];
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                uint64_t bin;
                if (kv_bin(v, t, hashes[w * 32 + b], bin)) kv_bucket_inc_exact(v, t, bin);
            }
        }
}

#define KV_TILE_THREADS 512
#define KV_TILE_BATCH 60000u     // offsets applied between two clamps: 255 + 60000 < 65536

// Spanning sketches (the tables spread over the HBM of `n` ranks): the slabs of ALL ranks feed a region,
// and only the rank whose HBM holds the region applies them -- the all-to-all of the update stream happens
// inside this kernel, as peer loads over NVLink of exactly the slabs a CTA needs.
struct KvTileSources {
    int n;                                   // 1: only `ti` (ordinary sketch)
    int rank;
    const uint32_t *cursor[KV_MAX_RANKS];
    const uint16_t *slab[KV_MAX_RANKS];
    uint64_t piece[KV_TABLES_DEV];           // bytes of table t per rank
    uint32_t own_lo[KV_TABLES_DEV];          // first region of table t that lives in this rank's HBM
    uint32_t own_n[KV_TABLES_DEV];           // how many of them
    uint32_t own_total;
};

template <int BITS, bool SPAN>
__global__ void __launch_bounds__(KV_TILE_THREADS) kv_tile_apply_kernel(const __grid_constant__ KvView v, const __grid_constant__ KvTileInfo ti0,
                                                                        uint32_t direct_below, const __grid_constant__ KvTileSources src)
{
    extern __shared__ uint32_t sm_tile[];   // BITS 8/4: (1 << rb) / 2 words of two 16-bit counters; BITS 1: (1 << rb) / 32 words
    __shared__ uint32_t s_cnt[KV_MAX_RANKS];
    // ordinary sketch: one CTA per run.  Spanning sketch: a resident grid strides over the runs whose regions
    // live in THIS rank's HBM (src.own_*), so no CTA is spent on a region somebody else applies.
    const uint32_t n_work = SPAN ? src.own_total : ti0.run_base[KV_TABLES_DEV];
    const int n_src = SPAN ? src.n : 1;
    // work item -> (table, run)
    auto locate = [&](uint32_t work, int &t) -> uint32_t {
        t = 0;
        if (SPAN) {
            uint32_t first = 0;
            while (t + 1 < v.n_tables && work >= first + src.own_n[t]) { first += src.own_n[t]; t++; }
            return ti0.run_base[t] + src.own_lo[t] + (work - first);
        }
        while (t + 1 < v.n_tables && work >= ti0.run_base[t + 1]) t++;
        return work;
    };
    // how many offsets source q filed for a region: thread q fetches its (remote) cursor one region AHEAD, so
    // the NVLink round trip overlaps the work on the current region
    uint32_t ahead = 0;
    if ((int)threadIdx.x < n_src && blockIdx.x < n_work) {
        int t0;
        const uint32_t r0 = locate(blockIdx.x, t0);
        ahead = SPAN ? src.cursor[threadIdx.x][r0] : ti0.cursor[r0];
    }
  for (uint32_t work = blockIdx.x; work < n_work; work += gridDim.x) {
    if (work != blockIdx.x) __syncthreads();   // the previous region is stored: shared memory may be reused
    int t;
    const uint32_t run = locate(work, t);
    const uint64_t bucket0 = (uint64_t)(run - ti0.run_base[t]) << ti0.rb;
    const uint32_t nb = (uint32_t)((v.size[t] - bucket0) < (1ull << ti0.rb) ? (v.size[t] - bucket0) : (1ull << ti0.rb));
    if ((int)threadIdx.x < n_src) {
        s_cnt[threadIdx.x] = ahead < ti0.cap ? ahead : ti0.cap;
        if (work + gridDim.x < n_work) {
            int tn;
            const uint32_t rn = locate(work + gridDim.x, tn);
            ahead = SPAN ? src.cursor[threadIdx.x][rn] : ti0.cursor[rn];
        }
    }
    __syncthreads();
    uint32_t total = 0;
    for (int q = 0; q < n_src; q++) total += s_cnt[q];
    if (total == 0) continue;
    if (total < direct_below) {   // sparse region: in place
        for (int q = 0; q < n_src; q++) {
            KvTileInfo ti = ti0;
            if (SPAN) { ti.cursor = const_cast<uint32_t *>(src.cursor[q]); ti.slab = const_cast<uint16_t *>(src.slab[q]); }
            const uint32_t cnt = s_cnt[q];
            for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) kv_bucket_inc_exact(v, t, bucket0 + ti.slab[kv_slab_index(ti, run, i)]);
        }
        continue;
    }
    if (BITS == 1) {
        uint8_t *g = v.tab[t] + (bucket0 >> 3);
        const uint32_t nbytes = (nb + 7) >> 3, nwords = (nbytes + 3) >> 2;
        for (uint32_t w = threadIdx.x; w < nwords; w += blockDim.x) sm_tile[w] = 0;
        __syncthreads();
        for (int q = 0; q < n_src; q++) {
            KvTileInfo ti = ti0;
            if (SPAN) { ti.cursor = const_cast<uint32_t *>(src.cursor[q]); ti.slab = const_cast<uint16_t *>(src.slab[q]); }
            const uint32_t cnt = s_cnt[q];
            for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
                const uint32_t o = ti.slab[kv_slab_index(ti, run, i)];
                atomicOr(&sm_tile[o >> 5], 1u << (o & 31));
            }
        }
        __syncthreads();
        // OR the collected bits into the table (bytes of a region are not shared with other regions:
        // bucket0 is a multiple of 2^rb >= 8)
        for (uint32_t b = threadIdx.x; b < nbytes; b += blockDim.x) {
            const uint8_t add = (uint8_t)(sm_tile[b >> 2] >> (8 * (b & 3)));
            if (add) g[b] |= add;
        }
        continue;
    }
    const unsigned maxv = BITS == 8 ? 255u : 15u;
    // ---- load + widen: 16 counters per thread and iteration
    if (BITS == 8) {
        const uint8_t *g = v.tab[t] + bucket0;          // 2^rb-aligned inside a 256-byte aligned table
        const uint32_t nvec = nb >> 4;
        const uint4 *gv = (const uint4 *)g;
        uint4 *sv = (uint4 *)sm_tile;
        for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) {
            const uint4 x = __ldcs(gv + i);
            uint4 lo, hi;
            lo.x = __byte_perm(x.x, 0, 0x4140); lo.y = __byte_perm(x.x, 0, 0x4342);
            lo.z = __byte_perm(x.y, 0, 0x4140); lo.w = __byte_perm(x.y, 0, 0x4342);
            hi.x = __byte_perm(x.z, 0, 0x4140); hi.y = __byte_perm(x.z, 0, 0x4342);
            hi.z = __byte_perm(x.w, 0, 0x4140); hi.w = __byte_perm(x.w, 0, 0x4342);
            sv[2 * i] = lo;
            sv[2 * i + 1] = hi;
        }
        for (uint32_t b = (nvec << 4) + threadIdx.x; b < nb + (nb & 1); b += blockDim.x)   // ragged end of the table
            ((uint16_t *)sm_tile)[b] = b < nb ? g[b] : 0;
    } else {
        const uint8_t *g = v.tab[t] + (bucket0 >> 1);   // even bucket = high nibble
        const uint32_t nbytes = (nb + 1) >> 1, nvec = nbytes >> 4;
        const uint4 *gv = (const uint4 *)g;
        uint4 *sv = (uint4 *)sm_tile;
        for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) {   // 16 bytes = 32 counters -> 16 words
            const uint4 x = __ldcs(gv + i);
            const uint32_t q[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t hi = (q[j] >> 4) & 0x0f0f0f0fu, lo = q[j] & 0x0f0f0f0fu;
                uint4 w;
                w.x = __byte_perm(hi, lo, 0x0400) & 0x00ff00ffu;
                w.y = __byte_perm(hi, lo, 0x0501) & 0x00ff00ffu;
                w.z = __byte_perm(hi, lo, 0x0602) & 0x00ff00ffu;
                w.w = __byte_perm(hi, lo, 0x0703) & 0x00ff00ffu;
                sv[4 * i + j] = w;
            }
        }
        for (uint32_t y = (nvec << 4) + threadIdx.x; y < nbytes; y += blockDim.x) {   // ragged end of the table
            const unsigned byte = g[y];
            sm_tile[y] = (byte >> 4) | ((byte & 15u) << 16);
        }
    }
    __syncthreads();
    // ---- apply, at most KV_TILE_BATCH offsets between clamps
    const uint32_t nwords = (nb + 1) >> 1;
    uint32_t since = 0;   // offsets applied since the last clamp
    if (SPAN && total <= KV_TILE_BATCH) {
        // common case: everything fits between two clamps -- one loop over the offsets of ALL sources, so
        // local and peer loads are in flight together
        for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) {
            uint32_t i = j;
            int q = 0;
            while (i >= s_cnt[q]) { i -= s_cnt[q]; q++; }
            KvTileInfo ti = ti0;
            ti.slab = const_cast<uint16_t *>(src.slab[q]);
            const uint32_t o = __ldcs(ti.slab + kv_slab_index(ti, run, i));
            atomicAdd(&sm_tile[o >> 1], 1u << (16 * (o & 1)));
        }
    } else
    for (int q = 0; q < n_src; q++) {
        KvTileInfo ti = ti0;
        if (SPAN) { ti.cursor = const_cast<uint32_t *>(src.cursor[q]); ti.slab = const_cast<uint16_t *>(src.slab[q]); }
        const uint32_t cnt = s_cnt[q];
        for (uint32_t base = 0; base < cnt;) {
            if (since == KV_TILE_BATCH) {
                __syncthreads();
                for (uint32_t w = threadIdx.x; w < nwords; w += blockDim.x) sm_tile[w] = __vminu2(sm_tile[w], maxv | (maxv << 16));
                __syncthreads();
                since = 0;
            }
            const uint32_t room = KV_TILE_BATCH - since;
            const uint32_t end = cnt - base < room ? cnt : base + room;
            for (uint32_t i = base + threadIdx.x; i < end; i += blockDim.x) {
                const uint32_t o = __ldcs(ti.slab + kv_slab_index(ti, run, i));
                atomicAdd(&sm_tile[o >> 1], 1u << (16 * (o & 1)));
            }
            since += end - base;
            base = end;
        }
    }
    __syncthreads();
    // ---- clamp + narrow + store
    if (BITS == 8) {
        uint8_t *g = v.tab[t] + bucket0;
        const uint32_t nvec = nb >> 4;
        uint4 *gv = (uint4 *)g;
        const uint4 *sv = (const uint4 *)sm_tile;
        const uint32_t lim = 0x00ff00ffu;
        for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) {
            uint4 lo = sv[2 * i], hi = sv[2 * i + 1], x;
            x.x = __byte_perm(__vminu2(lo.x, lim), __vminu2(lo.y, lim), 0x6420);
            x.y = __byte_perm(__vminu2(lo.z, lim), __vminu2(lo.w, lim), 0x6420);
            x.z = __byte_perm(__vminu2(hi.x, lim), __vminu2(hi.y, lim), 0x6420);
            x.w = __byte_perm(__vminu2(hi.z, lim), __vminu2(hi.w, lim), 0x6420);
            gv[i] = x;
        }
        for (uint32_t b = (nvec << 4) + threadIdx.x; b < nb; b += blockDim.x) {
            const unsigned c = ((const uint16_t *)sm_tile)[b];
            g[b] = (uint8_t)(c > 255u ? 255u : c);
        }
    } else {
        uint8_t *g = v.tab[t] + (bucket0 >> 1);
        const uint32_t nbytes = (nb + 1) >> 1, nvec = nbytes >> 4;
        uint4 *gv = (uint4 *)g;
        const uint4 *sv = (const uint4 *)sm_tile;
        for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) {
            uint32_t q[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint4 w = sv[4 * i + j];
                const uint32_t a = __vminu2(w.x, 0x000f000fu), b = __vminu2(w.y, 0x000f000fu), c = __vminu2(w.z, 0x000f000fu),
                               d = __vminu2(w.w, 0x000f000fu);
                // byte k of the result = (even counter << 4) | odd counter of word k
                const uint32_t hi = __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
                const uint32_t lo = __byte_perm(__byte_perm(a, b, 0x0062), __byte_perm(c, d, 0x0062), 0x5410);
                q[j] = (hi << 4) | lo;
            }
            gv[i] = make_uint4(q[0], q[1], q[2], q[3]);
        }
        for (uint32_t y = (nvec << 4) + threadIdx.x; y < nbytes; y += blockDim.x) {
            const uint32_t w = __vminu2(sm_tile[y], 0x000f000fu);
            g[y] = (uint8_t)(((w & 15u) << 4) | (w >> 16));
        }
    }
  }
}

// ----------------------------------------------------------------------- K5
//
// khmer's n_unique_kmers counts the add() calls that found at least one of their T buckets
// empty, in single-threaded file order.  That number depends only on the
// stream of hashes and on which buckets were empty when the batch started -- not on the
// counter values -- so it is computed per chunk, BEFORE the chunk's increments:
//   occurrence g is new  <=>  for some table t its bucket was empty at chunk start (occ bit
//                             clear) and g is the smallest position of the chunk touching it.
// first[] is ONE u32 array as long as the largest table (or bucket range), reused for every table
// in turn.  An entry holds (epoch << 28 | position); every pass draws a fresh, SMALLER epoch, so
// atomicMin makes the entries of older passes lose against anything written now and nothing is
// ever swept or reset (the array is memset once every 15 passes).  Per chunk:
//   table 0, pass A   fused into kv_hash_kernel: first[bin0] = min(first[bin0], tag | g) for every
//                     k-mer whose table-0 bucket is empty;
//   kv_first_compact  per position: owner of its table-0 bucket == itself -> new; owner is an
//                     EARLIER position with the SAME hash -> a repeat: it can neither be new nor
//                     change any minimum (the owner touches the same buckets before it), so it is
//                     dropped; everything else -- owners, mere collisions, occupied or foreign
//                     buckets -- goes on a compact (hash, position) list.  At sequencing depth the
//                     list is a fraction of the stream (30x: ~1/5), and the passes over the other
//                     tables cost what the list costs, not what the chunk costs;
//   tables 1..T-1     kv_first_min_list (pass A over the list), then kv_first_own_list (pass B:
//                     the position recorded in first[] is marked in fresh[], counted once).
// Tables larger than first[] (KV_FIRST_RANGE_LOG2, default 2^31 entries = 8 GB) are walked in bucket
// ranges [bin_lo, bin_lo + bin_n); then table 0 is not fused and the list holds every valid position.
#define KV_POS_BITS 28
#define KV_POS_MASK ((1u << KV_POS_BITS) - 1u)

// all tables' occupancy bitmaps in one launch: blockIdx.y = table
template <int BITS>
__global__ void __launch_bounds__(256) kv_occ_rebuild_all_kernel(const __grid_constant__ KvView v)
{
    const int t = blockIdx.y;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_words = (v.size[t] + 31) / 32;
    const uint64_t nbytes = BITS == 8 ? v.size[t] : v.size[t] / 2 + 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
        uint32_t word = 0;
        if (BITS == 8) {
            const uint64_t b0 = i * 32;
            if (b0 + 32 <= nbytes) {
                const uint4 *p = (const uint4 *)(v.tab[t] + b0);
                uint4 a = __ldcs(p), c = __ldcs(p + 1);
                const uint32_t q[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    uint32_t nz = __vcmpne4(q[j], 0u);   // 0xff per non-zero byte
                    word |= ((nz & 1u) | ((nz >> 7) & 2u) | ((nz >> 14) & 4u) | ((nz >> 21) & 8u)) << (4 * j);
                }
            } else {
                for (int j = 0; j < 32 && b0 + j < v.size[t]; j++)
                    if (v.tab[t][b0 + j]) word |= 1u << j;
            }
        } else {
            // 32 buckets = 16 bytes; even bucket = high nibble
            const uint64_t y0 = i * 16;
            for (int j = 0; j < 16 && y0 + j < nbytes; j++) {
                unsigned byte = v.tab[t][y0 + j];
                if (byte >> 4) word |= 1u << (2 * j);
                if ((byte & 15u) && i * 32 + 2 * j + 1 < v.size[t]) word |= 1u << (2 * j + 1);
            }
        }
        v.occ[t][i] = word;
    }
}

// FUSED0: table 0's pass A ran inside the hash kernel with tag `tag0`.  Writes fresh[] (every word of
// the chunk), adds the new positions to *n_unique, and files the positions that still matter on the
// list.  The list is segmented: the CTA working on positions [seg << KV_SEG_LOG2, +2^KV_SEG_LOG2) appends
// to the list slots of the same index range through a shared-memory cursor and leaves the count in
// seg_cnt[seg] -- no global counter that a million warps would fight over.
#define KV_SEG_LOG2 12

// KV_COMPACT_U = positions per thread and iteration, their random loads in flight together, first[] fetched before
// the occupancy bit is known.  4 for a sample counted into its own sketch (most buckets empty at chunk start; C2 at
// N = 1: 401 -> 351 us); 1 when another rank's occupancy is the occupied set or a merge shares the GPU (the host picks).
template <bool FUSED0, int KV_COMPACT_U>
__global__ void __launch_bounds__(256) kv_first_compact_kernel(const __grid_constant__ KvView v, const uint32_t *__restrict__ first,
                                                               const uint64_t *__restrict__ hashes,
                                                               const uint32_t *__restrict__ valid, uint64_t total,
                                                               uint32_t *__restrict__ fresh, uint64_t *__restrict__ list_h,
                                                               uint32_t *__restrict__ list_p, uint32_t *__restrict__ seg_cnt,
                                                               unsigned long long *n_unique)
{
    __shared__ unsigned s_cur;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t n_segs = (total + (1u << KV_SEG_LOG2) - 1) >> KV_SEG_LOG2;
    unsigned n_new = 0;
    for (uint64_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
        if (threadIdx.x == 0) s_cur = 0;
        __syncthreads();
        const uint64_t base = seg << KV_SEG_LOG2;
        for (unsigned off0 = threadIdx.x; off0 < (1u << KV_SEG_LOG2); off0 += 256 * KV_COMPACT_U) {
            // the chain per position is hash -> (occupancy bit, first[bin]) -> hash of the bucket's owner: each stage is
            // issued for all KV_COMPACT_U positions before the next one looks at the results
            uint64_t g[KV_COMPACT_U], h[KV_COMPACT_U], bin[KV_COMPACT_U];
            bool keep[KV_COMPACT_U], is_new[KV_COMPACT_U], mine[KV_COMPACT_U];
            uint32_t owner[KV_COMPACT_U];
#pragma unroll
            for (int u = 0; u < KV_COMPACT_U; u++) {
                g[u] = base + off0 + 256u * u;
                keep[u] = g[u] < total && (!valid || ((__ldg(valid + (g[u] >> 5)) >> (g[u] & 31)) & 1u));
                is_new[u] = false;
                mine[u] = false;
                h[u] = keep[u] ? __ldcs(hashes + g[u]) : 0;
            }
            if (FUSED0) {
#pragma unroll
                for (int u = 0; u < KV_COMPACT_U; u++) {
                    mine[u] = keep[u] && kv_bin(v, 0, h[u], bin[u]);
                    owner[u] = mine[u] ? __ldcg(first + bin[u]) & KV_POS_MASK : 0;   // (speculative: only empty buckets took part in pass A)
                }
#pragma unroll
                for (int u = 0; u < KV_COMPACT_U; u++) mine[u] = mine[u] && kv_bucket_empty(v, 0, bin[u]);
                uint64_t oh[KV_COMPACT_U];
#pragma unroll
                for (int u = 0; u < KV_COMPACT_U; u++) {
                    is_new[u] = mine[u] && owner[u] == (uint32_t)g[u];
                    oh[u] = mine[u] && !is_new[u] ? __ldg(hashes + owner[u]) : ~h[u];
                }
#pragma unroll
                for (int u = 0; u < KV_COMPACT_U; u++)
                    if (oh[u] == h[u]) keep[u] = false;   // repeat of an earlier occurrence
            }
#pragma unroll
            for (int u = 0; u < KV_COMPACT_U; u++) {
                const unsigned new_bal = __ballot_sync(0xffffffffu, is_new[u]);
                const unsigned keep_bal = __ballot_sync(0xffffffffu, keep[u]);
                if (lane == 0 && g[u] < ((total + 31) & ~(uint64_t)31)) {
                    fresh[g[u] >> 5] = new_bal;
                    n_new += __popc(new_bal);
                }
                if (keep_bal) {
                    unsigned wbase = 0;
                    if (lane == 0) wbase = atomicAdd(&s_cur, (unsigned)__popc(keep_bal));
                    wbase = __shfl_sync(0xffffffffu, wbase, 0);
                    if (keep[u]) {
                        const uint64_t slot = base + wbase + __popc(keep_bal & ((1u << lane) - 1u));
                        list_h[slot] = h[u];
                        list_p[slot] = (uint32_t)g[u];
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) seg_cnt[seg] = s_cur;
    }
    if (lane == 0 && n_new) atomicAdd(n_unique, (unsigned long long)n_new);
}

// pass A of table 0 over stored hashes (kv_unique_last_batch; otherwise this runs inside kv_hash_kernel)
__global__ void __launch_bounds__(256) kv_first_min0_kernel(const __grid_constant__ KvView v, uint32_t *__restrict__ first, uint32_t tag,
                                                            const uint64_t *__restrict__ hashes, const uint32_t *__restrict__ valid,
                                                            uint64_t n)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        if (!((__ldg(valid + (g >> 5)) >> (g & 31)) & 1u)) continue;
        uint64_t bin;
        if (kv_bin(v, 0, __ldg(hashes + g), bin) && kv_bucket_empty(v, 0, bin)) atomicMin(first + bin, tag | (uint32_t)g);
    }
}

// pass A of table t over the list (KV_LIST_U entries per thread and iteration, their occupancy loads in flight together;
// 4 or 1 like KV_COMPACT_U above)
template <int KV_LIST_U>
__global__ void __launch_bounds__(256) kv_first_min_list_kernel(const __grid_constant__ KvView v, int t, uint32_t *__restrict__ first,
                                                                uint32_t tag, const uint64_t *__restrict__ list_h,
                                                                const uint32_t *__restrict__ list_p,
                                                                const uint32_t *__restrict__ seg_cnt, uint64_t n_segs,
                                                                uint64_t bin_lo, uint64_t bin_n)
{
    for (uint64_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
        const uint32_t cnt = seg_cnt[seg];
        const uint64_t base = seg << KV_SEG_LOG2;
        for (uint32_t i0 = threadIdx.x; i0 < cnt; i0 += 256 * KV_LIST_U) {
            uint64_t bin[KV_LIST_U];
            uint32_t pos[KV_LIST_U];
            bool hit[KV_LIST_U];
#pragma unroll
            for (int u = 0; u < KV_LIST_U; u++) {
                const uint32_t i = i0 + 256u * u;
                hit[u] = i < cnt;
                const uint64_t h = hit[u] ? __ldg(list_h + base + i) : 0;
                pos[u] = hit[u] ? __ldg(list_p + base + i) : 0;
                hit[u] = hit[u] && kv_bin(v, t, h, bin[u]) && bin[u] - bin_lo < bin_n;
            }
#pragma unroll
            for (int u = 0; u < KV_LIST_U; u++) hit[u] = hit[u] && kv_bucket_empty(v, t, bin[u]);
#pragma unroll
            for (int u = 0; u < KV_LIST_U; u++)
                if (hit[u]) atomicMin(first + (bin[u] - bin_lo), tag | pos[u]);
        }
    }
}

// pass B of table t over the list: whoever is recorded in first[] is new (counted once over all tables)
template <int KV_LIST_U>
__global__ void __launch_bounds__(256) kv_first_own_list_kernel(const __grid_constant__ KvView v, int t, const uint32_t *__restrict__ first,
                                                                uint32_t tag, const uint64_t *__restrict__ list_h,
                                                                const uint32_t *__restrict__ list_p,
                                                                const uint32_t *__restrict__ seg_cnt, uint64_t n_segs,
                                                                uint64_t bin_lo, uint64_t bin_n, uint32_t *__restrict__ fresh,
                                                                unsigned long long *n_unique)
{
    unsigned n_new = 0;
    for (uint64_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
        const uint32_t cnt = seg_cnt[seg];
        const uint64_t base = seg << KV_SEG_LOG2;
        for (uint32_t i0 = threadIdx.x; i0 < cnt; i0 += 256 * KV_LIST_U) {
            uint64_t bin[KV_LIST_U];
            uint32_t pos[KV_LIST_U], rec[KV_LIST_U];
            bool hit[KV_LIST_U];
#pragma unroll
            for (int u = 0; u < KV_LIST_U; u++) {
                const uint32_t i = i0 + 256u * u;
                hit[u] = i < cnt;
                const uint64_t h = hit[u] ? __ldg(list_h + base + i) : 0;
                pos[u] = hit[u] ? __ldg(list_p + base + i) : 0;
                hit[u] = hit[u] && kv_bin(v, t, h, bin[u]) && bin[u] - bin_lo < bin_n;
                rec[u] = hit[u] ? __ldcg(first + (bin[u] - bin_lo)) : 0;   // (speculative: only empty buckets took part in pass A)
            }
#pragma unroll
            for (int u = 0; u < KV_LIST_U; u++) hit[u] = hit[u] && kv_bucket_empty(v, t, bin[u]);
#pragma unroll
            for (int u = 0; u < KV_LIST_U; u++)
                if (hit[u] && rec[u] == (tag | pos[u])) {
                    const uint32_t bit = 1u << (pos[u] & 31);
                    n_new += !(atomicOr(fresh + (pos[u] >> 5), bit) & bit);
                }
        }
    }
    n_new = __reduce_add_sync(0xffffffffu, n_new);
    if ((threadIdx.x & 31) == 0 && n_new) atomicAdd(n_unique, (unsigned long long)n_new);
}

// mark every bucket of the listed k-mers as occupied (kv_unique_batch between the chunks of one rank)
__global__ void __launch_bounds__(256) kv_occ_mark_list_kernel(const __grid_constant__ KvView v, const uint64_t *__restrict__ list_h,
                                                               const uint32_t *__restrict__ seg_cnt, uint64_t n_segs)
{
    for (uint64_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
        const uint32_t cnt = seg_cnt[seg];
        const uint64_t base = seg << KV_SEG_LOG2;
        for (uint32_t i = threadIdx.x; i < cnt; i += 256) {
            const uint64_t h = __ldg(list_h + base + i);
            for (int t = 0; t < v.n_tables; t++) {
                uint64_t bin;
                if (kv_bin(v, t, h, bin)) atomicOr(v.occ[t] + (bin >> 5), 1u << (bin & 31));
            }
        }
    }
}

// khmer abundance_distribution (kevlar/dist.py:55): hist[counts.get(h)] += 1 for every position the
// first-touch passes marked fresh in the TRACKING sketch.  Per-CTA histogram in shared memory.
__global__ void __launch_bounds__(256) kv_abund_dist_kernel(KvView counts, const uint64_t *__restrict__ hashes,
                                                            const uint32_t *__restrict__ fresh, uint64_t n,
                                                            unsigned long long *__restrict__ hist)
{
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride)
        if ((__ldg(fresh + (g >> 5)) >> (g & 31)) & 1u) atomicAdd(&sh[kv_get(counts, hashes[g]) & 255u], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}

// khmer _occupied_bins: non-zero buckets of table 0
__global__ void kv_occupied_kernel(KvView v, unsigned long long *out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n = v.size[0];
    unsigned mine = 0;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride)
        mine += kv_bucket_get(v, 0, b) != 0;
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, (unsigned long long)mine);
}

// ----------------------------------------------------------------------- K4

struct KvNovelParams {
    const uint8_t *bases;
    const uint64_t *offsets;
    const uint32_t *tile_first;
    uint64_t total, n_tiles;
    int k;
    int n_case, n_ctrl;
    int case_min, ctrl_max, screen;
    int banded;                 // kevlar/novel.py:144-147 bit test (App. B.1)
    uint64_t band_mask;         // num_bands - 1
    long long band_minus_1;
    kv_hit *hits;
    unsigned long long max_hits;
    unsigned long long *n_hits;
    uint32_t *read_flags;       // u32 per read (no byte atomics on the device)
    uint32_t *discard_pos;      // per read, atomicMin; NULL when screen <= 0
    KvView sk[KV_MAX_SAMPLES];  // cases then controls
    const uint8_t *pre[KV_MAX_SAMPLES];   // non-NULL: abundance of sample s at every base position, already
                                          // computed (sharded sketches: all-reduced partial minima)
};

// FAST: no abundance screen and no precomputed counts -> order-free evaluation (see below)
template <int HASHER, int KW, bool FAST>
__global__ void __launch_bounds__(KV_THREADS) kv_novel_kernel(const __grid_constant__ KvNovelParams p)
{
    __shared__ KvTileSmem sm;
    __shared__ KvTileList ls;
    const uint64_t tile = blockIdx.x;
    const uint64_t tile_start = tile * KV_TILE;
    kv_tile_load<HASHER == KV_HASH_TWOBIT, true>(sm, p.bases, tile_start, p.total);
    kv_tile_bounds(sm, p.offsets, p.tile_first, tile);
    __syncthreads();

    // any byte outside ACGT poisons the whole read (kevlar/novel.py:136-139)
    for (int it = 0; it < KV_TILE / KV_THREADS; it++) {
        const int l = it * KV_THREADS + threadIdx.x;
        const int L = l + KV_FRONT;
        if (tile_start + l < p.total && ((sm.bad[L >> 5] >> (L & 31)) & 1u)) {
            uint64_t read, rs, re;
            kv_find_read(sm, p.offsets, p.tile_first, tile, tile_start + l, read, rs, re);
            atomicOr(p.read_flags + read, KV_READ_SKIPPED);
        }
    }
    // k-mers of the tile (none of them touches a byte outside ACGT)
    const unsigned n_list = kv_tile_kmer_list(sm, ls, p.offsets, p.tile_first, tile, p.total, p.k, true);

#pragma unroll 1
    for (unsigned i = threadIdx.x; i < n_list; i += KV_THREADS) {
        const int l = ls.pos[i];
        const uint64_t g = tile_start + l;
        const uint64_t h = kv_tile_hash<HASHER, KW>(sm, l, p.k);
        if (p.banded && (long long)(h & p.band_mask) != p.band_minus_1) continue;

        uint8_t ab[KV_MAX_SAMPLES];
        bool interesting = true;
        if (FAST) {
            // kmer_is_interesting (kevlar/novel.py:21-53) without the abundance screen is a pure
            // predicate -- every case abundance >= case_min and every control abundance <= ctrl_max --
            // so the lookups may run in ANY order, and a Count-Min abundance is a minimum over tables:
            //   control passes as soon as ONE table is <= ctrl_max, fails only if ALL tables exceed it;
            //   case    fails as soon as ONE table is <  case_min.
            // Controls first: at sequencing depth most k-mers are inherited and die on the first control
            // (4 loads); error k-mers pass each control on its first table and die on the first case
            // table (3 loads) -- against 8 and 1-2 loads in the reference's case-first order.  Random
            // 32-byte sector loads are what bounds this kernel, so fewer loads is the whole game.
            for (int s = 0; s < p.n_ctrl && interesting; s++) {
                const KvView &v = p.sk[p.n_case + s];
                uint64_t bin;
                kv_bin(v, 0, h, bin);
                unsigned m = kv_bucket_get(v, 0, bin);
                if ((int)m > p.ctrl_max) {
#pragma unroll 3
                    for (int t = 1; t < v.n_tables; t++) {
                        kv_bin(v, t, h, bin);
                        unsigned c = kv_bucket_get(v, t, bin);
                        m = c < m ? c : m;
                    }
                    if ((int)m > p.ctrl_max) interesting = false;
                }
            }
            for (int s = 0; s < p.n_case && interesting; s++) {
                const KvView &v = p.sk[s];
                for (int t = 0; t < v.n_tables; t++) {
                    uint64_t bin;
                    kv_bin(v, t, h, bin);
                    if ((int)kv_bucket_get(v, t, bin) < p.case_min) { interesting = false; break; }
                }
            }
            if (!interesting) continue;
            // a hit (rare): now the full abundances for the annotation
            for (int s = 0; s < p.n_case + p.n_ctrl; s++) ab[s] = (uint8_t)kv_get(p.sk[s], h);
        } else {
        // kmer_is_interesting (kevlar/novel.py:21-53), same evaluation order and early exits
        for (int s = 0; s < p.n_case; s++) {
            int a = p.pre[s] ? (int)p.pre[s][g] : (int)kv_get(p.sk[s], h);
            if (a < p.case_min) {
                interesting = false;
                if (p.screen > 0 && a < p.screen) {
                    uint64_t read, rs, re;
                    kv_find_read(sm, p.offsets, p.tile_first, tile, g, read, rs, re);
                    atomicMin(p.discard_pos + read, (uint32_t)(g - rs));
                }
                break;
            }
            ab[s] = (uint8_t)a;
        }
        if (!interesting) continue;
        for (int s = 0; s < p.n_ctrl; s++) {
            int a = p.pre[p.n_case + s] ? (int)p.pre[p.n_case + s][g] : (int)kv_get(p.sk[p.n_case + s], h);
            if (a > p.ctrl_max) { interesting = false; break; }
            ab[p.n_case + s] = (uint8_t)a;
        }
        if (!interesting) continue;
        }
        unsigned long long slot = atomicAdd(p.n_hits, 1ULL);
        if (slot < p.max_hits) {
            uint64_t read, rs, re;
            kv_find_read(sm, p.offsets, p.tile_first, tile, g, read, rs, re);
            kv_hit hit;
            hit.read = (uint32_t)read;
            hit.offset = (uint32_t)(g - rs);
#pragma unroll
            for (int s = 0; s < KV_MAX_SAMPLES; s++) hit.abund[s] = s < p.n_case + p.n_ctrl ? ab[s] : 0;
            p.hits[slot] = hit;
        }
    }
}

// ----------------------------------------------------------------------- K6

// Widened copies for an NCCL sum.  NCCL has no 16-bit integer type, so 8-bit counters travel as
// IEEE half: integers up to 2048 are exact in fp16 and 8 ranks x 255 = 2040, so the sum is exact
// for up to 8 ranks (larger worlds use the all-gather / peer-to-peer merge).
// 8-bit: u8 -> half;  4-bit: one u8 per nibble (sums <= 15 x 17 fit);  1-bit: raw bytes.
__global__ void kv_widen_kernel(const uint8_t *__restrict__ flat, uint64_t nbytes, int bits, void *out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += stride) {
        uint8_t b = flat[i];
        if (bits == 8) ((__half *)out)[i] = __ushort2half_rn(b);
        else if (bits == 4) { ((uint8_t *)out)[2 * i] = b >> 4; ((uint8_t *)out)[2 * i + 1] = b & 15; }
        else ((uint8_t *)out)[i] = b;
    }
}

__global__ void kv_narrow_kernel(uint8_t *__restrict__ flat, uint64_t nbytes, int bits, const void *in)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += stride) {
        if (bits == 8) {
            unsigned s = __half2uint_rn(((const __half *)in)[i]);
            flat[i] = (uint8_t)(s > 255u ? 255u : s);
        } else if (bits == 4) {
            unsigned hi = ((const uint8_t *)in)[2 * i], lo = ((const uint8_t *)in)[2 * i + 1];
            hi = hi > 15u ? 15u : hi; lo = lo > 15u ? 15u : lo;
            flat[i] = (uint8_t)((hi << 4) | lo);
        } else flat[i] = ((const uint8_t *)in)[i];
    }
}

struct KvPeers {
    uint4 *peer[KV_MAX_RANKS - 1];
    int n;
};

// Saturating merge over NVLink: each thread pulls 16 bytes from every peer's table (peer-mapped
// device pointers) and folds them into the local table.  8-bit counters use the SIMD
// per-byte unsigned saturating add; nibbles are split into two byte lanes and clamped at 15.
__device__ __forceinline__ uint32_t kv_sat_merge_word(uint32_t a, uint32_t b, int bits)
{
    if (bits == 8) return __vaddus4(a, b);
    if (bits == 1) return a | b;
    uint32_t lo = __vminu4(__vaddus4(a & 0x0f0f0f0fu, b & 0x0f0f0f0fu), 0x0f0f0f0fu);
    uint32_t hi = __vminu4(__vaddus4((a >> 4) & 0x0f0f0f0fu, (b >> 4) & 0x0f0f0f0fu), 0x0f0f0f0fu);
    return lo | (hi << 4);
}

// Every thread keeps 4 to 8 loads over NVLink in flight: U vectors x NP peers, all issued -- unconditionally, so
// that the compiler cannot serialise them behind predicates -- before the first one is used.  EXACT: the sketch has
// exactly NP peers (worlds up to 8: one instantiation per peer count); otherwise peers are taken in groups of NP = 8
// and a short last group re-reads its last peer and discards the copy.  The depth comes from each thread because on
// the merge lane (kv_merge_fork) the kernel gets one CTA per SM.  Measured at 8 ranks, 3 x 4 GB sketches: 10.5 GB
// read from and 10.5 GB stored to the peers per rank in 34 ms -- every link direction carries the read responses of
// one side plus the stores of the other, 614 GB/s per direction, about what NVLink 5 delivers in practice; the wide
// single-load version of round 1 reached the same rate with 16x the CTAs (profiles/r02_notes.md, section 4).
// Vectors past the end are clamped for the loads and skipped by the stores.
template <bool PUSH, int NP, int U, bool EXACT>
__global__ void __launch_bounds__(256) kv_merge_peers_kernel(uint4 *__restrict__ local, uint64_t n_vec, int bits,
                                                             const __grid_constant__ KvPeers peers)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int n_peers = EXACT ? NP : peers.n;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += stride * U) {
        uint64_t idx[U];
        uint4 acc[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t i = i0 + (uint64_t)u * stride;
            idx[u] = i < n_vec ? i : n_vec - 1;
            acc[u] = local[idx[u]];
        }
        for (int p0 = 0; p0 < n_peers; p0 += NP) {
            uint4 o[U][NP];
#pragma unroll
            for (int j = 0; j < NP; j++) {
                const uint4 *src = peers.peer[EXACT ? j : (p0 + j < n_peers ? p0 + j : n_peers - 1)];
#pragma unroll
                for (int u = 0; u < U; u++) o[u][j] = src[idx[u]];
            }
#pragma unroll
            for (int j = 0; j < NP; j++) {
                if (!EXACT && p0 + j >= n_peers) break;
#pragma unroll
                for (int u = 0; u < U; u++) {
                    acc[u].x = kv_sat_merge_word(acc[u].x, o[u][j].x, bits);
                    acc[u].y = kv_sat_merge_word(acc[u].y, o[u][j].y, bits);
                    acc[u].z = kv_sat_merge_word(acc[u].z, o[u][j].z, bits);
                    acc[u].w = kv_sat_merge_word(acc[u].w, o[u][j].w, bits);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (i0 + (uint64_t)u * stride >= n_vec) break;
            local[idx[u]] = acc[u];
            // all-reduce in one pass: this rank owns the slice, so it also stores the finished
            // vector into every peer's table (nobody else reads or writes this slice anywhere)
            if (PUSH) {
                if (EXACT) {
#pragma unroll
                    for (int p = 0; p < NP; p++) peers.peer[p][idx[u]] = acc[u];
                } else
                    for (int p = 0; p < n_peers; p++) peers.peer[p][idx[u]] = acc[u];
            }
        }
    }
}


// ------------------------------------------------------------------ device-side rank barrier
//
// Every rank owns an array of u32 flags in its HBM that its peers map over CUDA IPC; flag[p] is
// written only by rank p.  Barrier number `epoch` (1, 2, ...; the same count on every rank because
// barriers are collective): thread p publishes `epoch` into MY slot of peer p's array (release,
// system scope -- everything earlier kernels of this stream wrote is visible to the peer before
// the flag is) and then spins on peer p's slot of my array.  A peer can be at most one barrier
// ahead, so ">= epoch" is the condition.  No host round trip; the next kernel in the stream
// starts when all peers have arrived.
struct KvPeerFlags {
    uint32_t *peer[KV_MAX_RANKS];   // peer p's flag array (mapped), NULL for p == rank
    uint32_t *mine;
    int rank, world;
};

__global__ void kv_peer_barrier_kernel(KvPeerFlags f, uint32_t epoch, unsigned long long timeout_ns, unsigned *timed_out)
{
    const int p = threadIdx.x;
    if (p >= f.world || p == f.rank) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.peer[p] + f.rank), "r"(epoch) : "memory");
    unsigned long long t0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t seen;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f.mine + p) : "memory");
        if ((int32_t)(seen - epoch) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) { atomicExch(timed_out, 1u); break; }
        __nanosleep(200);
    }
}

// ------------------------------------------------------------------ synthetic reads (measurement fixture)
//
// wgsim-style reads drawn on the device (kevlar/tests/data/minitrio/README recipe, 100 bp,
// iid substitution errors): read r of a sample is a pure function of (seed, r), so any rank can
// produce any slice of a sample's read set and the union over ranks never depends on the number
// of ranks.  Counter-based generator: splitmix64 finaliser over (seed, index).
struct KvSynthParams {
    const uint8_t *hap[8];
    uint64_t hap_len[8];
    int n_haps;
    uint64_t n_reads, first_read;
    uint32_t read_len;
    uint32_t err_q32;     // substitution probability * 2^32
    uint64_t seed;
    uint8_t *out;         // n_reads * read_len bases
    uint64_t *offsets;    // n_reads + 1 (may be NULL)
};

__device__ __forceinline__ uint64_t kv_mix64(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) kv_synth_reads_kernel(KvSynthParams p)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_bases = p.n_reads * p.read_len;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_bases; i += stride) {
        const uint64_t r = i / p.read_len, j = i - r * p.read_len, gr = p.first_read + r;
        const uint64_t a = kv_mix64(p.seed ^ kv_mix64(gr));
        const int h = (int)((a & 0xffff) % (uint64_t)p.n_haps);
        const uint64_t start = (a >> 17) % (p.hap_len[h] - p.read_len);
        const bool minus = (a >> 16) & 1;
        uint8_t b = p.hap[h][start + (minus ? p.read_len - 1 - j : j)];
        // codes in ACGT order: A=0 C=1 G=2 T=3; complement = 3 - code
        unsigned code = b == 'A' ? 0u : b == 'C' ? 1u : b == 'G' ? 2u : 3u;
        if (minus) code = 3u - code;
        const uint64_t e = kv_mix64((p.seed + 0x632be59bd9b4e019ULL) ^ kv_mix64(gr * p.read_len + j));
        if ((uint32_t)e < p.err_q32) code = (code + 1u + (unsigned)((e >> 32) % 3u)) & 3u;
        p.out[i] = "ACGT"[code];
        if (p.offsets && j == 0) p.offsets[r] = i;
    }
    if (p.offsets && blockIdx.x == 0 && threadIdx.x == 0) p.offsets[p.n_reads] = n_bases;
}

// reads the novel scan has something to say about (skipped, or discarded by the abundance screen) as a
// compact list: at sequencing scale that is a handful out of millions, and the host only needs those
struct KvReadNote {
    uint32_t read, flags, discard;
};

__global__ void kv_read_notes_kernel(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ discard, uint64_t n_reads,
                                     KvReadNote *__restrict__ notes, unsigned long long cap, unsigned long long *n_notes)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        const uint32_t f = flags[r], d = discard ? discard[r] : 0xffffffffu;
        if (f || d != 0xffffffffu) {
            const unsigned long long slot = atomicAdd(n_notes, 1ULL);
            if (slot < cap) notes[slot] = KvReadNote{(uint32_t)r, f, d};
        }
    }
}

// flag reads shorter than k as skipped (kevlar/novel.py:134): one thread per read
__global__ void kv_short_reads_kernel(const uint64_t *__restrict__ offsets, uint64_t n_reads, int k, uint32_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride)
        if (offsets[r + 1] - offsets[r] < (uint64_t)k) atomicOr(flags + r, KV_READ_SKIPPED);
}
