/*
 * kvsketch.h -- C ABI of libkvsketch.so: the B200 (sm_100a) implementation of the khmer
 * sketch operations that kevlar's `count` -> `novel` -> `filter` path calls.
 *
 * The drop-in boundary of the reference is the `khmer` Python namespace (SURVEY.md 8b;
 * /root/reference has no FFI of its own -- khmer is a Cython/C++ dependency).  Each entry
 * point below names the khmer call, and the kevlar call site (file:line under
 * /root/reference), that it replaces.  The host-side mirror that binds these with ctypes is
 * kevlar_b200/khmer/; INTEGRATION.md shows the binding a kevlar maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns 0 on success or a negative
 *    KV_E* code, with a human-readable message available from kv_last_error() (thread local).
 *  - there is NO CPU fallback: every call that computes fails with KV_ENODEVICE when no
 *    CUDA device is usable.
 *  - a "batch" is `bases` = the concatenated sequence bytes of n_reads reads (ASCII, as read
 *    from FASTA/FASTQ) and `offsets` = n_reads+1 uint64 byte offsets into it.  `where` says
 *    whether both pointers are host (KV_MEM_HOST; pinned memory recommended) or device
 *    (KV_MEM_DEVICE) pointers.
 *  - work is enqueued on the library's per-device stream (kv_stream); calls that return
 *    values to the host synchronise, the others are asynchronous: host batch buffers must
 *    stay valid until kv_sync().
 *  - handles are caller-owned.  A sketch may be consumed into from several host threads
 *    (increments commute; calls are serialised per device); kv_sketch_save/stats need
 *    quiescence.
 */
#ifndef KVSKETCH_H
#define KVSKETCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KV_ABI_VERSION 1

/* hash function: khmer *table types use MurmurHash3, *graph types the 2-bit encoding */
#define KV_HASH_MURMUR 0 /* Counttable / SmallCounttable / Nodetable   (kevlar/sketch.py:107-116) */
#define KV_HASH_TWOBIT 1 /* Countgraph / SmallCountgraph / Nodegraph   (kevlar/sketch.py:102-106) */

#define KV_MEM_HOST 0
#define KV_MEM_DEVICE 1

#define KV_MAX_TABLES 8   /* tables per sketch (kevlar always builds 4: kevlar/count.py:29) */
#define KV_MAX_SAMPLES 16 /* case + control sketches in one novel scan */
#define KV_MAX_KSIZE_MURMUR 64
#define KV_MAX_KSIZE_TWOBIT 32

#define KV_OK 0
#define KV_EINVAL (-1)    /* bad argument                      -> ValueError */
#define KV_EIO (-2)       /* unreadable / malformed sketch file -> OSError    */
#define KV_ENOMEM (-3)    /* host or device allocation failed  -> MemoryError */
#define KV_ECUDA (-4)     /* CUDA runtime error                -> RuntimeError */
#define KV_ENODEVICE (-5) /* no usable CUDA device             -> RuntimeError */
#define KV_EOVERFLOW (-6) /* output buffer too small           -> RuntimeError */
#define KV_ESTATE (-7)    /* a shortcut's precondition does not hold; the call changed nothing -> caller takes the general path */

typedef struct kv_sketch kv_sketch;

/* one interesting k-mer found by kv_novel_batch (kevlar/novel.py:157-162: irecord.annotate) */
typedef struct kv_hit {
    uint32_t read;                  /* read index within the batch */
    uint32_t offset;                /* k-mer offset within the read */
    uint8_t abund[KV_MAX_SAMPLES];  /* case abundances, then control abundances */
} kv_hit;

/* per-read flags written by kv_novel_batch */
#define KV_READ_SKIPPED 1u   /* shorter than k, or contains a byte outside ACGT (kevlar/novel.py:134-139) */
#define KV_READ_DISCARDED 2u /* abundance screen tripped (kevlar/novel.py:152-154) */

const char *kv_last_error(void);
int kv_abi_version(void);
int kv_device_count(int *n);

/* khmer get_n_primes_near_x(n, x) -- the table sizes every khmer constructor derives from
 * `starting_size` (kevlar/sketch.py:118, kevlar/filter.py:29).  Host arithmetic. */
int kv_primes_below(uint64_t x, int n, uint64_t *out);

/* khmer.{Count,SmallCount,Node}{table,graph}(ksize, starting_size, n_tables) with the table
 * sizes already chosen (kevlar/sketch.py:99-119).  bits = 8 | 4 | 1.  Zero-filled tables are
 * allocated in the HBM of `device`. */
int kv_sketch_create(int hasher, int bits, int ksize, int n_tables, const uint64_t *sizes, int device,
                     kv_sketch **out);
int kv_sketch_destroy(kv_sketch *s);
/* Zero every counter and the n_unique bookkeeping (a fresh sketch without a reallocation; asynchronous). */
int kv_sketch_clear(kv_sketch *s);

/* khmer <Type>.load(filename) / .save(filename): OXLI v4 files (kevlar/sketch.py:14-27,77-92;
 * kevlar/count.py:95; kevlar/novel.py:92).  The file does not record table-vs-graph, so the
 * caller passes the hasher; expect_bits (8|4|1, or 0 for any) is checked against the file. */
int kv_sketch_load(const char *path, int hasher, int expect_bits, int device, kv_sketch **out);
int kv_sketch_save(kv_sketch *s, const char *path);

/* .ksize() / .hashsizes() / n_tables (kevlar/sketch.py:68) */
int kv_sketch_info(const kv_sketch *s, int *hasher, int *bits, int *ksize, int *n_tables, uint64_t *sizes,
                   int *device);

/* .n_occupied() and .n_unique_kmers() (kevlar/sketch.py:70, kevlar/count.py:84).
 * n_unique is exact (khmer's single-threaded, file-order value) when unique tracking was on
 * for every consume since creation; otherwise *n_unique_valid is 0. */
int kv_sketch_stats(kv_sketch *s, uint64_t *n_occupied, uint64_t *n_unique, int *n_unique_valid);

/* Turn the exact n_unique_kmers bookkeeping on/off (default on; costs two extra passes per
 * table over each batch and a scratch array of 4 bytes per bucket of the LARGEST table, shared
 * per device).  on = 2, "deferred": off, but a consume that fits one chunk already runs table 0's
 * first-touch pass inside its hash kernel, for the kv_unique_last_batch that follows (multi-GPU counts). */
int kv_sketch_set_unique_tracking(kv_sketch *s, int on);

/* Raw table storage, for collectives that the host runs over it (torch.distributed/NCCL)
 * and for tests: device pointer and byte length of table t, khmer layout.
 * All tables of a sketch live in ONE allocation; kv_sketch_flat gives that allocation. */
int kv_sketch_table(kv_sketch *s, int t, void **dev_ptr, uint64_t *nbytes);
int kv_sketch_flat(kv_sketch *s, void **dev_ptr, uint64_t *nbytes);
int kv_sketch_read_table(kv_sketch *s, int t, uint8_t *host_out, uint64_t nbytes);   /* D2H copy */
int kv_sketch_write_table(kv_sketch *s, int t, const uint8_t *host_in, uint64_t nbytes);

/* .consume_seqfile / .consume_seqfile_banding / .consume_seqfile_with_mask /
 * .consume_seqfile_banding_with_mask applied to one batch of reads (kevlar/count.py:50-71,
 * kevlar/sketch.py:148-152) and .consume(seq) (one-read batch).
 *   num_bands <= 0: unbanded; otherwise keep hashes in khmer's band interval (0-based band).
 *   mask == NULL: none; otherwise count a k-mer iff
 *       consume_masked == 0:  mask.get(hash) <= mask_threshold
 *       consume_masked != 0:  mask.get(hash) >= mask_threshold          
 *   n_kmers_out: if non-NULL the call synchronises and returns the number of k-mers counted
 *   (khmer's n_consumed); if NULL the call is asynchronous. */
int kv_consume_batch(kv_sketch *s, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                     int where, int num_bands, int band, const kv_sketch *mask, int mask_threshold,
                     int consume_masked, uint64_t *n_kmers_out);

/* The read loop of kevlar.novel.novel + kmer_is_interesting (kevlar/novel.py:21-53,123-169)
 * for one batch: every k-mer of every read is hashed once and looked up in all case and
 * control sketches in one pass.
 *   screen <= 0: no abundance screen.   num_bands <= 0: no band filter; otherwise keep a
 *   k-mer iff (hash & (num_bands-1)) == band_minus_1 -- the reference's bit test on the
 *   already 0-based band, reproduced as is (kevlar/novel.py:144-147).
 * Outputs (host pointers): hits[0..*n_hits) sorted by (read, offset); read_flags[n_reads];
 * discard_pos[n_reads] (may be NULL when screen <= 0) = offset of the first k-mer that
 * tripped the screen or 0xFFFFFFFF.  Hits of skipped reads are not reported; hits that follow
 * a read's discard position are not reported (the reference breaks out of the read there).
 * Returns KV_EOVERFLOW (and the required count in *n_hits) if max_hits is too small. */
int kv_novel_batch(const kv_sketch *const *cases, int n_case, const kv_sketch *const *ctrls, int n_ctrl,
                   const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads, int where, int case_min,
                   int ctrl_max, int screen, int num_bands, int64_t band_minus_1, kv_hit *hits,
                   uint64_t max_hits, uint64_t *n_hits, uint8_t *read_flags, uint32_t *discard_pos);

/* .hash(kmer) for n k-mers of length ksize laid out back to back (kevlar/novel.py:145).
 * ok[i] = 0 where the k-mer holds a byte outside ACGT (khmer raises there). */
int kv_hash_kmers(int hasher, int ksize, const uint8_t *kmers, uint64_t n, int device, uint64_t *hashes_out,
                  uint8_t *ok_out);

/* .get(hash) / .add(hash) for n hashes, host arrays (kevlar/novel.py:38,48; kevlar/filter.py:32-34,67).
 * kv_add_hashes applies them in array order for the n_unique bookkeeping. */
int kv_get_hashes(const kv_sketch *s, const uint64_t *hashes, uint64_t n, uint8_t *counts_out);
int kv_add_hashes(kv_sketch *s, const uint64_t *hashes, uint64_t n);

/* .get_kmer_counts(seq) / .get_kmer_hashes(seq) for a batch: one value per base position
 * (positions that do not start a k-mer get 0 and valid[i] = 0).  Host outputs, any may be NULL.
 * (kevlar/simlike.py:23-29; kevlar/tests/test_novel.py:76) */
int kv_kmer_counts_batch(const kv_sketch *s, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                         int where, uint64_t *hashes_out, uint8_t *counts_out, uint8_t *valid_out);

/* khmer `counts.abundance_distribution(parser, tracking)` for one batch of reads (kevlar/dist.py:55,
 * SURVEY 8f rank 4): in read order, every k-mer not yet present in `tracking` (any of its buckets
 * empty) is added to it and dist_out[counts.get(kmer)] is incremented.  dist_out[256] (host) is
 * OVERWRITTEN with this batch's histogram -- khmer's list has 65536 entries, all beyond 255 are 0.
 * `tracking` is updated in place (kevlar builds it as Nodetable(k, 1, 1, primes=counts.hashsizes())),
 * so consecutive batches continue one another exactly like consecutive reads of one parser. */
int kv_abund_dist_batch(const kv_sketch *counts, kv_sketch *tracking, const uint8_t *bases, const uint64_t *offsets,
                        uint64_t n_reads, int where, uint64_t *dist_out);

/* Multi-GPU merge of per-GPU partial sketches (plan A).  The host runs the
 * collective (NCCL via torch.distributed) on a widened copy between these two kernels:
 *   kv_sketch_widen:  counters -> one IEEE half (8-bit counters; NCCL has no 16-bit integer
 *                     type and integers <= 2048 are exact in fp16, so sums are exact for up to
 *                     8 ranks), uint8 (4-bit: one per nibble) or uint8 (1-bit: the raw bytes,
 *                     merge = bitwise OR) element per bucket, written to dev_out (device
 *                     pointer, *n_elems elements of *elem_bytes bytes).
 *   kv_sketch_narrow: summed elements -> clamp to 255 / 15 / (bytes as is) and store back.
 * kv_sketch_merge_peers does the same in ONE kernel over peer-mapped tables of the other
 * GPUs (device pointers valid on this device, e.g. from CUDA IPC): saturating add of bytes
 * [byte_lo, byte_hi) of n_peers flat tables into this sketch (multiples of 256; hi = 0 means
 * the whole table).  With each rank merging only its own 1/N slice and then pulling the
 * finished slices of its peers with kv_sketch_copy_from_peer, this is a reduce-scatter +
 * all-gather made of plain NVLink loads, with no staging copy.  Both calls are asynchronous
 * (kv_sync before telling the peers that the data may be read / overwritten). */
int kv_sketch_widen(kv_sketch *s, void *dev_out, uint64_t *n_elems, int *elem_bytes);
int kv_sketch_narrow(kv_sketch *s, const void *dev_in);
int kv_sketch_merge_peers(kv_sketch *s, const void *const *peer_flat, int n_peers, uint64_t byte_lo, uint64_t byte_hi);
int kv_sketch_copy_from_peer(kv_sketch *s, const void *peer_flat, uint64_t byte_lo, uint64_t byte_hi);
/* One-pass all-reduce of a byte slice: like kv_sketch_merge_peers, and the finished slice is also
 * STORED into every peer's table over NVLink.  Correct when every rank calls it on its own,
 * disjoint slice between two rank barriers (nobody else reads or writes a rank's slice): the
 * reduce-scatter and the all-gather of the merge become one kernel with no barrier in between. */
int kv_sketch_allreduce_peers(kv_sketch *s, void *const *peer_flat, int n_peers, uint64_t byte_lo, uint64_t byte_hi);
/* Bin-range-sharded sketches (plan B; needed when one sketch exceeds one GPU): shard i
 * of n holds, of every table, a contiguous range of bins (multiples of 8 bins, so the shards'
 * bytes concatenate to khmer's exact table bytes).  A k-mer's buckets generally live on different
 * shards, so every shard must see the whole hash stream: ranks hash their own reads into device
 * buffers (kv_hash_batch_dev), exchange the buffers (NCCL all-gather, run by the host), and each
 * applies all of them to its ranges (kv_add_hashes_dev; hashes that fall outside the shard's
 * ranges are skipped in-kernel).  Queries return the minimum over the buckets held locally, 255
 * when none is (kv_get_hashes_dev); an element-wise MIN all-reduce over the shards gives the true
 * abundance, which kv_novel_from_counts turns into hits exactly like kv_novel_batch.
 * n_unique_kmers is not available on shards; n_occupied is the local count (sum over shards). */
int kv_sketch_create_shard(int hasher, int bits, int ksize, int n_tables, const uint64_t *sizes, int shard,
                           int n_shards, int device, kv_sketch **out);
int kv_sketch_shard_info(const kv_sketch *s, int *shard, int *n_shards, uint64_t *lo, uint64_t *count);
/* one piece of an OXLI file written cooperatively: 0 header (shard 0, creates the file; n_occupied =
 * sum over shards), 1 size field of table t (shard 0), 2 this shard's bytes of table t (all shards,
 * in shard order), 3 trailer (shard 0) */
int kv_sketch_save_part(kv_sketch *s, const char *path, int piece, int t, uint64_t n_occupied);
/* dev_hashes: capacity u64, dev_valid: capacity/32+1 u32 (device memory of `device`); capacity must cover
 * the batch rounded up to 1024 positions.  Synchronous. */
int kv_hash_batch_dev(int hasher, int ksize, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                      int where, int num_bands, int band, int device, uint64_t *dev_hashes, uint32_t *dev_valid,
                      uint64_t capacity, uint64_t *n_positions, uint64_t *n_kmers);
int kv_add_hashes_dev(kv_sketch *s, const uint64_t *dev_hashes, const uint32_t *dev_valid, uint64_t n);
int kv_get_hashes_dev(const kv_sketch *s, const uint64_t *dev_hashes, const uint32_t *dev_valid, uint64_t n,
                      uint8_t *dev_counts);
/* kv_novel_batch with the abundances of sample i at every base position of the batch given in
 * dev_counts[i] (device pointers, one byte per position); `like` supplies k, hasher and device. */
int kv_novel_from_counts(const kv_sketch *like, int n_case, int n_ctrl, const uint8_t *const *dev_counts,
                         const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads, int where, int case_min,
                         int ctrl_max, int screen, int num_bands, int64_t band_minus_1, kv_hit *hits,
                         uint64_t max_hits, uint64_t *n_hits, uint8_t *read_flags, uint32_t *discard_pos);

/* Spanning sketches (SURVEY 8e plan B, second design; no reference counterpart): a sketch larger than one GPU
 * whose tables live in ONE virtual address range mapped on every rank, piece r of every table in the HBM of
 * rank r (CUDA virtual memory management; pieces are shared between the ranks' processes as POSIX file
 * descriptors).  Every call that reads the sketch -- kv_get_hashes, kv_novel_batch, kv_kmer_counts_batch,
 * kv_sketch_stats, kv_sketch_save, kv_sketch_read_table -- works on it unchanged from any rank (remote
 * pieces are reached over NVLink).  Updates are COLLECTIVE: kv_consume_batch_span takes THIS rank's reads,
 * files their updates by (table, region) in slabs that the other ranks map over CUDA IPC, and the rank
 * holding a region applies the slabs of all ranks in shared memory -- the all-to-all of the update stream
 * (2 bytes per update) is fused into that kernel.  Protocol:
 *   kv_sketch_create_span   on every rank (same arguments but `rank`): returns the handle, one file descriptor
 *                           per table for this rank's physical memory (mem_fds_out[n_tables]) and a
 *                           128-byte handle for its exchange buffers;
 *   kv_sketch_span_attach   once per peer, with the peer's n_tables descriptors (passed over a Unix socket,
 *                           SCM_RIGHTS) and exchange handle;
 *   kv_sketch_span_ready    after all peers are attached.
 * kv_consume_batch_span: every rank passes the same n_chunks >= ceil(its positions / chunk_positions) (the
 * maximum over the ranks; ranks that run out of reads take part with empty slabs) and its kv_peer_sync.
 * kv_sketch_clear clears the local pieces (call it on every rank).  n_unique_kmers is not tracked. */
int kv_sketch_create_span(int hasher, int bits, int ksize, int n_tables, const uint64_t *sizes, int rank, int world,
                          int device, uint64_t chunk_positions, kv_sketch **out, int *mem_fds_out,
                          uint8_t xchg_handle_out[128]);
int kv_sketch_span_attach(kv_sketch *s, int peer_rank, const int *mem_fds, const uint8_t xchg_handle[128]);
int kv_sketch_span_ready(kv_sketch *s);
int kv_sketch_span_info(const kv_sketch *s, int *rank, int *world, uint64_t *piece_bytes, uint64_t *chunk_positions);
struct kv_peer_sync;
int kv_consume_batch_span(kv_sketch *s, struct kv_peer_sync *ps, const uint8_t *bases, const uint64_t *offsets,
                          uint64_t n_reads, int where, uint64_t n_chunks, int num_bands, int band, const kv_sketch *mask,
                          int mask_threshold, int consume_masked, uint64_t *n_kmers_out);

/* n_unique_kmers (kevlar/count.py:84 logs it) when the reads of one sample are sharded over R ranks in file
 * order.  After every rank has counted its shard into a ZEROED partial sketch:
 *   kv_sketch_occupancy   device pointers to the sketch's per-table occupancy bitmaps (1 bit per bucket, LSB
 *                         first; the refresh from the counters is ENQUEUED on kv_stream; owned by the sketch);
 *   the caller ORs the bitmaps of all LOWER ranks into scratch bitmaps of its own (rank 0: all zero);
 *   kv_unique_batch       the first-touch passes over THIS rank's reads with those bitmaps as the occupied set
 *                         (they are updated as the batch's chunks go by) -> this rank's share of n_unique;
 *   kv_sketch_set_unique  stores the sum over the ranks in the merged sketch. */
int kv_sketch_occupancy(kv_sketch *s, uint32_t **dev_words_out, uint64_t *n_words_out);
int kv_unique_batch(const kv_sketch *like, uint32_t *const *dev_occupied, const uint8_t *bases, const uint64_t *offsets,
                    uint64_t n_reads, int where, int num_bands, int band, const kv_sketch *mask, int mask_threshold,
                    int consume_masked, uint64_t *n_unique_out, uint64_t *dev_n_unique_out);
/* kv_unique_batch for the batch that kv_consume_batch counted into `like` LAST on this device, straight from the
 * hashes that call left in the device scratch: no second host-to-device copy of the reads, no second hash.
 * Applies when that batch fitted one chunk and nothing has used the scratch since; otherwise KV_ESTATE and
 * nothing was done -- call kv_unique_batch. */
int kv_unique_last_batch(const kv_sketch *like, uint32_t *const *dev_occupied, uint64_t *n_unique_out,
                         uint64_t *dev_n_unique_out);
int kv_sketch_set_unique(kv_sketch *s, uint64_t n_unique);
/* Stream-ordered forms (no host synchronisation; everything runs on the device's kv_stream): kv_sketch_occupancy
 * only enqueues the refresh; kv_unique_batch with n_unique_out == NULL leaves this rank's share in
 * *dev_n_unique_out (device memory); kv_sketch_set_unique_dev copies the final number from device memory. */
int kv_sketch_set_unique_dev(kv_sketch *s, const uint64_t *dev_n_unique);

/* CUDA IPC plumbing for kv_sketch_merge_peers: export this sketch's flat allocation
 * (64-byte handle) / map a peer's.  */
int kv_sketch_ipc_export(kv_sketch *s, uint8_t handle_out[64]);
int kv_ipc_open(int device, const uint8_t handle[64], void **dev_ptr);
int kv_ipc_close(int device, void *dev_ptr);

/* Device-side barrier between the ranks of a multi-GPU merge (no reference counterpart; it replaces
 * host barriers between the merge phases).  Every rank creates one object -- a small flag array in
 * its HBM, exported as a CUDA IPC handle --, connects the handles of all its peers, and then
 * kv_peer_barrier() ENQUEUES one tiny kernel on the device's compute stream that publishes "I am
 * here" into every peer's array and spins until every peer has done the same: work enqueued after
 * it starts only when all ranks' earlier work is complete and visible.  Calls are collective (same
 * number on every rank).  A peer that does not arrive within the timeout (KV_PEER_TIMEOUT_MS,
 * default 30000) makes the kernel give up; kv_peer_sync_status() then returns KV_ECUDA after
 * draining the stream (0 when every barrier so far completed). */
typedef struct kv_peer_sync kv_peer_sync;
int kv_peer_sync_create(int device, int rank, int world, kv_peer_sync **out, uint8_t handle_out[64]);
int kv_peer_sync_connect(kv_peer_sync *ps, int peer_rank, const uint8_t handle[64]);
int kv_peer_barrier(kv_peer_sync *ps);
int kv_peer_sync_status(kv_peer_sync *ps);
int kv_peer_sync_destroy(kv_peer_sync *ps);
/* The merge lane: a second stream per device for the peer-to-peer merge, so that the merge of one sample's sketch
 * crosses NVLink while the next sample is being counted (kevlar count runs once per sample: three independent
 * sketches per trio, kevlar/workflows/mark-I/Snakefile).  kv_merge_fork(device): work enqueued on the merge lane
 * from now on starts after everything enqueued on kv_stream so far; until kv_merge_join(device) -- which makes
 * kv_stream wait for the lane -- kv_sketch_allreduce_peers / kv_sketch_merge_peers launch on the lane.
 * kv_peer_barrier follows its object: kv_peer_sync_set_lane(ps, 1) binds all barriers of `ps` to the lane (use
 * one object per lane: barriers of one object must execute in the order they were enqueued). */
int kv_peer_sync_set_lane(kv_peer_sync *ps, int lane);
int kv_merge_fork(int device);
int kv_merge_join(int device);

/* khmer.ReadParser(filename) (kevlar/count.py:40, kevlar/__init__.py:125-128): FASTA/FASTQ, plain or
 * gzip.  kv_reader_next parses at least one and at most ~max_bases bases' worth of records, in
 * file order, straight into the batch layout above; n_reads = 0 means end of file.  All output
 * pointers refer to reader-owned host memory that stays valid until the next call on the same
 * reader: names / quals are the header lines (without '@' / '>') and quality strings of the
 * batch back to back, with n_reads+1 offsets each; is_fastq[i] tells whether record i has a
 * quality string.  Any of the text outputs may be NULL; with names == NULL and quals == NULL the
 * header / quality text is not collected at all (the counting path needs sequences only).  File
 * I/O and inflate run in a read-ahead thread owned by the reader.  Not thread-safe per reader. */
typedef struct kv_reader kv_reader;
int kv_reader_open(const char *path, kv_reader **out);
int kv_reader_next(kv_reader *r, uint64_t max_bases, const uint8_t **bases, const uint64_t **offsets,
                   uint64_t *n_reads, const char **names, const uint64_t **name_offsets, const char **quals,
                   const uint64_t **qual_offsets, const uint8_t **is_fastq);
int kv_reader_num_reads(const kv_reader *r, uint64_t *n);
int kv_reader_close(kv_reader *r);
/* The same with batches the caller keeps: kv_reader_next_batch hands out a batch object (*out = NULL at the end
 * of the input) whose arrays (kv_batch_arrays, same meaning as the outputs of kv_reader_next) stay valid until
 * kv_batch_release -- also across further kv_reader_next_batch calls and after kv_reader_close -- so a consumer
 * can pass them to kv_consume_batch / kv_novel_batch without copying while the next batch is being parsed.
 * Released batches are recycled by their reader.  kv_batch_release may be called from any thread.
 * Plain files are memory-mapped and parsed by KV_READER_THREADS threads (default min(8, cores / LOCAL_WORLD_SIZE));
 * BGZF (bgzip) files are mapped too and their blocks inflated by the same threads; an ordinary gzip stream is
 * inflated by one read-ahead thread. */
typedef struct kv_batch kv_batch;
int kv_reader_next_batch(kv_reader *r, uint64_t max_bases, int keep_text, kv_batch **out);
int kv_batch_arrays(const kv_batch *b, const uint8_t **bases, const uint64_t **offsets, uint64_t *n_reads, const char **names,
                    const uint64_t **name_offsets, const char **quals, const uint64_t **qual_offsets, const uint8_t **is_fastq);
int kv_batch_release(kv_batch *b);

/* Measurement fixture, no reference counterpart in the product path: wgsim-style synthetic reads
 * drawn on the device (the recipe of kevlar/tests/data/minitrio/README: fixed-length reads from a
 * random haplotype / position / strand, iid substitution errors), so that inputs of BASELINE
 * configs 3-4 size never cross PCIe.  Read r (global index first_read + r) is a pure function of
 * (seed, r): ranks can generate disjoint slices of one sample.  dev_haplotypes: n_haps device
 * pointers to upper-case ACGT bytes.  dev_offsets (n_reads + 1 u64, device) may be NULL. */
int kv_synth_reads(int device, const uint8_t *const *dev_haplotypes, const uint64_t *hap_lens, int n_haps,
                   uint64_t n_reads, uint64_t first_read, uint32_t read_len, double error_rate, uint64_t seed,
                   uint8_t *dev_bases, uint64_t *dev_offsets);

/* Stream plumbing: the cudaStream_t all work for `device` is enqueued on, so a host
 * framework can record events on it; kv_sync waits for it. */
int kv_stream(int device, void **cuda_stream);
int kv_sync(int device);

/* Per-kernel-class device timing, measured with CUDA events on the launching stream.
 * enable: 1 = start (and reset), 0 = stop (and reset), 2 = read only.  ms_out / n_out receive
 * KV_PROF_CLASSES totals accumulated since the last reset: milliseconds and launch counts for
 * [0] other, [1] hash, [2] increment, [3] unique-tracking probe, [4] novel, [5] merge,
 * [6] overflow fix-up (rollback + exact redo; normally two empty launches per chunk),
 * [7] region partitioning of the updates (large sketches: hist + scan + scatter). */
#define KV_PROF_CLASSES 8
int kv_profile(int device, int enable, double *ms_out, uint64_t *n_out);

/* How many chunks had to be rolled back and redone with the exact update path because the
 * speculative pass saw a counter overflow (diagnostics; tests assert the path is exercised). */
int kv_redo_count(int device, uint64_t *n);

/* Number of kernels this library has launched on `device` since load (bench accounting). */
int kv_launch_count(int device, uint64_t *n);

#ifdef __cplusplus
}
#endif
#endif /* KVSKETCH_H */
