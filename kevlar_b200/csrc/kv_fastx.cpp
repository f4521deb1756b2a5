// kv_fastx.cpp -- native FASTA/FASTQ(.gz) reader behind the kv_reader_* entry points.
//
// Replaces khmer.ReadParser (kevlar/count.py:40, kevlar/__init__.py:125-128) on the host side of
// the boundary: it parses straight into the batch layout the kernels take (concatenated sequence
// bytes + offsets) and, on request, keeps the header / quality text of the batch so that the few
// reads `kevlar novel` reports can be turned back into records.  Same record rules as the Python
// reader in kevlar_b200/fastx.py (which the tests keep as a cross-check): '@' starts a 4-line
// FASTQ record, '>' a FASTA record whose sequence may span lines, blank lines are skipped, CR is
// stripped, the name is the whole header line without its first character.
//
// Plain (uncompressed) regular files are mapped and parsed by several threads at once: every call cuts
// the next stretch of the file into slices that start at record boundaries ('@' line whose next-but-one
// line starts with '+'; '>' line in a FASTA file), each worker parses its slice with the same record rules
// into private arrays, and the pieces are copied into the batch arrays in file order (KV_READER_THREADS,
// default min(8, cores / local ranks)).  One thread parses ~1.1 GB/s of FASTQ text; a B200 counts the k-mers
// of 10 GB/s.
//
// BGZF (bgzip) files take the same route: the compressed file is mapped, the pool inflates its blocks -- each one an
// independent deflate stream whose compressed and uncompressed sizes are in its header and trailer, CRC checked --
// into a text window that ends at the last record start, and the parser above runs over the window (0.3 -> 1.4 GB/s
// of text with 8 threads).
//
// Everything else (ordinary gzip streams, pipes) goes through two stages, so that inflate (the slow part of a .gz input)
// overlaps both the parsing and the GPU work of the caller:
//   producer thread   read() for plain files, zlib for gzip (detected by magic number); fills
//                     4 MB blocks of raw text into a small queue, always a few blocks ahead;
//   kv_reader_next    appends blocks to its text window and splits records.  Complete 4-line FASTQ
//                     records are cut with four memchr calls and copied once, straight into the
//                     batch arrays; everything else (FASTA, the ragged end of the file) goes
//                     through the line-by-line path.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/kvsketch.h"

int kv_fail_public(int code, const char *fmt, ...);   // kvsketch.cu

namespace {

const size_t KV_BLOCK = 4u << 20;   // raw text per producer block
const size_t KV_AHEAD = 4;          // blocks the producer may run ahead

struct Block {
    std::vector<char> data;
    size_t len = 0;
};

// Allocator of the batch arrays: no value-initialisation on resize (every byte is overwritten), and large blocks
// are 2 MB-aligned and marked for transparent huge pages -- a fresh 64 MB batch buffer otherwise costs 16 k page
// faults before the first record lands in it, more than parsing the text does.
template <class T>
struct BatchAlloc {
    using value_type = T;
    BatchAlloc() = default;
    template <class U> BatchAlloc(const BatchAlloc<U> &) {}
    T *allocate(size_t n)
    {
        const size_t bytes = n * sizeof(T), huge = 2u << 20;
        void *p = nullptr;
        if (bytes >= 2 * huge) {
            if (posix_memalign(&p, huge, (bytes + huge - 1) / huge * huge) != 0) throw std::bad_alloc();
            madvise(p, (bytes + huge - 1) / huge * huge, MADV_HUGEPAGE);
        } else if (!(p = malloc(bytes ? bytes : 1)))
            throw std::bad_alloc();
        return (T *)p;
    }
    void deallocate(T *p, size_t) { free(p); }
    template <class U> void construct(U *p) noexcept { ::new ((void *)p) U; }
    template <class U, class... A> void construct(U *p, A &&...a) { ::new ((void *)p) U(std::forward<A>(a)...); }
    template <class U> bool operator==(const BatchAlloc<U> &) const { return true; }
    template <class U> bool operator!=(const BatchAlloc<U> &) const { return false; }
};
template <class T> using BatchVec = std::vector<T, BatchAlloc<T>>;

// what one worker made of one slice of a mapped file
struct Piece {
    BatchVec<uint8_t> bases;
    BatchVec<uint64_t> ends, name_ends, qual_ends;   // cumulative, local to the piece
    BatchVec<char> names, quals;
    BatchVec<uint8_t> is_fastq;
    bool mode_flip = false;   // a '>' header at record position inside a FASTQ file: the rest is FASTA for the serial rules
    void clear() { bases.clear(); ends.clear(); name_ends.clear(); qual_ends.clear(); names.clear(); quals.clear(); is_fastq.clear(); mode_flip = false; }
};

// a FASTA record whose end (next header or EOF) has not been seen yet
struct ParseState {
    bool fasta_open = false;
    std::string name, seq;
};

}   // namespace

struct kv_reader;

// One batch in the layout the kernels take.  A batch handed out by kv_reader_next_batch belongs to the caller
// until kv_batch_release: the reader recycles released batches (their buffers keep their capacity -- and their
// huge pages), and a batch that outlives its reader frees itself on release.
struct kv_batch {
    BatchVec<uint8_t> bases;
    BatchVec<uint64_t> offsets, name_offsets, qual_offsets;
    BatchVec<char> names, quals;
    BatchVec<uint8_t> is_fastq;   // per record: 1 = FASTQ (has a quality string), 0 = FASTA
    kv_reader *owner = nullptr;
    bool leased = false;
};

static std::mutex g_batch_mu;   // free lists and lease bookkeeping of all readers (release may come from any thread)

struct kv_reader {
    int fd = -1;
    gzFile gz = nullptr;
    std::string path;
    // producer side
    std::thread producer;
    std::mutex mu;
    std::condition_variable cv_full, cv_free;
    std::deque<Block *> full, spare;
    bool input_done = false, stop = false;
    // text window of the parser: unread bytes are buf[pos, end)
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    std::string io_error;            // set by the producer: the input ended because of an I/O / inflate error
    uint64_t num_reads = 0;
    // a FASTA record whose end (next header or EOF) has not been seen yet
    bool fasta_open = false;
    std::string fasta_name, fasta_seq;
    // the batch handed out by the last kv_reader_next call
    bool keep_text = true;
    kv_batch *cur = nullptr;                  // the batch being filled / handed out by the last kv_reader_next call
    std::vector<kv_batch *> spare_batches;    // released, ready for reuse
    std::vector<kv_batch *> leased_batches;   // handed out by kv_reader_next_batch, not yet released
    // mapped plain file, parsed in parallel
    const char *map = nullptr;
    size_t map_len = 0, mpos = 0;
    int mode = 0;                    // '@' FASTQ file, '>' FASTA file (first record decides)
    bool serial_rest = false;        // a FASTQ file turned FASTA half-way: one thread, carried state, from here on
    ParseState mstate;
    int n_threads = 1;
    size_t slice_bytes = 8u << 20;
    std::vector<Piece> pieces;
    std::vector<std::thread> pool;
    std::mutex pool_mu;
    std::condition_variable pool_go, pool_done;
    uint64_t pool_epoch = 0;
    int pool_pending = 0;
    bool pool_stop = false;
    std::function<void(int)> pool_job;
    // BGZF (bgzip) input: the compressed file is mapped, its blocks are inflated by the pool into `text`, and the
    // parser above works on windows of that text: map = text.data(), map_len = end of the last complete record
    void *file_map = nullptr;        // what munmap gets (the text of a plain file, the compressed bytes of a BGZF file)
    size_t file_map_len = 0;
    const uint8_t *cmap = nullptr;
    size_t cmap_len = 0, cpos = 0;
    bool bgzf = false, c_eof = false;
    BatchVec<char> text;
    size_t text_len = 0;
};

static void producer_main(kv_reader *r)
{
    for (;;) {
        Block *b = nullptr;
        {
            std::unique_lock<std::mutex> lk(r->mu);
            r->cv_free.wait(lk, [r] { return r->stop || !r->spare.empty() || r->full.size() < KV_AHEAD; });
            if (r->stop) return;
            if (!r->spare.empty()) { b = r->spare.back(); r->spare.pop_back(); }
        }
        if (!b) { b = new Block(); b->data.resize(KV_BLOCK); }
        long got = r->gz ? (long)gzread(r->gz, b->data.data(), (unsigned)KV_BLOCK) : (long)read(r->fd, b->data.data(), KV_BLOCK);
        // a read error or a damaged / truncated gzip stream is an ERROR, not the end of the file: khmer's
        // ReadParser raises there, and a sketch built from part of the input would be silently wrong
        std::string problem;
        if (got < 0) problem = r->gz ? "gzip stream is damaged" : "read error";
        else if (got == 0 && r->gz) {
            int zerr = Z_OK;
            const char *msg = gzerror(r->gz, &zerr);
            if (zerr != Z_OK && zerr != Z_STREAM_END) problem = msg && *msg ? msg : "gzip stream is truncated or damaged";
        }
        std::unique_lock<std::mutex> lk(r->mu);
        if (got <= 0) {
            r->spare.push_back(b);
            r->io_error = problem;
            r->input_done = true;
            r->cv_full.notify_all();
            return;
        }
        b->len = (size_t)got;
        r->full.push_back(b);
        r->cv_full.notify_all();
    }
}

// append the next raw block to the window; false at end of input
static bool reader_fill(kv_reader *r)
{
    if (r->eof) return false;
    Block *b = nullptr;
    {
        std::unique_lock<std::mutex> lk(r->mu);
        r->cv_full.wait(lk, [r] { return !r->full.empty() || r->input_done; });
        if (r->full.empty()) { r->eof = true; return false; }
        b = r->full.front();
        r->full.pop_front();
    }
    if (r->pos > 0) {
        memmove(r->buf.data(), r->buf.data() + r->pos, r->end - r->pos);
        r->end -= r->pos;
        r->pos = 0;
    }
    if (r->buf.size() - r->end < b->len) r->buf.resize(r->end + b->len + KV_BLOCK);
    memcpy(r->buf.data() + r->end, b->data.data(), b->len);
    r->end += b->len;
    {
        std::unique_lock<std::mutex> lk(r->mu);
        r->spare.push_back(b);
    }
    r->cv_free.notify_one();
    return true;
}

// next line without its terminator (and without a trailing CR); false at end of input
static bool reader_line(kv_reader *r, const char **line, size_t *len)
{
    for (;;) {
        const char *start = r->buf.data() + r->pos;
        const char *nl = (const char *)memchr(start, '\n', r->end - r->pos);
        if (nl) {
            *line = start;
            *len = (size_t)(nl - start);
            r->pos += *len + 1;
            break;
        }
        if (!reader_fill(r)) {
            if (r->pos == r->end) return false;
            *line = r->buf.data() + r->pos;      // last line without newline
            *len = r->end - r->pos;
            r->pos = r->end;
            break;
        }
    }
    if (*len && (*line)[*len - 1] == '\r') --*len;
    return true;
}

static inline void reader_emit(kv_reader *r, const char *name, size_t nlen, const char *seq, size_t slen, const char *qual,
                               size_t qlen, bool fastq)
{
    r->cur->bases.insert(r->cur->bases.end(), (const uint8_t *)seq, (const uint8_t *)seq + slen);
    r->cur->offsets.push_back(r->cur->bases.size());
    if (r->keep_text) {
        r->cur->names.insert(r->cur->names.end(), name, name + nlen);
        if (fastq) r->cur->quals.insert(r->cur->quals.end(), qual, qual + qlen);
    }
    r->cur->name_offsets.push_back(r->cur->names.size());
    r->cur->qual_offsets.push_back(r->cur->quals.size());
    r->cur->is_fastq.push_back(fastq ? 1 : 0);
    r->num_reads++;
}

// A complete 4-line FASTQ record at the head of the window: cut it with four memchr calls and
// emit it without intermediate copies.  Returns false (window untouched) when the window does not
// hold four newlines from here -- the caller then refills or, at end of input, takes the
// line-by-line path, which copes with a truncated record.
static inline bool reader_fastq_fast(kv_reader *r)
{
    const char *p = r->buf.data() + r->pos, *end = r->buf.data() + r->end;
    const char *nl[4];
    const char *q = p;
    for (int i = 0; i < 4; i++) {
        nl[i] = (const char *)memchr(q, '\n', (size_t)(end - q));
        if (!nl[i]) return false;
        q = nl[i] + 1;
    }
    auto strip = [](const char *b, const char *e) { return (size_t)(e - b) - ((e > b && e[-1] == '\r') ? 1 : 0); };
    reader_emit(r, p + 1, strip(p + 1, nl[0]), nl[0] + 1, strip(nl[0] + 1, nl[1]), nl[2] + 1, strip(nl[2] + 1, nl[3]), true);
    r->pos = (size_t)(q - r->buf.data());
    return true;
}

// ---------------------------------------------------------------- mapped files: parallel parsing

static inline void piece_emit(Piece &o, bool keep_text, const char *name, size_t nlen, const char *seq, size_t slen,
                              const char *qual, size_t qlen, bool fastq)
{
    o.bases.insert(o.bases.end(), (const uint8_t *)seq, (const uint8_t *)seq + slen);
    o.ends.push_back(o.bases.size());
    if (keep_text) {
        o.names.insert(o.names.end(), name, name + nlen);
        if (fastq) o.quals.insert(o.quals.end(), qual, qual + qlen);
    }
    o.name_ends.push_back(o.names.size());
    o.qual_ends.push_back(o.quals.size());
    o.is_fastq.push_back(fastq ? 1 : 0);
}

// The record rules of kv_reader_next over a memory range.  Parses [p, end) until the piece holds max_bases bases
// (checked between records) and returns where it stopped.  `final`: `end` is a record boundary or the end of the
// file, so an open FASTA record ends there.  `fastq_only`: stop with mode_flip at a '>' header (see Piece).
static const char *parse_range(const char *p, const char *end, ParseState &st, bool final, uint64_t max_bases, bool keep_text,
                               Piece &o, bool fastq_only)
{
    auto next_line = [&](const char *&line, size_t &len) -> bool {
        if (p >= end) return false;
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        line = p;
        len = nl ? (size_t)(nl - p) : (size_t)(end - p);
        p = nl ? nl + 1 : end;
        if (len && line[len - 1] == '\r') --len;
        return true;
    };
    auto strip = [](const char *b, const char *e) { return (size_t)(e - b) - ((e > b && e[-1] == '\r') ? 1 : 0); };
    while (p < end && (o.bases.size() < max_bases || o.ends.empty())) {
        if (!st.fasta_open && *p == '@') {   // a complete 4-line record: four memchr calls, one copy
            const char *nl[4], *q = p;
            int i = 0;
            for (; i < 4; i++) {
                nl[i] = (const char *)memchr(q, '\n', (size_t)(end - q));
                if (!nl[i]) break;
                q = nl[i] + 1;
            }
            if (i == 4) {
                piece_emit(o, keep_text, p + 1, strip(p + 1, nl[0]), nl[0] + 1, strip(nl[0] + 1, nl[1]), nl[2] + 1,
                           strip(nl[2] + 1, nl[3]), true);
                p = q;
                continue;
            }
        }
        const char *line;
        size_t len;
        if (!next_line(line, len)) break;
        if (len == 0) continue;
        if (line[0] == '@' && !st.fasta_open) {   // the ragged end: whatever lines are left
            const char *name = line + 1, *seq = nullptr, *qual = nullptr;
            size_t nlen = len - 1, slen = 0, qlen = 0, unused;
            const char *tmp;
            if (next_line(seq, slen)) {
                if (next_line(tmp, unused)) {
                    if (!next_line(qual, qlen)) { qual = nullptr; qlen = 0; }
                }
            } else { seq = nullptr; slen = 0; }
            piece_emit(o, keep_text, name, nlen, seq ? seq : name, slen, qual ? qual : name, qlen, true);
        } else if (line[0] == '>') {
            if (fastq_only) { o.mode_flip = true; return line; }
            if (st.fasta_open) piece_emit(o, keep_text, st.name.data(), st.name.size(), st.seq.data(), st.seq.size(), nullptr, 0, false);
            st.fasta_open = true;
            st.name.assign(line + 1, len - 1);
            st.seq.clear();
        } else if (st.fasta_open) {
            st.seq.append(line, len);
        }
    }
    if (p >= end && final && st.fasta_open) {
        piece_emit(o, keep_text, st.name.data(), st.name.size(), st.seq.data(), st.seq.size(), nullptr, 0, false);
        st.fasta_open = false;
    }
    return p;
}

// first record start at or after `from` (which need not be a line start unless it is the start of the map),
// or `limit` if there is none before it
static const char *next_record_start(const char *base, const char *from, const char *limit, int mode)
{
    const char *p = from;
    if (p > base && p[-1] != '\n') {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(limit - p));
        if (!nl) return limit;
        p = nl + 1;
    }
    while (p < limit) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(limit - p));
        if (mode == '>') {
            if (*p == '>') return p;
        } else if (*p == '@' && nl) {
            // a header, not a quality string that happens to start with '@': the line after next starts with '+'
            const char *nl2 = (const char *)memchr(nl + 1, '\n', (size_t)(limit - (nl + 1)));
            if (nl2 && nl2 + 1 < limit && nl2[1] == '+') return p;
        }
        if (!nl) return limit;
        p = nl + 1;
    }
    return limit;
}

static void pool_main(kv_reader *r, int id)
{
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(r->pool_mu);
            r->pool_go.wait(lk, [&] { return r->pool_stop || r->pool_epoch != seen; });
            if (r->pool_stop) return;
            seen = r->pool_epoch;
        }
        r->pool_job(id);
        std::unique_lock<std::mutex> lk(r->pool_mu);
        if (--r->pool_pending == 0) r->pool_done.notify_all();
    }
}

// run job(0..n_threads-1): worker 0 is the calling thread
static void pool_run(kv_reader *r, const std::function<void(int)> &job)
{
    if (r->n_threads <= 1) { job(0); return; }
    if (r->pool.empty())
        for (int i = 1; i < r->n_threads; i++) r->pool.emplace_back(pool_main, r, i);
    {
        std::unique_lock<std::mutex> lk(r->pool_mu);
        r->pool_job = job;
        r->pool_pending = r->n_threads - 1;
        r->pool_epoch++;
    }
    r->pool_go.notify_all();
    job(0);
    std::unique_lock<std::mutex> lk(r->pool_mu);
    r->pool_done.wait(lk, [&] { return r->pool_pending == 0; });
}

// append the pieces to the batch arrays, in order; the copies run on the pool
static void append_pieces(kv_reader *r, int n_pieces)
{
    std::vector<size_t> b0(n_pieces + 1), n0(n_pieces + 1), t0(n_pieces + 1), q0(n_pieces + 1);
    b0[0] = r->cur->bases.size(); n0[0] = r->cur->offsets.size() - 1; t0[0] = r->cur->names.size(); q0[0] = r->cur->quals.size();
    for (int i = 0; i < n_pieces; i++) {
        const Piece &pc = r->pieces[i];
        b0[i + 1] = b0[i] + pc.bases.size(); n0[i + 1] = n0[i] + pc.ends.size();
        t0[i + 1] = t0[i] + pc.names.size(); q0[i + 1] = q0[i] + pc.quals.size();
    }
    r->cur->bases.resize(b0[n_pieces]);
    r->cur->offsets.resize(n0[n_pieces] + 1); r->cur->name_offsets.resize(n0[n_pieces] + 1); r->cur->qual_offsets.resize(n0[n_pieces] + 1);
    r->cur->is_fastq.resize(n0[n_pieces]);
    r->cur->names.resize(t0[n_pieces]); r->cur->quals.resize(q0[n_pieces]);
    pool_run(r, [&](int id) {
        for (int i = id; i < n_pieces; i += r->n_threads) {
            const Piece &pc = r->pieces[i];
            const size_t n = pc.ends.size();
            if (!pc.bases.empty()) memcpy(r->cur->bases.data() + b0[i], pc.bases.data(), pc.bases.size());
            if (!pc.names.empty()) memcpy(r->cur->names.data() + t0[i], pc.names.data(), pc.names.size());
            if (!pc.quals.empty()) memcpy(r->cur->quals.data() + q0[i], pc.quals.data(), pc.quals.size());
            if (n) memcpy(r->cur->is_fastq.data() + n0[i], pc.is_fastq.data(), n);
            for (size_t k = 0; k < n; k++) {
                r->cur->offsets[n0[i] + 1 + k] = b0[i] + pc.ends[k];
                r->cur->name_offsets[n0[i] + 1 + k] = t0[i] + pc.name_ends[k];
                r->cur->qual_offsets[n0[i] + 1 + k] = q0[i] + pc.qual_ends[k];
            }
        }
    });
    r->num_reads += n0[n_pieces] - n0[0];
}

// ---------------------------------------------------------------- BGZF: block-parallel inflate

// a BGZF block header at p (RFC 1952 member with the 'BC' extra subfield holding the block size - 1)?
static inline size_t bgzf_block_len(const uint8_t *p, size_t avail)
{
    if (avail < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
    const size_t xlen = p[10] | ((size_t)p[11] << 8);
    if (avail < 12 + xlen) return 0;
    for (size_t i = 12; i + 4 <= 12 + xlen;) {
        const size_t slen = p[i + 2] | ((size_t)p[i + 3] << 8);
        if (p[i] == 'B' && p[i + 1] == 'C' && slen == 2 && i + 6 <= 12 + xlen) return (size_t)(p[i + 4] | ((size_t)p[i + 5] << 8)) + 1;
        i += 4 + slen;
    }
    return 0;
}

// start of the last record in text[from, len) whose start can be recognised with the data at hand (or `from`)
static size_t last_record_start(const char *text, size_t from, size_t len, int mode)
{
    size_t end = len;
    for (;;) {
        const char *nl = end > from ? (const char *)memrchr(text + from, '\n', end - from) : nullptr;
        const size_t line = nl ? (size_t)(nl - text) + 1 : from;
        if (line < len) {
            if (mode == '>') {
                if (text[line] == '>') return line;
            } else if (text[line] == '@') {
                const char *n1 = (const char *)memchr(text + line, '\n', len - line);
                const char *n2 = n1 ? (const char *)memchr(n1 + 1, '\n', len - (size_t)(n1 + 1 - text)) : nullptr;
                if (n2 && (size_t)(n2 + 1 - text) < len && n2[1] == '+') return line;
            }
        }
        if (!nl) return from;
        end = (size_t)(nl - text);
    }
}

// Make the next window of text: carry the unparsed tail to the front, inflate further blocks behind it (all threads),
// and end the window at the last record start -- or at the end of the text when the file is exhausted.  Returns
// false on a damaged file (io_error set).
static bool bgzf_refill(kv_reader *r)
{
    const size_t tail = r->text_len - r->mpos;   // what the parser has not consumed: the rest of the window and the text behind it
    if (tail && r->mpos) memmove(r->text.data(), r->text.data() + r->mpos, tail);
    r->text_len = tail;
    r->map_len = 0;
    r->mpos = 0;
    const size_t target = std::max<size_t>((size_t)r->n_threads * r->slice_bytes, 1u << 20);
    struct Blk { size_t coff, clen, uoff, ulen; };
    for (;;) {
        std::vector<Blk> blocks;
        size_t fresh = 0;
        while (r->cpos < r->cmap_len && fresh < target) {
            const size_t blen = bgzf_block_len(r->cmap + r->cpos, r->cmap_len - r->cpos);
            if (blen < 26 || r->cpos + blen > r->cmap_len) { r->io_error = "damaged or truncated BGZF block"; return false; }
            const uint8_t *t = r->cmap + r->cpos + blen - 4;
            const size_t ulen = t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
            blocks.push_back({r->cpos, blen, r->text_len + fresh, ulen});
            fresh += ulen;
            r->cpos += blen;
        }
        if (r->cpos >= r->cmap_len) r->c_eof = true;
        if (r->text.size() < r->text_len + fresh) r->text.resize(r->text_len + fresh + (1u << 20));
        std::atomic<bool> bad(false);
        char *text = r->text.data();
        const uint8_t *cmap = r->cmap;
        const int nb = (int)blocks.size();
        pool_run(r, [&](int id) {
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) { bad = true; return; }
            for (int i = id; i < nb; i += r->n_threads) {
                const Blk &b = blocks[i];
                const uint8_t *p = cmap + b.coff;
                const size_t hdr = 12 + (p[10] | ((size_t)p[11] << 8));
                if (hdr + 8 > b.clen) { bad = true; break; }
                inflateReset(&zs);
                zs.next_in = (Bytef *)(p + hdr);
                zs.avail_in = (uInt)(b.clen - hdr - 8);
                zs.next_out = (Bytef *)(text + b.uoff);
                zs.avail_out = (uInt)b.ulen;
                const int rc = inflate(&zs, Z_FINISH);
                const uint8_t *t = p + b.clen - 8;
                const uint32_t want_crc = t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
                if (rc != Z_STREAM_END || zs.total_out != b.ulen ||
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)(text + b.uoff), (uInt)b.ulen) != want_crc) { bad = true; break; }
            }
            inflateEnd(&zs);
        });
        if (bad) { r->io_error = "damaged BGZF block (inflate or CRC failure)"; return false; }
        r->text_len += fresh;
        r->map = r->text.data();
        if (r->mode == 0 && !r->serial_rest) {   // the first record decides how record boundaries are recognised
            size_t i = 0;
            while (i < r->text_len && (r->text[i] == '\n' || r->text[i] == '\r')) i++;
            if (i < r->text_len) {
                if (r->text[i] == '@' || r->text[i] == '>') r->mode = r->text[i];
                else r->serial_rest = true;
            }
        }
        if (r->c_eof) { r->map_len = r->text_len; return true; }
        if (!r->serial_rest && r->mode) {
            r->map_len = last_record_start(r->text.data(), 0, r->text_len, r->mode);
            if (r->map_len > 0) return true;
        }
        // no complete record yet (a FASTA record longer than the window), or serial rules, which need the whole rest
        // of the text before them: keep inflating
    }
}

// one round of the parser over the current window: up to n_threads record-aligned slices (or one serial stretch)
static void mapped_round(kv_reader *r, uint64_t max_bases)
{
    const char *base = r->map, *file_end = r->map + r->map_len;
    const uint64_t room = max_bases > r->cur->bases.size() ? max_bases - r->cur->bases.size() : 1;
    if (r->serial_rest || r->mode == 0) {   // one thread, carried state
        Piece &pc = r->pieces[0];
        pc.clear();
        const char *stop = parse_range(base + r->mpos, file_end, r->mstate, true, room, r->keep_text, pc, false);
        r->mpos = (size_t)(stop - base);
        append_pieces(r, 1);
        return;
    }
    // text per base: ~2.2 in FASTQ (header, '+', qualities), ~1 in FASTA; aim at the room that is left
    const size_t want = (size_t)std::min<uint64_t>(r->map_len - r->mpos, std::max<uint64_t>(256u << 10, room * (r->mode == '@' ? 2 : 1)));
    const size_t round = std::min(want, (size_t)r->n_threads * r->slice_bytes);
    int n = (int)std::min<size_t>((size_t)r->n_threads, std::max<size_t>(1, round / (64u << 10)));
    const char *round_end = r->mpos + round >= r->map_len ? file_end : next_record_start(base, base + r->mpos + round, file_end, r->mode);
    std::vector<const char *> cut(n + 1);
    cut[0] = base + r->mpos;
    cut[n] = round_end;
    const size_t span = (size_t)(round_end - cut[0]);
    for (int i = 1; i < n; i++) {
        const char *c = next_record_start(base, cut[0] + span / n * i, round_end, r->mode);
        cut[i] = std::max(c, cut[i - 1]);
    }
    pool_run(r, [&](int id) {
        for (int i = id; i < n; i += r->n_threads) {
            Piece &pc = r->pieces[i];
            pc.clear();
            ParseState st;
            if (cut[i] < cut[i + 1]) parse_range(cut[i], cut[i + 1], st, true, UINT64_MAX, r->keep_text, pc, r->mode == '@');
        }
    });
    bool flip = false;
    for (int i = 0; i < n; i++) flip = flip || r->pieces[i].mode_flip;
    if (flip) {   // rare: redo this stretch -- and everything after it -- with the serial rules
        r->serial_rest = true;
        return;
    }
    append_pieces(r, n);
    r->mpos = (size_t)(round_end - base);
}

// kv_reader_next for a mapped file: rounds of up to n_threads slices until the batch is full
static void mapped_next(kv_reader *r, uint64_t max_bases)
{
    if (r->pieces.size() < (size_t)r->n_threads) r->pieces.resize(r->n_threads);
    for (;;) {
        // BGZF: a new window when this one is used up -- or when the serial rules took over, which cannot stop at a
        // window end that the FASTQ boundary rule chose
        if (r->bgzf && !(r->c_eof && r->map_len == r->text_len) && (r->mpos >= r->map_len || r->serial_rest)) {
            if (!bgzf_refill(r)) { r->eof = true; return; }
        }
        if (!(r->mpos < r->map_len && (r->cur->bases.size() < max_bases || r->cur->offsets.size() == 1))) break;
        mapped_round(r, max_bases);
    }
    if (r->mpos >= r->map_len && (!r->bgzf || r->c_eof)) r->eof = true;
}



extern "C" int kv_reader_open(const char *path, kv_reader **out)
{
    if (!path || !out) return kv_fail_public(KV_EINVAL, "null argument");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return kv_fail_public(KV_EIO, "cannot open %s", path);
    unsigned char magic[32];
    memset(magic, 0, sizeof magic);
    ssize_t got = read(fd, magic, sizeof magic);
    lseek(fd, 0, SEEK_SET);
    kv_reader *r = new kv_reader();
    r->path = path;
    const bool gzip = got >= 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    const bool bgzf = gzip && bgzf_block_len(magic, (size_t)got) > 0;
    struct stat sb;
    const char *off = getenv("KV_READER_NO_MMAP");
    // plain text and BGZF files are mapped and handled by the thread pool; everything else streams through zlib / read()
    if ((!gzip || bgzf) && !(off && *off && *off != '0') && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
        void *m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) {
            madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL);
            r->fd = fd;
            r->file_map = m;
            r->file_map_len = (size_t)sb.st_size;
            unsigned hw = std::thread::hardware_concurrency();
            int ranks = 1;
            if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
            r->n_threads = (int)std::max(1u, std::min(8u, (hw ? hw : 1u) / (unsigned)ranks));
            if (const char *e = getenv("KV_READER_THREADS")) r->n_threads = std::max(1, std::min(64, atoi(e)));
            if (const char *e = getenv("KV_READER_SLICE_BYTES")) r->slice_bytes = (size_t)std::max(64ll, atoll(e));
            if (bgzf) {
                r->bgzf = true;
                r->cmap = (const uint8_t *)m;
                r->cmap_len = (size_t)sb.st_size;
                r->text.resize(1u << 20);
                r->map = r->text.data();   // (non-null: the mapped path; the first kv_reader_next call inflates the first window)
                r->map_len = 0;
            } else {
                r->map = (const char *)m;
                r->map_len = (size_t)sb.st_size;
                // the first record decides how record boundaries are recognised; anything else: one thread
                size_t i = 0;
                while (i < r->map_len && (r->map[i] == '\n' || r->map[i] == '\r')) i++;
                r->mode = i < r->map_len && (r->map[i] == '@' || r->map[i] == '>') ? r->map[i] : 0;
            }
            *out = r;
            return KV_OK;
        }
    }
    if (gzip) {
        r->gz = gzdopen(fd, "rb");
        if (!r->gz) { close(fd); delete r; return kv_fail_public(KV_EIO, "cannot open %s", path); }
        gzbuffer(r->gz, 1u << 20);
    } else {
        r->fd = fd;
#ifdef POSIX_FADV_SEQUENTIAL
        posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
    }
    r->buf.resize(2 * KV_BLOCK);
    r->producer = std::thread(producer_main, r);
    *out = r;
    return KV_OK;
}

extern "C" int kv_reader_close(kv_reader *r)
{
    if (!r) return KV_OK;
    {
        std::unique_lock<std::mutex> lk(r->mu);
        r->stop = true;
    }
    r->cv_free.notify_all();
    if (r->producer.joinable()) r->producer.join();
    {
        std::unique_lock<std::mutex> lk(r->pool_mu);
        r->pool_stop = true;
    }
    r->pool_go.notify_all();
    for (std::thread &t : r->pool) t.join();
    if (r->file_map) munmap(r->file_map, r->file_map_len);
    {
        std::unique_lock<std::mutex> lk(g_batch_mu);
        for (kv_batch *b : r->leased_batches) b->owner = nullptr;   // they free themselves when released
        for (kv_batch *b : r->spare_batches) delete b;
        delete r->cur;
    }
    for (Block *b : r->full) delete b;
    for (Block *b : r->spare) delete b;
    if (r->gz) gzclose(r->gz);
    if (r->fd >= 0) close(r->fd);
    delete r;
    return KV_OK;
}

// a batch to fill: the current one if the caller never took it over, else a recycled or a new one
static void reader_take_batch(kv_reader *r)
{
    if (r->cur) return;
    std::unique_lock<std::mutex> lk(g_batch_mu);
    if (!r->spare_batches.empty()) {
        r->cur = r->spare_batches.back();
        r->spare_batches.pop_back();
    } else {
        r->cur = new kv_batch();
        r->cur->owner = r;
    }
}

// parse the next batch into r->cur (n_reads = 0: end of input)
static int reader_fill_batch(kv_reader *r, uint64_t max_bases)
{
    reader_take_batch(r);
    r->cur->bases.clear(); r->cur->names.clear(); r->cur->quals.clear();
    r->cur->offsets.assign(1, 0); r->cur->name_offsets.assign(1, 0); r->cur->qual_offsets.assign(1, 0);
    r->cur->is_fastq.clear();
    if (r->cur->bases.capacity() < max_bases && max_bases <= (1ull << 31)) r->cur->bases.reserve((size_t)max_bases + (1u << 16));
    if (r->map) {
        mapped_next(r, max_bases);
        if (!r->io_error.empty()) return kv_fail_public(KV_EIO, "%s: %s", r->path.c_str(), r->io_error.c_str());
        return KV_OK;
    }
    const char *line;
    size_t len;
    std::string name, seq;   // slow path: lines are copied, the window may move while reading the record
    while (r->cur->bases.size() < max_bases || r->cur->offsets.size() == 1) {
        // fast path: the window starts with a complete FASTQ record
        if (!r->fasta_open && r->pos < r->end && r->buf[r->pos] == '@') {
            if (reader_fastq_fast(r)) continue;
            if (!r->eof && reader_fill(r)) continue;   // more text arrived: try again
        }
        if (!reader_line(r, &line, &len)) {
            if (r->fasta_open) {
                reader_emit(r, r->fasta_name.data(), r->fasta_name.size(), r->fasta_seq.data(), r->fasta_seq.size(),
                            nullptr, 0, false);
                r->fasta_open = false;
            }
            break;
        }
        if (len == 0) continue;
        if (line[0] == '@' && !r->fasta_open) {
            name.assign(line + 1, len - 1);
            seq.clear();
            std::string qual;
            if (reader_line(r, &line, &len)) seq.assign(line, len);
            if (reader_line(r, &line, &len)) { /* '+' line */ }
            if (reader_line(r, &line, &len)) qual.assign(line, len);
            reader_emit(r, name.data(), name.size(), seq.data(), seq.size(), qual.data(), qual.size(), true);
        } else if (line[0] == '>') {
            std::string next_name(line + 1, len - 1);
            if (r->fasta_open)
                reader_emit(r, r->fasta_name.data(), r->fasta_name.size(), r->fasta_seq.data(), r->fasta_seq.size(),
                            nullptr, 0, false);
            r->fasta_open = true;
            r->fasta_name.swap(next_name);
            r->fasta_seq.clear();
        } else if (r->fasta_open) {
            r->fasta_seq.append(line, len);
        }
    }
    if (r->eof) {
        std::string problem;
        {
            std::unique_lock<std::mutex> lk(r->mu);
            problem = r->io_error;
        }
        if (!problem.empty()) return kv_fail_public(KV_EIO, "%s: %s", r->path.c_str(), problem.c_str());
    }
    return KV_OK;
}

static void batch_outputs(const kv_batch *b, const uint8_t **bases, const uint64_t **offsets, uint64_t *n_reads, const char **names,
                          const uint64_t **name_offsets, const char **quals, const uint64_t **qual_offsets, const uint8_t **is_fastq)
{
    if (bases) *bases = b->bases.data();
    if (offsets) *offsets = b->offsets.data();
    if (n_reads) *n_reads = b->offsets.size() - 1;
    if (names) *names = b->names.data();
    if (name_offsets) *name_offsets = b->name_offsets.data();
    if (quals) *quals = b->quals.data();
    if (qual_offsets) *qual_offsets = b->qual_offsets.data();
    if (is_fastq) *is_fastq = b->is_fastq.data();
}

extern "C" int kv_reader_next(kv_reader *r, uint64_t max_bases, const uint8_t **bases, const uint64_t **offsets,
                              uint64_t *n_reads, const char **names, const uint64_t **name_offsets, const char **quals,
                              const uint64_t **qual_offsets, const uint8_t **is_fastq)
{
    if (!r || !bases || !offsets || !n_reads) return kv_fail_public(KV_EINVAL, "null argument");
    r->keep_text = names != nullptr || quals != nullptr;   // sequences only: skip the header / quality copies
    int rc = reader_fill_batch(r, max_bases);
    if (rc != KV_OK) return rc;
    batch_outputs(r->cur, bases, offsets, n_reads, names, name_offsets, quals, qual_offsets, is_fastq);
    return KV_OK;
}

extern "C" int kv_reader_next_batch(kv_reader *r, uint64_t max_bases, int keep_text, kv_batch **out)
{
    if (!r || !out) return kv_fail_public(KV_EINVAL, "null argument");
    *out = nullptr;
    r->keep_text = keep_text != 0;
    int rc = reader_fill_batch(r, max_bases);
    if (rc != KV_OK) return rc;
    if (r->cur->offsets.size() <= 1) return KV_OK;   // end of input: nothing handed out
    std::unique_lock<std::mutex> lk(g_batch_mu);
    r->cur->leased = true;
    r->leased_batches.push_back(r->cur);
    *out = r->cur;
    r->cur = nullptr;
    return KV_OK;
}

extern "C" int kv_batch_arrays(const kv_batch *b, const uint8_t **bases, const uint64_t **offsets, uint64_t *n_reads, const char **names,
                               const uint64_t **name_offsets, const char **quals, const uint64_t **qual_offsets,
                               const uint8_t **is_fastq)
{
    if (!b) return kv_fail_public(KV_EINVAL, "null argument");
    batch_outputs(b, bases, offsets, n_reads, names, name_offsets, quals, qual_offsets, is_fastq);
    return KV_OK;
}

extern "C" int kv_batch_release(kv_batch *b)
{
    if (!b) return KV_OK;
    std::unique_lock<std::mutex> lk(g_batch_mu);
    if (!b->leased) return kv_fail_public(KV_EINVAL, "batch is not on loan");
    b->leased = false;
    kv_reader *r = b->owner;
    if (!r) { lk.unlock(); delete b; return KV_OK; }   // its reader is gone
    r->leased_batches.erase(std::find(r->leased_batches.begin(), r->leased_batches.end(), b));
    if (r->spare_batches.size() < 3) r->spare_batches.push_back(b);
    else { lk.unlock(); delete b; }
    return KV_OK;
}

extern "C" int kv_reader_num_reads(const kv_reader *r, uint64_t *n)
{
    if (!r || !n) return kv_fail_public(KV_EINVAL, "null argument");
    *n = r->num_reads;
    return KV_OK;
}
