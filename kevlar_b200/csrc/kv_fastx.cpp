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
// Two stages, so that inflate (the slow part of a .gz input) overlaps both the parsing and the
// GPU work of the caller:
//   producer thread   read() for plain files, zlib for gzip (detected by magic number); fills
//                     4 MB blocks of raw text into a small queue, always a few blocks ahead;
//   kv_reader_next    appends blocks to its text window and splits records.  Complete 4-line FASTQ
//                     records are cut with four memchr calls and copied once, straight into the
//                     batch arrays; everything else (FASTA, the ragged end of the file) goes
//                     through the line-by-line path.
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
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

}   // namespace

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
    std::vector<uint8_t> bases;
    std::vector<uint64_t> offsets, name_offsets, qual_offsets;
    std::vector<char> names, quals;
    std::vector<uint8_t> is_fastq;   // per record: 1 = FASTQ (has a quality string), 0 = FASTA
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
    r->bases.insert(r->bases.end(), (const uint8_t *)seq, (const uint8_t *)seq + slen);
    r->offsets.push_back(r->bases.size());
    if (r->keep_text) {
        r->names.insert(r->names.end(), name, name + nlen);
        if (fastq) r->quals.insert(r->quals.end(), qual, qual + qlen);
    }
    r->name_offsets.push_back(r->names.size());
    r->qual_offsets.push_back(r->quals.size());
    r->is_fastq.push_back(fastq ? 1 : 0);
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

extern "C" int kv_reader_open(const char *path, kv_reader **out)
{
    if (!path || !out) return kv_fail_public(KV_EINVAL, "null argument");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return kv_fail_public(KV_EIO, "cannot open %s", path);
    unsigned char magic[2] = {0, 0};
    ssize_t got = read(fd, magic, 2);
    lseek(fd, 0, SEEK_SET);
    kv_reader *r = new kv_reader();
    r->path = path;
    if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
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
    for (Block *b : r->full) delete b;
    for (Block *b : r->spare) delete b;
    if (r->gz) gzclose(r->gz);
    if (r->fd >= 0) close(r->fd);
    delete r;
    return KV_OK;
}

extern "C" int kv_reader_next(kv_reader *r, uint64_t max_bases, const uint8_t **bases, const uint64_t **offsets,
                              uint64_t *n_reads, const char **names, const uint64_t **name_offsets, const char **quals,
                              const uint64_t **qual_offsets, const uint8_t **is_fastq)
{
    if (!r || !bases || !offsets || !n_reads) return kv_fail_public(KV_EINVAL, "null argument");
    r->keep_text = names != nullptr || quals != nullptr;   // sequences only: skip the header / quality copies
    r->bases.clear(); r->names.clear(); r->quals.clear();
    r->offsets.assign(1, 0); r->name_offsets.assign(1, 0); r->qual_offsets.assign(1, 0);
    r->is_fastq.clear();
    if (r->bases.capacity() < max_bases && max_bases <= (1ull << 31)) r->bases.reserve((size_t)max_bases + (1u << 16));
    const char *line;
    size_t len;
    std::string name, seq;   // slow path: lines are copied, the window may move while reading the record
    while (r->bases.size() < max_bases || r->offsets.size() == 1) {
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
    *bases = r->bases.data();
    *offsets = r->offsets.data();
    *n_reads = r->offsets.size() - 1;
    if (names) *names = r->names.data();
    if (name_offsets) *name_offsets = r->name_offsets.data();
    if (quals) *quals = r->quals.data();
    if (qual_offsets) *qual_offsets = r->qual_offsets.data();
    if (is_fastq) *is_fastq = r->is_fastq.data();
    return KV_OK;
}

extern "C" int kv_reader_num_reads(const kv_reader *r, uint64_t *n)
{
    if (!r || !n) return kv_fail_public(KV_EINVAL, "null argument");
    *n = r->num_reads;
    return KV_OK;
}
