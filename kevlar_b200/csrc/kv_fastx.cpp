// kv_fastx.cpp -- native FASTA/FASTQ(.gz) reader behind the kv_reader_* entry points.
//
// Replaces khmer.ReadParser (kevlar/count.py:40, kevlar/__init__.py:125-128) on the host side of
// the boundary: it parses straight into the batch layout the kernels take (concatenated sequence
// bytes + offsets) and keeps the header / quality text of the batch so that the few reads `kevlar
// novel` reports can be turned back into records.  Same record rules as the Python reader in
// kevlar_b200/fastx.py (which the tests keep as a cross-check): '@' starts a 4-line FASTQ record,
// '>' a FASTA record whose sequence may span lines, blank lines are skipped, CR is stripped, the
// name is the whole header line without its first character.
#include <zlib.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/kvsketch.h"

int kv_fail_public(int code, const char *fmt, ...);   // kvsketch.cu

struct kv_reader {
    gzFile fh = nullptr;
    std::string path;
    std::vector<char> buf;      // raw text window
    size_t pos = 0, end = 0;    // unread bytes are buf[pos, end)
    bool eof = false;
    uint64_t num_reads = 0;
    // a FASTA record whose end (next header or EOF) has not been seen yet
    bool fasta_open = false;
    std::string fasta_name, fasta_seq;
    // the batch handed out by the last kv_reader_next call
    std::vector<uint8_t> bases;
    std::vector<uint64_t> offsets, name_offsets, qual_offsets;
    std::vector<char> names, quals;
    std::vector<uint8_t> is_fastq;   // per record: 1 = FASTQ (has a quality string), 0 = FASTA
};

static bool reader_fill(kv_reader *r)
{
    if (r->eof) return false;
    if (r->pos > 0) {
        memmove(r->buf.data(), r->buf.data() + r->pos, r->end - r->pos);
        r->end -= r->pos;
        r->pos = 0;
    }
    if (r->buf.size() - r->end < (1u << 20)) r->buf.resize(r->buf.size() + (4u << 20));
    int got = gzread(r->fh, r->buf.data() + r->end, (unsigned)(r->buf.size() - r->end));
    if (got <= 0) { r->eof = true; return false; }
    r->end += (size_t)got;
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

static void reader_emit(kv_reader *r, const char *name, size_t nlen, const char *seq, size_t slen, const char *qual,
                        size_t qlen, bool fastq)
{
    r->bases.insert(r->bases.end(), (const uint8_t *)seq, (const uint8_t *)seq + slen);
    r->offsets.push_back(r->bases.size());
    r->names.insert(r->names.end(), name, name + nlen);
    r->name_offsets.push_back(r->names.size());
    if (fastq) r->quals.insert(r->quals.end(), qual, qual + qlen);
    r->qual_offsets.push_back(r->quals.size());
    r->is_fastq.push_back(fastq ? 1 : 0);
    r->num_reads++;
}

extern "C" int kv_reader_open(const char *path, kv_reader **out)
{
    if (!path || !out) return kv_fail_public(KV_EINVAL, "null argument");
    gzFile fh = gzopen(path, "rb");   // transparently reads plain files too
    if (!fh) return kv_fail_public(KV_EIO, "cannot open %s", path);
    gzbuffer(fh, 1u << 20);
    kv_reader *r = new kv_reader();
    r->fh = fh;
    r->path = path;
    r->buf.resize(8u << 20);
    *out = r;
    return KV_OK;
}

extern "C" int kv_reader_close(kv_reader *r)
{
    if (!r) return KV_OK;
    if (r->fh) gzclose(r->fh);
    delete r;
    return KV_OK;
}

extern "C" int kv_reader_next(kv_reader *r, uint64_t max_bases, const uint8_t **bases, const uint64_t **offsets,
                              uint64_t *n_reads, const char **names, const uint64_t **name_offsets, const char **quals,
                              const uint64_t **qual_offsets, const uint8_t **is_fastq)
{
    if (!r || !bases || !offsets || !n_reads) return kv_fail_public(KV_EINVAL, "null argument");
    r->bases.clear(); r->names.clear(); r->quals.clear();
    r->offsets.assign(1, 0); r->name_offsets.assign(1, 0); r->qual_offsets.assign(1, 0);
    r->is_fastq.clear();
    const char *line;
    size_t len;
    std::string name, seq;   // FASTQ: lines must be copied, the window may move while reading the record
    while (r->bases.size() < max_bases || r->offsets.size() == 1) {
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
