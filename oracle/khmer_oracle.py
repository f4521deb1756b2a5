"""khmer-shaped Python face of the CPU oracle (oracle/kmer_oracle.c).

TEST INFRASTRUCTURE ONLY.  Nothing under kevlar_b200/ may import this module; it is
loaded by tests/, by bench.py's cpu_baseline / --impl reference leg and by
__graft_entry__.smoke() as the checker.

It mirrors the slice of the ``khmer`` namespace that kevlar's count/novel/filter path
touches, so the reference's own Python modules
(kevlar/count.py, novel.py, filter.py, sketch.py) can be run unchanged on top of it
when generating golden vectors (tests/golden/make_golden.py).

Parity status: pinned against the reference's golden sketches, see kmer_oracle.c.
"""
import ctypes
import gzip
import os
import subprocess
import threading
from ctypes import (POINTER, byref, c_char_p, c_int, c_int64, c_uint8, c_uint32, c_uint64,
                    c_void_p, create_string_buffer)

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, 'libkmer_oracle.so')

HASH_MURMUR = 0
HASH_TWOBIT = 1


def build(force=False):
    src = os.path.join(_HERE, 'kmer_oracle.c')
    if force or not os.path.exists(_LIBPATH) or os.path.getmtime(_LIBPATH) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s', '-B', 'libkmer_oracle.so'])
    return _LIBPATH


class ko_hit(ctypes.Structure):
    _fields_ = [('read', c_uint64), ('offset', c_uint32), ('abund', c_uint8 * 12)]


HIT_DTYPE = np.dtype([('read', '<u8'), ('offset', '<u4'), ('abund', 'u1', (12,))], align=True)
assert HIT_DTYPE.itemsize == ctypes.sizeof(ko_hit)


def _load():
    lib = ctypes.CDLL(build())
    u8p = POINTER(c_uint8)
    u64p = POINTER(c_uint64)
    lib.ko_murmur3_lo.restype = c_uint64
    lib.ko_murmur3_lo.argtypes = [c_char_p, c_int, c_uint32]
    lib.ko_hash.restype = c_uint64
    lib.ko_hash.argtypes = [c_int, c_char_p, c_int, POINTER(c_int)]
    lib.ko_reverse_hash_twobit.argtypes = [c_uint64, c_int, c_char_p]
    lib.ko_primes_below.argtypes = [c_uint64, c_int, u64p]
    lib.ko_create.restype = c_void_p
    lib.ko_create.argtypes = [c_int, c_int, c_int, c_int, u64p]
    lib.ko_destroy.argtypes = [c_void_p]
    lib.ko_n_unique.restype = c_uint64
    lib.ko_n_unique.argtypes = [c_void_p]
    lib.ko_n_occupied.restype = c_uint64
    lib.ko_n_occupied.argtypes = [c_void_p]
    lib.ko_count_occupied.restype = c_uint64
    lib.ko_count_occupied.argtypes = [c_void_p]
    lib.ko_table_ptr.restype = c_void_p
    lib.ko_table_ptr.argtypes = [c_void_p, c_int]
    lib.ko_table_nbytes.restype = c_uint64
    lib.ko_table_nbytes.argtypes = [c_void_p, c_int]
    lib.ko_info.argtypes = [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int), u64p]
    lib.ko_get_hash.restype = ctypes.c_uint
    lib.ko_get_hash.argtypes = [c_void_p, c_uint64]
    lib.ko_add_hash.argtypes = [c_void_p, c_uint64]
    lib.ko_band_interval.argtypes = [c_int, c_int, u64p, u64p]
    lib.ko_consume_read.restype = c_uint64
    lib.ko_consume_read.argtypes = [c_void_p, c_char_p, c_uint64, c_int, c_int, c_void_p, c_int, c_int]
    lib.ko_consume_batch.restype = c_uint64
    lib.ko_consume_batch.argtypes = [c_void_p, c_void_p, c_void_p, c_uint64, c_int, c_int, c_void_p,
                                     c_int, c_int, c_int]
    lib.ko_abund_dist_batch.restype = None
    lib.ko_abund_dist_batch.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]
    lib.ko_novel_batch.restype = c_uint64
    lib.ko_novel_batch.argtypes = [POINTER(c_void_p), c_int, POINTER(c_void_p), c_int, c_void_p, c_void_p,
                                   c_uint64, c_int, c_int, c_int, c_int, c_int64, c_void_p, c_uint64,
                                   c_void_p, c_int]
    lib.ko_save.argtypes = [c_void_p, c_char_p]
    lib.ko_load.restype = c_void_p
    lib.ko_load.argtypes = [c_char_p, c_int, c_int, c_char_p, c_int]
    return lib


_lib = _load()

_buckets_per_byte = {'countgraph': 1, 'smallcountgraph': 2, 'nodegraph': 8}


def primes_below(x, n):
    out = (c_uint64 * n)()
    if _lib.ko_primes_below(int(x), n, out):
        raise ValueError('cannot find {} primes below {}'.format(n, x))
    return list(out)


def murmur3_lo(data, seed=0):
    return _lib.ko_murmur3_lo(data, len(data), seed)


def band_interval(num_bands, band):
    lo, hi = c_uint64(), c_uint64()
    if _lib.ko_band_interval(num_bands, band, byref(lo), byref(hi)):
        raise ValueError('Band number must be less than number of bands')
    return lo.value, hi.value


# ------------------------------------------------------------------ sequence input

class Read(object):
    __slots__ = ('name', 'sequence', 'quality')

    def __init__(self, name, sequence, quality=None):
        self.name = name
        self.sequence = sequence
        self.quality = quality


def _open_maybe_gz(filename):
    with open(filename, 'rb') as fh:
        magic = fh.read(2)
    if magic == b'\x1f\x8b':
        return gzip.open(filename, 'rt')
    return open(filename, 'r')


class ReadParser(object):
    """Stand-in for khmer.ReadParser (kevlar/count.py:40, kevlar/__init__.py:125-128):
    FASTA/FASTQ, optionally gzipped; ``name`` is the full header line."""

    def __init__(self, filename):
        self.filename = filename
        self.num_reads = 0
        self._iter = None
        self._lock = threading.Lock()

    def __iter__(self):
        with _open_maybe_gz(self.filename) as fh:
            name, chunks = None, []
            line = fh.readline()
            while line:
                line = line.rstrip('\r\n')
                if not line:
                    line = fh.readline()
                    continue
                if line[0] == '@' and name is None:
                    seq = fh.readline().rstrip('\r\n')
                    fh.readline()
                    qual = fh.readline().rstrip('\r\n')
                    self.num_reads += 1
                    yield Read(line[1:], seq, qual)
                elif line[0] == '>':
                    if name is not None:
                        self.num_reads += 1
                        yield Read(name, ''.join(chunks))
                    name, chunks = line[1:], []
                else:
                    chunks.append(line)
                line = fh.readline()
            if name is not None:
                self.num_reads += 1
                yield Read(name, ''.join(chunks))

    def shared_iter(self):
        """Iterator that several consumer threads may drain concurrently
        (kevlar/count.py:40-77 hands one parser to numthreads consumers)."""
        with self._lock:
            if self._iter is None:
                self._iter = iter(self)
        while True:
            with self._lock:
                try:
                    read = next(self._iter)
                except StopIteration:
                    return
            yield read


def reads_to_batch(seqs):
    """list of str/bytes -> (uint8 bases, uint64 offsets[n+1]) in the C-ABI batch layout."""
    bs = [s.encode('ascii') if isinstance(s, str) else s for s in seqs]
    offs = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    bases = np.frombuffer(b''.join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return bases, offs


# ------------------------------------------------------------------------ sketches

class _Sketch(object):
    _hasher = HASH_MURMUR
    _bits = 8

    def __init__(self, ksize, starting_size, n_tables, primes=None, _handle=None):
        if _handle is not None:
            self._h = _handle
        else:
            if self._hasher == HASH_TWOBIT and ksize > 32:
                raise ValueError('k-mer size must be <= 32 for graph types')
            sizes = list(primes) if primes else primes_below(int(starting_size), int(n_tables))
            arr = (c_uint64 * len(sizes))(*sizes)
            self._h = _lib.ko_create(self._hasher, self._bits, int(ksize), len(sizes), arr)
            if not self._h:
                raise MemoryError('cannot allocate sketch')
        hs, b, k, nt = c_int(), c_int(), c_int(), c_int()
        sz = (c_uint64 * 16)()
        _lib.ko_info(self._h, byref(hs), byref(b), byref(k), byref(nt), sz)
        self._ksize = k.value
        self._sizes = list(sz)[:nt.value]

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h and _lib is not None:
            _lib.ko_destroy(h)

    # -- metadata
    def ksize(self):
        return self._ksize

    def hashsizes(self):
        return list(self._sizes)

    def n_tables(self):
        return len(self._sizes)

    def n_occupied(self):
        return _lib.ko_n_occupied(self._h)

    def n_unique_kmers(self):
        return _lib.ko_n_unique(self._h)

    def table_bytes(self, t):
        n = _lib.ko_table_nbytes(self._h, t)
        return ctypes.string_at(_lib.ko_table_ptr(self._h, t), n)

    # -- hashing
    def hash(self, kmer):
        if isinstance(kmer, int):
            return kmer
        if len(kmer) != self._ksize:
            raise ValueError('k-mer length must equal the sketch k-mer size')
        ok = c_int()
        h = _lib.ko_hash(self._hasher, kmer.encode('ascii'), len(kmer), byref(ok))
        if not ok.value:
            raise ValueError('invalid DNA character in k-mer: ' + kmer)
        return h

    def reverse_hash(self, h):
        if self._hasher != HASH_TWOBIT:
            raise ValueError('not implemented for this hash function')
        buf = create_string_buffer(self._ksize + 1)
        _lib.ko_reverse_hash_twobit(h, self._ksize, buf)
        return buf.value.decode('ascii')

    def get_kmers(self, seq):
        k = self._ksize
        return [seq[i:i + k] for i in range(len(seq) - k + 1)]

    def get_kmer_hashes(self, seq):
        return [self.hash(km) for km in self.get_kmers(seq)]

    def get_kmer_counts(self, seq):
        return [self.get(km) for km in self.get_kmers(seq)]

    # -- point ops
    def get(self, kmer):
        return _lib.ko_get_hash(self._h, self.hash(kmer))

    def add(self, kmer):
        return bool(_lib.ko_add_hash(self._h, self.hash(kmer)))

    count = add

    # -- bulk ops
    def consume(self, seq):
        if len(seq) < self._ksize:
            raise ValueError('sequence length ({}) must >= the hashtable k-mer size ({})'.format(
                len(seq), self._ksize))
        return _lib.ko_consume_read(self._h, seq.encode('ascii'), len(seq), 0, 0, None, 0, 0)

    def _consume_parser(self, parser, num_bands, band, mask, threshold, consume_masked):
        if isinstance(parser, str):
            parser = ReadParser(parser)
        if num_bands:
            band_interval(num_bands, band)   # validates
        n_reads, n_kmers = 0, 0
        mh = mask._h if mask is not None else None
        for read in parser.shared_iter():
            n_reads += 1
            seq = read.sequence.encode('ascii')
            n_kmers += _lib.ko_consume_read(self._h, seq, len(seq), num_bands or 0, band or 0, mh,
                                            int(threshold), int(bool(consume_masked)))
        return n_reads, n_kmers

    def consume_seqfile(self, parser):
        return self._consume_parser(parser, 0, 0, None, 0, False)

    def consume_seqfile_banding(self, parser, num_bands, band):
        return self._consume_parser(parser, num_bands, band, None, 0, False)

    def consume_seqfile_with_mask(self, parser, mask, threshold=0, consume_masked=False):
        return self._consume_parser(parser, 0, 0, mask, threshold, consume_masked)

    def consume_seqfile_banding_with_mask(self, parser, num_bands, band, mask, threshold=0,
                                          consume_masked=False):
        return self._consume_parser(parser, num_bands, band, mask, threshold, consume_masked)

    def consume_batch(self, bases, offs, num_bands=0, band=0, mask=None, threshold=0,
                      consume_masked=False, threads=1):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        return _lib.ko_consume_batch(self._h, bases.ctypes.data, offs.ctypes.data, len(offs) - 1,
                                     num_bands or 0, band or 0, mask._h if mask is not None else None,
                                     int(threshold), int(bool(consume_masked)), threads)

    def abundance_distribution(self, parser, tracking):
        """khmer abundance_distribution (kevlar/dist.py:55): list of 65536 counts."""
        if isinstance(parser, str):
            parser = ReadParser(parser)
        dist = np.zeros(256, dtype=np.uint64)
        seqs = [read.sequence for read in parser.shared_iter()]
        if seqs:
            bases, offs = reads_to_batch(seqs)
            self.abundance_distribution_batch(bases, offs, tracking, dist)
        return [int(x) for x in dist] + [0] * (65536 - 256)

    def abundance_distribution_batch(self, bases, offs, tracking, dist):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        _lib.ko_abund_dist_batch(self._h, tracking._h, bases.ctypes.data, offs.ctypes.data, len(offs) - 1,
                                 dist.ctypes.data)
        return dist

    # -- I/O
    def save(self, filename):
        if _lib.ko_save(self._h, filename.encode()):
            raise OSError('cannot write ' + filename)

    @classmethod
    def load(cls, filename):
        err = create_string_buffer(512)
        h = _lib.ko_load(filename.encode(), cls._hasher, cls._bits, err, 512)
        if not h:
            raise OSError(err.value.decode())
        return cls(0, 0, 0, _handle=h)


class Counttable(_Sketch):
    _hasher, _bits = HASH_MURMUR, 8


class SmallCounttable(_Sketch):
    _hasher, _bits = HASH_MURMUR, 4


class Nodetable(_Sketch):
    _hasher, _bits = HASH_MURMUR, 1


class Countgraph(_Sketch):
    _hasher, _bits = HASH_TWOBIT, 8


class SmallCountgraph(_Sketch):
    _hasher, _bits = HASH_TWOBIT, 4


class Nodegraph(_Sketch):
    _hasher, _bits = HASH_TWOBIT, 1


def novel_batch(cases, ctrls, bases, offs, case_min, ctrl_max, screen=None, numbands=None,
                band_minus_1=0, threads=1, max_hits=None):
    """Batch form of the kevlar.novel.novel read loop (kevlar/novel.py:123-169).
    Returns (hits sorted by (read, offset), per-read flags)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    n_reads = len(offs) - 1
    if max_hits is None:
        max_hits = max(1024, int(len(bases)))
    hits = np.zeros(max_hits, dtype=HIT_DTYPE)
    flags = np.zeros(max(n_reads, 1), dtype=np.uint8)
    ca = (c_void_p * len(cases))(*[s._h for s in cases])
    co = (c_void_p * max(len(ctrls), 1))(*[s._h for s in ctrls])
    n = _lib.ko_novel_batch(ca, len(cases), co, len(ctrls), bases.ctypes.data, offs.ctypes.data, n_reads,
                            int(case_min), int(ctrl_max), int(screen or 0), int(numbands or 0),
                            int(band_minus_1), hits.ctypes.data, max_hits, flags.ctypes.data, threads)
    if n > max_hits:
        raise RuntimeError('hit buffer too small')
    hits = hits[:n]
    order = np.lexsort((hits['offset'], hits['read']))
    return hits[order], flags[:n_reads]


class khmer_args(object):
    @staticmethod
    def memory_setting(label):
        """khmer.khmer_args.memory_setting."""
        suffixes = {'K': 1000.0, 'M': 1000.0 ** 2, 'G': 1000.0 ** 3, 'T': 1000.0 ** 4}
        try:
            return float(label)
        except ValueError:
            prefix, suffix = label[:-1], label[-1:].upper()
            if suffix not in suffixes:
                raise ValueError('cannot parse memory setting "{}"'.format(label))
            return float(prefix) * suffixes[suffix]
