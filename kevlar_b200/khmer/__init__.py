"""Drop-in for the slice of the ``khmer`` namespace that kevlar's count -> novel -> filter
path uses, backed by libkvsketch.so on a B200.

Six sketch classes (MurmurHash3 "table" and 2-bit "graph" hashing x 8/4/1-bit counters),
``ReadParser``, ``khmer_args.memory_setting`` and ``_buckets_per_byte`` keep khmer's names,
argument meaning and exception types.  On top of the khmer methods every sketch has batch
entry points (``consume_batch``, ``get_many``, ``add_many``, ``hash_many``) which are what the
GPU-side rewrite of kevlar's loops calls.

All arithmetic happens in the CUDA library; nothing here computes a hash or touches a
counter on the CPU.
"""
from ctypes import byref, c_int, c_uint64, c_void_p

import numpy as np

from kevlar_b200 import _lib
from kevlar_b200._lib import HASH_MURMUR, HASH_TWOBIT, MEM_DEVICE, MEM_HOST, check, lib
from kevlar_b200.fastx import FastxReader, NativeFastxReader, Read, SeqBatch, batch_from_sequences  # noqa: F401
from kevlar_b200.khmer import khmer_args  # noqa: F401

ReadParser = NativeFastxReader

_buckets_per_byte = {'countgraph': 1, 'smallcountgraph': 2, 'nodegraph': 8}

_TWOBIT_LETTERS = 'ATCG'   # khmer's 2-bit code order

BATCH_BASES = 64 << 20     # bases per consume/novel batch handed to the GPU


def calc_expected_collisions(sketch, force=False, max_false_pos=0.2):
    """khmer.calc_expected_collisions: (n_occupied / min table size) ** n_tables."""
    sizes = sketch.hashsizes()
    fp_all = (float(sketch.n_occupied()) / min(sizes)) ** len(sizes)
    if fp_all > max_false_pos and not force:
        raise SystemExit('ERROR: the graph structure is too small for this data set')
    return fp_all


class _Sketch(object):
    _hasher = HASH_MURMUR
    _bits = 8

    def __init__(self, ksize, starting_size, n_tables, primes=None, device=None, _handle=None):
        self._h = None
        if _handle is None:
            ksize = int(ksize)
            sizes = [int(p) for p in primes] if primes else _lib.primes_below(int(starting_size), int(n_tables))
            arr = (c_uint64 * len(sizes))(*sizes)
            handle = c_void_p()
            dev = _lib.current_device() if device is None else int(device)
            check(lib().kv_sketch_create(self._hasher, self._bits, ksize, len(sizes), arr, dev, byref(handle)))
            _handle = handle
        self._h = _handle
        hs, bits, k, nt, dev = c_int(), c_int(), c_int(), c_int(), c_int()
        sz = (c_uint64 * _lib.MAX_TABLES)()
        check(lib().kv_sketch_info(self._h, byref(hs), byref(bits), byref(k), byref(nt), sz, byref(dev)))
        self._ksize, self._sizes, self._device = k.value, list(sz)[:nt.value], dev.value

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and _lib._lib is not None:
            if getattr(self, '_p2p_peers', None):   # peer tables mapped by a multi-GPU merge: unmap before freeing
                try:
                    from kevlar_b200 import multigpu
                    multigpu.close_p2p(self)
                except Exception:
                    pass
            self._h = None
            _lib._lib.kv_sketch_destroy(h)

    # ---------------------------------------------------------------- metadata
    def ksize(self):
        return self._ksize

    def hashsizes(self):
        return list(self._sizes)

    def n_tables(self):
        return len(self._sizes)

    @property
    def device(self):
        return self._device

    def n_occupied(self):
        n = c_uint64()
        check(lib().kv_sketch_stats(self._h, byref(n), None, None))
        return n.value

    def n_unique_kmers(self):
        n, valid = c_uint64(), c_int()
        check(lib().kv_sketch_stats(self._h, None, byref(n), byref(valid)))
        if not valid.value:
            raise RuntimeError('n_unique_kmers is not available: exact tracking was switched off '
                               '(or the sketch was merged across GPUs)')
        return n.value

    def clear(self):
        """Reset to an empty sketch in place (no reallocation)."""
        check(lib().kv_sketch_clear(self._h))

    def set_unique_tracking(self, on):
        """True / False: the exact n_unique_kmers bookkeeping of every consume.  'deferred': not tracked, but a
        consume that fits one chunk prepares the first-touch passes that `multigpu.count_sharded` runs once the
        other ranks' occupancy is known (kv_unique_last_batch)."""
        check(lib().kv_sketch_set_unique_tracking(self._h, 2 if on == 'deferred' else int(bool(on))))

    def table_bytes(self, t):
        """Raw khmer-layout bytes of table t (D2H copy)."""
        ptr, nbytes = c_void_p(), c_uint64()
        check(lib().kv_sketch_table(self._h, t, byref(ptr), byref(nbytes)))
        out = np.empty(nbytes.value, dtype=np.uint8)
        check(lib().kv_sketch_read_table(self._h, t, out.ctypes.data, nbytes.value))
        return out.tobytes()

    def flat_device_buffer(self):
        ptr, nbytes = c_void_p(), c_uint64()
        check(lib().kv_sketch_flat(self._h, byref(ptr), byref(nbytes)))
        return ptr.value, nbytes.value

    # ----------------------------------------------------------------- hashing
    def _check_kmer(self, kmer):
        if len(kmer) != self._ksize:
            raise ValueError('k-mer length {} does not match the sketch k-mer size {}'.format(len(kmer), self._ksize))

    def hash_many(self, kmers):
        """hash() for a list of k-mer strings -> np.uint64 array (one GPU call)."""
        if not len(kmers):
            return np.zeros(0, dtype=np.uint64)
        for km in kmers:
            self._check_kmer(km)
        buf = np.frombuffer(''.join(kmers).encode('ascii'), dtype=np.uint8)
        out = np.empty(len(kmers), dtype=np.uint64)
        ok = np.empty(len(kmers), dtype=np.uint8)
        check(lib().kv_hash_kmers(self._hasher, self._ksize, buf.ctypes.data, len(kmers), self._device,
                                  out.ctypes.data, ok.ctypes.data))
        if not ok.all():
            bad = kmers[int(np.argmin(ok))]
            raise ValueError('invalid DNA character in k-mer: ' + bad)
        return out

    def hash(self, kmer):
        if isinstance(kmer, (int, np.integer)):
            return int(kmer)
        return int(self.hash_many([kmer])[0])

    def reverse_hash(self, khash):
        if self._hasher != HASH_TWOBIT:
            raise ValueError('reverse_hash is not implemented for this hash function (MurmurHash is one-way)')
        khash = int(khash)
        return ''.join(_TWOBIT_LETTERS[(khash >> (2 * i)) & 3] for i in range(self._ksize - 1, -1, -1))

    def get_kmers(self, sequence):
        k = self._ksize
        return [sequence[i:i + k] for i in range(len(sequence) - k + 1)]

    def _per_position(self, sequence, want_hashes, want_counts):
        if len(sequence) < self._ksize:
            raise ValueError('sequence length ({}) must >= the hashtable k-mer size ({})'.format(
                len(sequence), self._ksize))
        batch = batch_from_sequences([sequence])
        n = len(batch.bases)
        hashes = np.empty(n, dtype=np.uint64) if want_hashes else None
        counts = np.empty(n, dtype=np.uint8) if want_counts else None
        valid = np.empty(n, dtype=np.uint8)
        check(lib().kv_kmer_counts_batch(self._h, batch.bases.ctypes.data, batch.offsets.ctypes.data, 1, MEM_HOST,
                                         hashes.ctypes.data if want_hashes else None,
                                         counts.ctypes.data if want_counts else None, valid.ctypes.data))
        nk = n - self._ksize + 1
        if not valid[:nk].all():
            raise ValueError('invalid DNA character in sequence')
        return (hashes[:nk] if want_hashes else None), (counts[:nk] if want_counts else None)

    def get_kmer_hashes(self, sequence):
        return [int(h) for h in self._per_position(sequence, True, False)[0]]

    def get_kmer_counts(self, sequence):
        return [int(c) for c in self._per_position(sequence, False, True)[1]]

    def get_kmer_counts_many(self, sequences):
        """get_kmer_counts for a list of sequences in ONE GPU call (e.g. all call windows of a
        `simlike` run); returns one uint8 array per sequence.  Sequences shorter than k give an
        empty array."""
        batch = batch_from_sequences(sequences)
        n, total = len(sequences), len(batch.bases)
        if total == 0:
            return [np.zeros(0, dtype=np.uint8) for _ in sequences]
        counts = np.empty(total, dtype=np.uint8)
        valid = np.empty(total, dtype=np.uint8)
        check(lib().kv_kmer_counts_batch(self._h, batch.bases.ctypes.data, batch.offsets.ctypes.data, n, MEM_HOST, None,
                                         counts.ctypes.data, valid.ctypes.data))
        out = []
        for i in range(n):
            lo, hi = int(batch.offsets[i]), int(batch.offsets[i + 1])
            nk = max(0, hi - lo - self._ksize + 1)
            if nk and not valid[lo:lo + nk].all():
                raise ValueError('invalid DNA character in sequence {}'.format(i))
            out.append(counts[lo:lo + nk])
        return out

    # --------------------------------------------------------------- point ops
    def _to_hashes(self, items):
        if isinstance(items, np.ndarray) and items.dtype == np.uint64:
            return np.ascontiguousarray(items)
        if len(items) and isinstance(items[0], str):
            return self.hash_many(list(items))
        return np.asarray(list(items), dtype=np.uint64)

    def get_many(self, items):
        """get() for a list of k-mer strings or an array of hashes -> np.uint8 counts."""
        hashes = self._to_hashes(items)
        out = np.empty(len(hashes), dtype=np.uint8)
        if len(hashes):
            check(lib().kv_get_hashes(self._h, hashes.ctypes.data, len(hashes), out.ctypes.data))
        return out

    def add_many(self, items):
        """add()/count() for a list of k-mer strings or hashes, applied in order."""
        hashes = self._to_hashes(items)
        if len(hashes):
            check(lib().kv_add_hashes(self._h, hashes.ctypes.data, len(hashes)))

    def get(self, kmer):
        return int(self.get_many([kmer] if isinstance(kmer, str) else np.array([kmer], dtype=np.uint64))[0])

    def add(self, kmer):
        self.add_many([kmer] if isinstance(kmer, str) else np.array([kmer], dtype=np.uint64))

    count = add

    # ---------------------------------------------------------------- bulk ops
    def consume_batch(self, bases, offsets, num_bands=None, band=None, mask=None, threshold=0,
                      consume_masked=False, where=MEM_HOST, wait=True):
        """One kv_consume_batch call.  ``bases``/``offsets`` are numpy arrays (host) or raw
        device pointers (where=MEM_DEVICE, offsets = (ptr, n_reads)).  Returns the number of
        k-mers counted when ``wait`` is true."""
        if where == MEM_HOST:
            bases, offsets = _lib.as_u8(bases), _lib.as_u64(offsets)
            bptr, optr, n_reads = bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1
        else:
            bptr, (optr, n_reads) = bases, offsets
        n = c_uint64()
        check(lib().kv_consume_batch(self._h, bptr, optr, n_reads, where, int(num_bands or 0), int(band or 0),
                                     mask._h if mask is not None else None, int(threshold),
                                     int(bool(consume_masked)), byref(n) if wait else None))
        return n.value if wait else None

    def consume(self, sequence):
        if len(sequence) < self._ksize:
            raise ValueError('sequence length ({}) must >= the hashtable k-mer size ({})'.format(
                len(sequence), self._ksize))
        if set(sequence) - set('ACGT'):
            raise ValueError('invalid DNA character in sequence')
        batch = batch_from_sequences([sequence])
        return self.consume_batch(batch.bases, batch.offsets)

    def _consume_parser(self, parser, num_bands, band, mask, threshold, consume_masked):
        if isinstance(parser, str):
            parser = ReadParser(parser)
        if mask is not None and not isinstance(mask, _Sketch):
            raise TypeError('mask must be a sketch')
        n_reads, n_kmers = 0, 0
        for batch in parser.batches(BATCH_BASES):
            n_reads += len(batch)
            n_kmers += self.consume_batch(batch.bases, batch.offsets, num_bands, band, mask, threshold,
                                          consume_masked)
        return n_reads, n_kmers

    def consume_seqfile(self, parser):
        return self._consume_parser(parser, None, None, None, 0, False)

    def consume_seqfile_banding(self, parser, num_bands, band):
        if band is None or band < 0 or band >= num_bands:
            raise ValueError('Band number must be less than number of bands')
        return self._consume_parser(parser, num_bands, band, None, 0, False)

    def consume_seqfile_with_mask(self, parser, mask, threshold=0, consume_masked=False):
        return self._consume_parser(parser, None, None, mask, threshold, consume_masked)

    def consume_seqfile_banding_with_mask(self, parser, num_bands, band, mask, threshold=0,
                                          consume_masked=False):
        if band is None or band < 0 or band >= num_bands:
            raise ValueError('Band number must be less than number of bands')
        return self._consume_parser(parser, num_bands, band, mask, threshold, consume_masked)

    def abundance_distribution_batch(self, bases, offsets, tracking, where=MEM_HOST):
        """One kv_abund_dist_batch call: numpy uint64[256] histogram of ``self.get(kmer)`` over the
        k-mers of the batch that ``tracking`` had not seen yet (``tracking`` is updated)."""
        if not isinstance(tracking, _Sketch):
            raise TypeError('tracking must be a sketch')
        if where == MEM_HOST:
            bases, offsets = _lib.as_u8(bases), _lib.as_u64(offsets)
            bptr, optr, n_reads = bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1
        else:
            bptr, (optr, n_reads) = bases, offsets
        dist = np.zeros(256, dtype=np.uint64)
        check(lib().kv_abund_dist_batch(self._h, tracking._h, bptr, optr, n_reads, where, dist.ctypes.data))
        return dist

    def abundance_distribution(self, parser, tracking):
        """khmer ``abundance_distribution(parser, tracking)`` (kevlar/dist.py:55): list of 65536
        counts, entry i = number of distinct (first-seen in ``tracking``) k-mers with abundance i."""
        if isinstance(parser, str):
            parser = ReadParser(parser)
        total = np.zeros(256, dtype=np.uint64)
        for batch in parser.batches(BATCH_BASES):
            total += self.abundance_distribution_batch(batch.bases, batch.offsets, tracking)
        return [int(x) for x in total] + [0] * (65536 - 256)

    # --------------------------------------------------------------------- I/O
    def save(self, filename):
        check(lib().kv_sketch_save(self._h, str(filename).encode()))

    @classmethod
    def load(cls, filename, device=None):
        handle = c_void_p()
        dev = _lib.current_device() if device is None else int(device)
        check(lib().kv_sketch_load(str(filename).encode(), cls._hasher, cls._bits, dev, byref(handle)))
        return cls(0, 0, 0, _handle=handle)


class Counttable(_Sketch):
    """MurmurHash3, 8-bit saturating counters (what `kevlar count` builds by default)."""
    _hasher, _bits = HASH_MURMUR, 8


class SmallCounttable(_Sketch):
    _hasher, _bits = HASH_MURMUR, 4


class Nodetable(_Sketch):
    _hasher, _bits = HASH_MURMUR, 1


class Countgraph(_Sketch):
    """2-bit canonical hashing (k <= 32), 8-bit saturating counters."""
    _hasher, _bits = HASH_TWOBIT, 8


class SmallCountgraph(_Sketch):
    _hasher, _bits = HASH_TWOBIT, 4


class Nodegraph(_Sketch):
    _hasher, _bits = HASH_TWOBIT, 1


def novel_batch(cases, ctrls, bases, offsets, case_min, ctrl_max, screen=None, num_bands=None,
                band_minus_1=0, where=MEM_HOST, max_hits=None):
    """One kv_novel_batch call: every k-mer of every read of the batch against all case and
    control sketches.  Returns (hits, read_flags, discard_pos): ``hits`` is a structured array
    (read, offset, abund[16]) sorted by (read, offset); see include/kvsketch.h for the flags."""
    if where == MEM_HOST:
        bases, offsets = _lib.as_u8(bases), _lib.as_u64(offsets)
        bptr, optr, n_reads = bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1
        total = len(bases)
    else:
        bptr, (optr, n_reads, total) = bases, offsets
    ca = (c_void_p * len(cases))(*[s._h for s in cases])
    co = (c_void_p * max(1, len(ctrls)))(*[s._h for s in ctrls])
    flags = np.zeros(max(1, n_reads), dtype=np.uint8)
    discard = np.full(max(1, n_reads), 0xffffffff, dtype=np.uint32)
    if max_hits is None:
        max_hits = max(4096, total // 64)
    while True:
        hits = np.empty(max_hits, dtype=_lib.HIT_DTYPE)
        n = c_uint64()
        rc = lib().kv_novel_batch(ca, len(cases), co, len(ctrls), bptr, optr, n_reads, where, int(case_min),
                                  int(ctrl_max), int(screen or 0), int(num_bands or 0), int(band_minus_1),
                                  hits.ctypes.data, max_hits, byref(n), flags.ctypes.data,
                                  discard.ctypes.data if screen else None)
        if rc == _lib.KV_EOVERFLOW:
            max_hits = int(n.value) + 1024
            continue
        check(rc)
        return hits[:n.value], flags[:n_reads], discard[:n_reads]
