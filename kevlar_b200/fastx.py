"""FASTA/FASTQ(.gz) input for the GPU path.

Replaces ``khmer.ReadParser`` (kevlar/count.py:40, kevlar/__init__.py:125-128).  Besides the
record-at-a-time iteration the reference uses, it can hand out whole *batches* -- the
concatenated sequence bytes plus read offsets that ``kv_consume_batch`` / ``kv_novel_batch``
take (include/kvsketch.h) -- so the hot loops never touch Python strings.
"""
import gzip
import threading

import numpy as np


class Read(object):
    """One sequence record; ``name`` is the whole header line, ``quality`` is None for FASTA."""
    __slots__ = ('name', 'sequence', 'quality')

    def __init__(self, name, sequence, quality=None):
        self.name = name
        self.sequence = sequence
        self.quality = quality

    def __len__(self):
        return len(self.sequence)


class SeqBatch(object):
    """A batch in the C-ABI layout plus what is needed to rebuild records for a few reads."""
    __slots__ = ('bases', 'offsets', 'names', 'quals')

    def __init__(self, bases, offsets, names, quals):
        self.bases = bases        # np.uint8[total]
        self.offsets = offsets    # np.uint64[n+1]
        self.names = names        # list[bytes]
        self.quals = quals        # list[bytes] or None

    def __len__(self):
        return len(self.offsets) - 1

    def record(self, i):
        lo, hi = int(self.offsets[i]), int(self.offsets[i + 1])
        qual = self.quals[i].decode('ascii') if self.quals is not None and self.quals[i] is not None else None
        return Read(self.names[i].decode('ascii'), self.bases[lo:hi].tobytes().decode('ascii'), qual)


def batch_from_sequences(seqs, names=None, quals=None):
    """Build a SeqBatch from a list of str/bytes sequences."""
    bs = [s.encode('ascii') if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        np.cumsum(np.fromiter((len(b) for b in bs), dtype=np.uint64, count=len(bs)), out=offsets[1:])
    joined = b''.join(bs)
    bases = np.frombuffer(joined, dtype=np.uint8) if joined else np.zeros(0, dtype=np.uint8)
    if names is None:
        names = [b''] * len(bs)
    return SeqBatch(bases, offsets, names, quals)


def _open_binary(filename):
    fh = open(filename, 'rb')
    magic = fh.read(2)
    fh.seek(0)
    if magic == b'\x1f\x8b':
        return gzip.open(fh, 'rb')
    return fh


class FastxReader(object):
    """Streaming FASTA/FASTQ parser.  Thread-safe: several consumers may pull batches from one
    reader, as kevlar/count.py:40-77 does with one ReadParser and ``numthreads`` consumers."""

    BLOCK = 32 << 20

    def __init__(self, filename):
        self.filename = filename
        self.num_reads = 0
        self._fh = None
        self._tail = b''        # bytes after the last newline of the previous block
        self._carry = []        # lines of a FASTQ record cut by a block boundary
        self._eof = False
        self._pending = []      # parsed (name, seq, qual) tuples ...
        self._pos = 0           # ... of which [0, _pos) were already handed out
        self._pname = None      # FASTA record that may continue in the next block
        self._pchunks = []
        self._lock = threading.Lock()

    def _available(self):
        return len(self._pending) - self._pos

    def _refill(self):
        """Parse further blocks until at least one complete record is pending (or EOF)."""
        if self._fh is None:
            self._fh = _open_binary(self.filename)
        if self._pos:
            del self._pending[:self._pos]
            self._pos = 0
        while not self._pending and not self._eof:
            block = self._fh.read(self.BLOCK)
            if not block:
                self._eof = True
                self._fh.close()
                data, self._tail = self._tail, b''
            else:
                data = self._tail + block
                cut = data.rfind(b'\n')
                if cut < 0:
                    self._tail = data
                    continue
                data, self._tail = data[:cut], data[cut + 1:]
            lines = self._carry + data.split(b'\n')
            self._carry = []
            self._parse_lines(lines)
            if self._eof and self._pname is not None:
                self._pending.append((self._pname, b''.join(self._pchunks), None))
                self._pname, self._pchunks = None, []

    def _parse_lines(self, lines):
        out = self._pending
        i, n = 0, len(lines)
        while i < n:
            line = lines[i]
            if not line or line == b'\r':
                i += 1
                continue
            c = line[0:1]
            if c == b'@' and self._pname is None:
                if i + 3 >= n and not self._eof:       # record continues in the next block
                    self._carry = lines[i:]
                    return
                seq = lines[i + 1].rstrip(b'\r') if i + 1 < n else b''
                qual = lines[i + 3].rstrip(b'\r') if i + 3 < n else b''
                out.append((line[1:].rstrip(b'\r'), seq, qual))
                i += 4
            elif c == b'>':
                if self._pname is not None:
                    out.append((self._pname, b''.join(self._pchunks), None))
                self._pname, self._pchunks = line[1:].rstrip(b'\r'), []
                i += 1
            else:
                if self._pname is not None:
                    self._pchunks.append(line.rstrip(b'\r'))
                i += 1

    # -- record iteration (khmer.ReadParser protocol)
    def __iter__(self):
        while True:
            with self._lock:
                if not self._available():
                    self._refill()
                if not self._available():
                    return
                name, seq, qual = self._pending[self._pos]
                self._pos += 1
                self.num_reads += 1
            yield Read(name.decode('ascii'), seq.decode('ascii'), qual.decode('ascii') if qual is not None else None)

    # -- batch iteration (GPU path)
    def batches(self, max_bases=64 << 20, keep_text=False):
        """Yield SeqBatch objects of at most ~max_bases bases, in file order."""
        while True:
            with self._lock:
                if not self._available():
                    self._refill()
                if not self._available():
                    return
                pend, lo = self._pending, self._pos
                hi, total = lo, 0
                while hi < len(pend) and (hi == lo or total + len(pend[hi][1]) <= max_bases):
                    total += len(pend[hi][1])
                    hi += 1
                recs = pend[lo:hi]
                self._pos = hi
                self.num_reads += hi - lo
            seqs = [r[1] for r in recs]
            names = [r[0] for r in recs] if keep_text else None
            quals = [r[2] for r in recs] if keep_text else None
            yield batch_from_sequences(seqs, names, quals)


ReadParser = FastxReader
