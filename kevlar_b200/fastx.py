"""FASTA/FASTQ(.gz) input for the GPU path.

Replaces ``khmer.ReadParser`` (kevlar/count.py:40, kevlar/__init__.py:125-128).  Besides the
record-at-a-time iteration the reference uses, it can hand out whole *batches* -- the
concatenated sequence bytes plus read offsets that ``kv_consume_batch`` / ``kv_novel_batch``
take (include/kvsketch.h) -- so the hot loops never touch Python strings.
"""
import ctypes
import gzip
import os
import queue
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
    """A batch in the C-ABI layout plus what is needed to rebuild records for a few reads.
    Header and quality text is kept either as Python lists or as the packed blobs the native
    reader returns (blob + n+1 offsets), and only unpacked for the reads somebody asks for."""
    __slots__ = ('bases', 'offsets', '_names', '_quals', '_name_blob', '_name_offs', '_qual_blob', '_qual_offs',
                 '_is_fastq')

    def __init__(self, bases, offsets, names=None, quals=None, packed=None):
        self.bases = bases        # np.uint8[total]
        self.offsets = offsets    # np.uint64[n+1]
        self._names = names       # list[bytes] or None
        self._quals = quals       # list[bytes or None] or None
        self._name_blob = self._name_offs = self._qual_blob = self._qual_offs = self._is_fastq = None
        if packed is not None:
            self._name_blob, self._name_offs, self._qual_blob, self._qual_offs, self._is_fastq = packed

    def __len__(self):
        return len(self.offsets) - 1

    def name(self, i):
        if self._names is not None:
            return self._names[i]
        if self._name_blob is None:
            return b''
        return self._name_blob[int(self._name_offs[i]):int(self._name_offs[i + 1])]

    def qual(self, i):
        if self._quals is not None:
            return self._quals[i]
        if self._qual_blob is None or not self._is_fastq[i]:
            return None
        return self._qual_blob[int(self._qual_offs[i]):int(self._qual_offs[i + 1])]

    @property
    def names(self):
        if self._names is None:
            self._names = [self.name(i) for i in range(len(self))]
        return self._names

    def find_name(self, key):
        """Index of the first read called `key` (bytes), or -1."""
        if self._names is None and self._name_blob is not None:
            at = self._name_blob.find(key)
            while at >= 0:   # a hit must span exactly one name
                i = int(np.searchsorted(self._name_offs, at, side='right')) - 1
                if int(self._name_offs[i]) == at and int(self._name_offs[i + 1]) == at + len(key):
                    return i
                at = self._name_blob.find(key, at + 1)
            return -1
        try:
            return self.names.index(key)
        except ValueError:
            return -1

    def tail(self, start):
        """The batch without its first `start` reads."""
        offsets = np.ascontiguousarray(self.offsets[start:] - self.offsets[start])
        bases = self.bases[int(self.offsets[start]):]
        n = len(self)
        return SeqBatch(bases, offsets, [self.name(i) for i in range(start, n)], [self.qual(i) for i in range(start, n)])

    def record(self, i):
        lo, hi = int(self.offsets[i]), int(self.offsets[i + 1])
        qual = self.qual(i)
        return Read(self.name(i).decode('ascii'), self.bases[lo:hi].tobytes().decode('ascii'),
                    qual.decode('ascii') if qual is not None else None)


def batch_from_sequences(seqs, names=None, quals=None):
    """Build a SeqBatch from a list of str/bytes sequences."""
    bs = [s.encode('ascii') if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        np.cumsum(np.fromiter((len(b) for b in bs), dtype=np.uint64, count=len(bs)), out=offsets[1:])
    joined = b''.join(bs)
    bases = np.frombuffer(joined, dtype=np.uint8) if joined else np.zeros(0, dtype=np.uint8)
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



class _BatchLease(object):
    """Keeps a native batch (kv_batch) on loan while numpy views of its arrays are alive."""

    def __init__(self, lib, handle):
        self._lib, self._handle = lib, handle

    def view(self, address, count, typestr):
        if not count:
            return np.empty(0, dtype=np.dtype(typestr))

        class _Memory(object):
            pass
        memory = _Memory()
        memory.lease = self
        memory.__array_interface__ = {'shape': (int(count),), 'typestr': typestr, 'data': (int(address), True), 'version': 3}
        return np.asarray(memory)

    def __del__(self):
        handle, self._handle = getattr(self, '_handle', None), None
        if handle is not None and handle.value and self._lib._lib is not None:
            self._lib._lib.kv_batch_release(handle)


class NativeFastxReader(object):
    """The same interface on top of the native parser in libkvsketch.so (kv_reader_*: a read-ahead
    thread doing the file I/O / inflate, memchr record splitting straight into the batch arrays);
    the pure-Python reader above remains as the implementation the tests compare it with."""

    def __init__(self, filename):
        from kevlar_b200 import _lib
        self._lib = _lib
        self.filename = filename
        self.num_reads = 0
        self._lock = threading.Lock()
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().kv_reader_open(str(filename).encode(), ctypes.byref(self._h)))
        self._records = None   # record-at-a-time iteration state: (batch, next index)
        self._stash = []       # batches parsed ahead by an abandoned batches() generator: the next consumer gets them first

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h is not None and h.value and self._lib._lib is not None:
            self._lib._lib.kv_reader_close(h)

    def _next(self, max_bases, keep_text):
        """One kv_reader_next_batch call.  The batch's sequence bytes and offsets are numpy VIEWS of the native
        batch (no copy): the batch stays on loan until the last view is garbage-collected, then goes back to the
        reader for reuse.  Header / quality text (novel path only) is copied into bytes objects."""
        if self._stash:
            return self._stash.pop(0)
        c = ctypes
        handle = c.c_void_p()
        self._lib.check(self._lib.lib().kv_reader_next_batch(self._h, int(max_bases), int(bool(keep_text)), c.byref(handle)))
        if not handle.value:
            return None
        lease = _BatchLease(self._lib, handle)
        bases, offs, names, noffs, quals, qoffs, isfq = (c.c_void_p() for _ in range(7))
        n = c.c_uint64()
        text = (c.byref(names), c.byref(noffs), c.byref(quals), c.byref(qoffs), c.byref(isfq)) if keep_text else (None,) * 5
        self._lib.check(self._lib.lib().kv_batch_arrays(handle, c.byref(bases), c.byref(offs), c.byref(n), *text))
        n = n.value
        offsets = lease.view(offs.value, n + 1, '<u8')
        total = int(offsets[-1])
        b = lease.view(bases.value, total, '|u1')
        packed = None
        if keep_text:
            no = lease.view(noffs.value, n + 1, '<u8').copy()
            qo = lease.view(qoffs.value, n + 1, '<u8').copy()
            fq = lease.view(isfq.value, n, '|u1').copy()
            packed = (c.string_at(names, int(no[-1])), no, c.string_at(quals, int(qo[-1])), qo, fq)
        self.num_reads += n
        return SeqBatch(b, offsets, packed=packed)

    def batches(self, max_bases=64 << 20, keep_text=False, prefetch=True):
        """Batches in file order.  With ``prefetch`` a helper thread parses one batch ahead, so the
        parsing of batch i+1 overlaps whatever the consumer does with batch i (typically waiting
        for the GPU inside a ctypes call, which releases the GIL).  A generator that is abandoned
        early hands the batch it had parsed ahead to whoever iterates the parser next."""
        if not prefetch or os.environ.get('KV_NO_PREFETCH'):
            while True:
                with self._lock:
                    batch = self._next(max_bases, keep_text)
                if batch is None:
                    return
                yield batch
        ahead = queue.Queue(maxsize=1)
        stop = threading.Event()

        def parse_ahead():
            try:
                while not stop.is_set():
                    with self._lock:
                        batch = self._next(max_bases, keep_text)
                    ahead.put(batch)
                    if batch is None:
                        return
            except BaseException as exc:   # handed to the consumer
                ahead.put(exc)

        worker = threading.Thread(target=parse_ahead, name='kv-fastx-prefetch', daemon=True)
        worker.start()
        try:
            while True:
                item = ahead.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            leftover = []
            while worker.is_alive() or not ahead.empty():   # unblock a pending put, then let the thread see the flag
                try:
                    item = ahead.get(timeout=0.05)
                except queue.Empty:
                    continue
                if isinstance(item, SeqBatch):
                    leftover.append(item)   # parsed and counted but never yielded: keep it for the next consumer
            with self._lock:
                self._stash = leftover + self._stash

    def __iter__(self):
        while True:
            with self._lock:
                if self._records is None or self._records[1] >= len(self._records[0]):
                    batch = self._next(4 << 20, True)
                    if batch is None:
                        return
                    self._records = [batch, 0]
                batch, i = self._records
                self._records[1] = i + 1
            yield batch.record(i)


ReadParser = NativeFastxReader
