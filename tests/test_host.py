"""Host-side logic and the C-ABI surface -- runs without a GPU (`pytest -m "not gpu"`)."""
import ctypes
import io
import os
import re
import socket
import subprocess
import sys
import threading

import numpy as np
import pytest

from conftest import REPO, golden_data

import kevlar_b200 as kv
from kevlar_b200 import _lib, fastx, multigpu
from kevlar_b200.sequence import KmerOfInterest, Record


# ------------------------------------------------------------------ C ABI surface

def _declared_symbols():
    text = open(os.path.join(REPO, 'include', 'kvsketch.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(kv_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    declared = _declared_symbols()
    assert len(declared) >= 30
    handle = ctypes.CDLL(_lib.LIBPATH)
    for name in declared:
        assert hasattr(handle, name), 'libkvsketch.so does not export ' + name
    bound = sorted(name for name, _, _ in _lib.SYMBOLS)
    assert bound == declared, 'ctypes binding and header disagree'
    assert _lib.lib().kv_abi_version() == 1


def test_no_cpu_fallback():
    """Without a CUDA device every computing entry point must fail loudly."""
    if _lib.device_count() > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(_lib.KvError, match='no usable CUDA device'):
        kv.khmer.Counttable(21, 1e4, 4)
    with pytest.raises(_lib.KvError, match='no CPU fallback'):
        kv.sketch.load(golden_data('test.counttable'))
    with pytest.raises(_lib.KvError):
        kv.count.load_sample_seqfile([golden_data('bogus-genome/refr.fa')], 21, 1e6)


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, 'kevlar_b200')):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(root, fn)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, fn
                assert 'kmer_oracle' not in text or fn.endswith(('.cuh', '.cu')), fn


def test_primes_match_oracle(oracle):
    for x, n in [(2500, 4), (250000, 4), (125, 4), (1e4, 4), (100, 4), (16000000, 4), (7, 2)]:
        assert _lib.primes_below(x, n) == oracle.primes_below(x, n)
    with pytest.raises(ValueError):
        _lib.primes_below(2, 4)


def test_memory_setting():
    ms = kv.khmer.khmer_args.memory_setting
    assert ms('10K') == 1e4 and ms('1M') == 1e6 and ms('8G') == 8e9 and ms('2T') == 2e12
    assert ms('1e7') == 1e7 and ms('97') == 97.0 and ms('500k') == 5e5 and ms('1.5M') == 1.5e6
    with pytest.raises(ValueError):
        ms('lots')


# ------------------------------------------------------------------ CLI

def test_cli_defaults():
    """kevlar/tests/test_novel.py:40-56, kevlar/cli/count.py, kevlar/cli/filter.py defaults."""
    args = kv.cli.parser().parse_args(['novel', '--case', 'case1.fq', '--control', 'cntl1.fq', '--control',
                                       'cntl2.fq', '-k', '17'])
    assert (args.ksize, args.case_min, args.ctrl_max, args.num_bands, args.band) == (17, 6, 1, None, None)
    assert args.case == [['case1.fq']] and args.control == [['cntl1.fq'], ['cntl2.fq']]
    assert args.memory == 1e6 and args.max_fpr == 0.2 and args.threads == 1 and args.abund_screen is None
    args = kv.cli.parser().parse_args(['novel', '--num-bands', '8', '--band', '1', '--case', 'c.fq', '--control', 'd.fq'])
    assert (args.ksize, args.num_bands, args.band) == (31, 8, 1)
    args = kv.cli.parser().parse_args(['count', 'out.ct', 'a.fq', 'b.fq'])
    assert (args.ksize, args.counter_size, args.memory, args.max_fpr, args.mask, args.count_masked) == \
        (31, 8, 1e6, 0.2, None, False)
    assert args.seqfile == ['a.fq', 'b.fq'] and args.threads == 1
    args = kv.cli.parser().parse_args(['filter', '--mask', 'm.nt', 'reads.augfastq'])
    assert (args.memory, args.max_fpr, args.ctrl_max, args.case_min) == (1e6, 0.01, 1, 6)
    with pytest.raises(SystemExit):
        kv.cli.parser().parse_args(['count', '-c', '2', 'out', 'in.fq'])


def test_band_argument_errors():
    """kevlar/tests/test_novel.py:25-37,58-65; kevlar/tests/test_count.py:83-91."""
    with pytest.raises(ValueError, match=r'Must specify `numbands` and `band` together'):
        list(kv.novel.novel(None, [], [], numbands=4))
    with pytest.raises(ValueError, match=r'Must specify `numbands` and `band` together'):
        list(kv.novel.novel(None, [], [], band=0))
    with pytest.raises(ValueError, match=r'`band` must be a value between 0 and 3'):
        list(kv.novel.novel(None, [], [], numbands=4, band=-1))
    args = kv.cli.parser().parse_args(['novel', '--case', 'c.fq', '--control', 'd.fq', '--band', '1'])
    with pytest.raises(ValueError, match=r'Must specify --num-bands and --band together'):
        kv.novel.main(args)
    args = kv.cli.parser().parse_args(['count', '--band', '2', 'out', 'in.fq'])
    with pytest.raises(ValueError, match=r'Must specify --num-bands and --band together'):
        kv.count.main(args)


def test_sketch_extensions():
    """kevlar/tests/test_sketch.py:102-106."""
    ge = kv.sketch.get_extension
    assert ge() == ('.nt', '.nodetable')
    assert ge(count=True) == ('.ct', '.counttable')
    assert ge(count=True, smallcount=True) == ('.sct', '.smallcounttable')
    assert ge(count=True, graph=True) == ('.cg', '.countgraph')
    assert ge(graph=True) == ('.ng', '.nodegraph')
    assert sorted(kv.sketch.sketch_loader_by_filename_extension) == sorted(
        ['.nt', '.ng', '.ct', '.cg', '.sct', '.scg', '.nodetable', '.nodegraph', '.counttable', '.countgraph',
         '.smallcounttable', '.smallcountgraph'])
    with pytest.raises(kv.sketch.KevlarSketchTypeError, match='sketch type from filename'):
        kv.sketch.load(golden_data('test.notasketchtype'))
    assert issubclass(kv.sketch.KevlarUnsuitableFPRError, SystemExit)


# ------------------------------------------------------------------ sequence / augmented FASTX

def test_augfastx_writer_literal():
    """kevlar/tests/test_seqio.py:135-182: exact text."""
    out = io.StringIO()
    kv.print_augmented_fastx(Record(
        name='BasiliscusVulgarisRead84467/1', sequence='TTAACTCTAGATTAGGGGCGTGACTTAATAAGGTGTGGGCCTAAGCGTCT',
        quality='B' * 50, annotations=[KmerOfInterest(19, 13, (12, 1, 1)), KmerOfInterest(19, 15, (20, 0, 1))]), out)
    kv.print_augmented_fastx(Record(
        name='BasiliscusVulgarisRead90577/2', sequence='CTGTAATCCCAGCACTTTGGGAGGCCGAGGCAAGCAGATGATGCGGTCAG',
        quality='B' * 50, annotations=[KmerOfInterest(19, 2, (7, 10, 9)), KmerOfInterest(19, 1, (5, 7, 9))],
        mates=['CAGATGTGTCTTGTGGGCAGTGCAGCGGAGAGGTGCAAATATGGGTTTGG']), out)
    kv.print_augmented_fastx(Record(name='r3', sequence='ACGT'), out)
    assert out.getvalue() == (
        '@BasiliscusVulgarisRead84467/1\nTTAACTCTAGATTAGGGGCGTGACTTAATAAGGTGTGGGCCTAAGCGTCT\n+\n' + 'B' * 50 + '\n'
        '             AGGGGCGTGACTTAATAAG          12 1 1#\n'
        '               GGGCGTGACTTAATAAGGT          20 0 1#\n'
        '@BasiliscusVulgarisRead90577/2\nCTGTAATCCCAGCACTTTGGGAGGCCGAGGCAAGCAGATGATGCGGTCAG\n+\n' + 'B' * 50 + '\n'
        ' TGTAATCCCAGCACTTTGG          5 7 9#\n'
        '  GTAATCCCAGCACTTTGGG          7 10 9#\n'
        '#mateseq=CAGATGTGTCTTGTGGGCAGTGCAGCGGAGAGGTGCAAATATGGGTTTGG#\n'
        '>r3\nACGT\n')


@pytest.mark.parametrize('fn', ['example1.augfastq', 'example2.augfastq', 'collect.alpha.txt', 'trio1/novel_3_1,2.txt'])
def test_augfastx_round_trip(fn):
    text = open(golden_data(fn)).read()
    out = io.StringIO()
    recs = list(kv.parse_augmented_fastx(io.StringIO(text)))
    for rec in recs:
        kv.print_augmented_fastx(rec, out)
    assert out.getvalue() == text
    for rec in recs:
        for ik in rec.annotations:
            assert rec.ikmerseq(ik) in rec.ikmers and kv.revcom(rec.ikmerseq(ik)) in rec.ikmers


def test_augfastx_reader_details():
    recs = list(kv.parse_augmented_fastx(open(golden_data('example2.augfastq'))))
    mated = list(kv.parse_augmented_fastx(io.StringIO('@r/1\nACGTACGT\n+\nIIIIIIII\n CGTAC          9 0#\n#mateseq=TTGACA#\n')))
    assert mated[0].mates == ['TTGACA'] and mated[0].annotations[0] == KmerOfInterest(5, 1, (9, 0))
    assert recs[0].id == recs[0].name.split()[0]
    with pytest.raises(Exception):
        list(kv.parse_augmented_fastx(io.StringIO('@r\nACGT\n+\nIIII\nnot an annotation\n')))
    with pytest.raises(AssertionError):
        Record('r', 'ACGTACGT').annotate('TTTT', 0, (1,))
    assert kv.revcom('ACGTNacgtRY') == 'RYACGTNACGT'
    assert kv.revcommin('TTTTA') == 'TAAAA' and kv.revcommin('AAAAT') == 'AAAAT'
    assert kv.same_seq('ACCG', 'CGGT')
    with pytest.raises(ValueError):
        kv.open('x', 'a')


# ------------------------------------------------------------------ FASTA/FASTQ reader

@pytest.mark.parametrize('fn', ['simple-genome-case-reads.fa.gz', 'trio1/case1.fq.gz', 'bogus-genome/refr.fa',
                                'microtrios/trio-na-proband.fq.gz', 'ambig.fasta', 'screen-case.fa'])
def test_fastx_reader_matches_oracle_parser(oracle, fn, monkeypatch):
    """Both readers -- the native one in libkvsketch.so (the product's ReadParser) and the pure
    Python one -- against the oracle's independent parser, record by record and batch by batch."""
    assert kv.khmer.ReadParser is fastx.NativeFastxReader
    want = [(r.name, r.sequence, r.quality) for r in oracle.ReadParser(golden_data(fn))]
    for cls, block in ((fastx.NativeFastxReader, None), (fastx.FastxReader, 32 << 20), (fastx.FastxReader, 4096),
                       (fastx.FastxReader, 333)):
        if block:
            monkeypatch.setattr(fastx.FastxReader, 'BLOCK', block)
        got = [(r.name, r.sequence, r.quality) for r in cls(golden_data(fn))]
        assert got == want
        reader = cls(golden_data(fn))
        seen = 0
        for batch in reader.batches(7000, keep_text=True):
            assert batch.offsets[0] == 0 and batch.offsets[-1] == len(batch.bases)
            for i in (0, len(batch) - 1):
                rec = batch.record(i)
                assert (rec.name, rec.sequence, rec.quality) == want[seen + i]
            last = want[seen + len(batch) - 1][0].encode()
            assert batch.name(batch.find_name(last)) == last and batch.find_name(b'no such read') == -1
            tail = batch.tail(len(batch) - 1)
            assert len(tail) == 1 and tail.record(0).sequence == want[seen + len(batch) - 1][1]
            seen += len(batch)
        assert seen == len(want) == reader.num_reads


def test_native_reader_edge_cases(tmp_path):
    """CRLF line ends, blank lines, a last record without newline, an empty sequence, and a
    missing file."""
    path = tmp_path / 'odd.fq'
    path.write_bytes(b'@r1 first\r\nACGT\r\n+\r\nIIII\r\n\n@r2\n\n+\n\n@r3\nGG\n+\nII')
    for cls in (fastx.NativeFastxReader, fastx.FastxReader):
        recs = [(r.name, r.sequence, r.quality) for r in cls(str(path))]
        assert recs == [('r1 first', 'ACGT', 'IIII'), ('r2', '', ''), ('r3', 'GG', 'II')], cls
    fa = tmp_path / 'multi.fa'
    fa.write_bytes(b'>chr1 desc\nACGT\nTTGA\n\n>chr2\nGG\n>empty\n')
    for cls in (fastx.NativeFastxReader, fastx.FastxReader):
        recs = [(r.name, r.sequence, r.quality) for r in cls(str(fa))]
        assert recs == [('chr1 desc', 'ACGTTTGA', None), ('chr2', 'GG', None), ('empty', '', None)], cls
    with pytest.raises(OSError):
        fastx.NativeFastxReader(str(tmp_path / 'missing.fq'))


def test_native_reader_large_inputs(tmp_path):
    """The native reader hands 4 MB blocks from its read-ahead thread to the record splitter:
    a ~10 MB FASTQ with ragged read lengths puts records across many block boundaries.  Plain,
    gzip (two concatenated members), sequences-only batches, a truncated last record, and closing
    a reader that was never drained (the producer thread must stop)."""
    import gzip
    rng = np.random.default_rng(99)
    letters = np.frombuffer(b'ACGTN', dtype=np.uint8)
    recs, chunks = [], []
    for i in range(45000):
        n = int(rng.integers(1, 400))
        seq = letters[rng.integers(0, 5, size=n)].tobytes()
        name = b'read%d len=%d' % (i, n)
        recs.append((name.decode(), seq.decode(), 'I' * n))
        chunks.append(b'@' + name + b'\n' + seq + b'\n+\n' + b'I' * n + b'\n')
    text = b''.join(chunks)
    assert len(text) > 2 * (4 << 20)
    plain = tmp_path / 'big.fq'
    plain.write_bytes(text)
    gz = tmp_path / 'big.fq.gz'
    half = len(chunks) // 2
    gz.write_bytes(gzip.compress(b''.join(chunks[:half]), 1) + gzip.compress(b''.join(chunks[half:]), 1))
    for path in (plain, gz):
        got = [(r.name, r.sequence, r.quality) for r in fastx.NativeFastxReader(str(path))]
        assert got == recs, path
        reader = fastx.NativeFastxReader(str(path))
        seqs = []
        for batch in reader.batches(1 << 20):            # sequences only: no header / quality text collected
            assert batch.name(0) == b'' and batch.qual(0) is None
            seqs.extend(batch.bases[int(batch.offsets[i]):int(batch.offsets[i + 1])].tobytes().decode()
                        for i in range(len(batch)))
        assert seqs == [r[1] for r in recs] and reader.num_reads == len(recs)
    cut = tmp_path / 'cut.fq'
    cut.write_bytes(b''.join(chunks[:3]) + b'@last one\nACGTAC\n+')
    for cls in (fastx.NativeFastxReader, fastx.FastxReader):
        got = [(r.name, r.sequence, r.quality) for r in cls(str(cut))]
        assert got == recs[:3] + [('last one', 'ACGTAC', '')], cls
    for _ in range(3):
        reader = fastx.NativeFastxReader(str(plain))
        next(iter(reader.batches(1000)))
        del reader


@pytest.mark.parametrize('threads,slice_bytes', [(1, 8 << 20), (4, 64), (7, 1000), (3, 70000)])
def test_native_reader_parallel_slices(tmp_path, monkeypatch, threads, slice_bytes):
    """Plain files are mapped and parsed slice by slice on several threads; slices must start at record
    boundaries.  Quality strings that start with '@', '>' or '+', sequences of length 0, CRLF, blank lines,
    multi-line FASTA, a FASTQ file that turns into FASTA half-way, junk before the first record: the same
    records as the pure-Python reader and as the serial native path (KV_READER_NO_MMAP), whatever the slicing
    and the batch size."""
    rng = np.random.default_rng(5)
    letters = np.frombuffer(b'ACGTN', dtype=np.uint8)
    qchars = np.frombuffer(b'@>+I#5', dtype=np.uint8)

    def fastq(n, crlf=False):
        out = []
        for i in range(n):
            ln = int(rng.integers(0, 60))
            seq = letters[rng.integers(0, 5, size=ln)].tobytes()
            qual = qchars[rng.integers(0, 6, size=ln)].tobytes()
            eol = b'\r\n' if crlf and i % 3 == 0 else b'\n'
            out.append(b'@r%d x\n' % i + seq + eol + b'+\n' + qual + eol + (b'\n' if i % 17 == 0 else b''))
        return b''.join(out)

    def fasta(n):
        out = []
        for i in range(n):
            out.append(b'>c%d\n' % i)
            for _ in range(int(rng.integers(0, 4))):
                out.append(letters[rng.integers(0, 5, size=int(rng.integers(1, 50)))].tobytes() + b'\n')
        return b''.join(out)

    files = {'a.fq': fastq(3000, crlf=True), 'b.fa': fasta(2000), 'c.fq': fastq(1500) + fasta(300) + fastq(5),
             'd.fq': b'\n\n' + fastq(800)[:-1], 'e.txt': b'junk line\n' + fastq(50), 'f.fq': b'@only\nACGT\n+\nIIII'}
    for name, text in files.items():
        path = tmp_path / name
        path.write_bytes(text)
        want = [(r.name, r.sequence, r.quality) for r in fastx.FastxReader(str(path))]
        monkeypatch.setenv('KV_READER_NO_MMAP', '1')
        serial = [(r.name, r.sequence, r.quality) for r in fastx.NativeFastxReader(str(path))]
        monkeypatch.delenv('KV_READER_NO_MMAP')
        assert serial == want, name
        monkeypatch.setenv('KV_READER_THREADS', str(threads))
        monkeypatch.setenv('KV_READER_SLICE_BYTES', str(slice_bytes))
        got = [(r.name, r.sequence, r.quality) for r in fastx.NativeFastxReader(str(path))]
        assert got == want, name
        for max_bases in (1, 500, 64 << 20):
            reader = fastx.NativeFastxReader(str(path))
            seqs = []
            for batch in reader.batches(max_bases):
                seqs.extend(batch.bases[int(batch.offsets[i]):int(batch.offsets[i + 1])].tobytes().decode() for i in range(len(batch)))
            assert seqs == [w[1] for w in want] and reader.num_reads == len(want), (name, max_bases)


def _bgzf(text, block=5000, level=6):
    """`text` as a BGZF (bgzip) file: independent gzip members with the 'BC' block-size subfield, then the empty
    end-of-file block."""
    import struct
    import zlib
    out = []
    for i in list(range(0, len(text), block)) + [None]:
        chunk = b'' if i is None else text[i:i + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        data = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(data) + 8 - 1
        out.append(b'\x1f\x8b\x08\x04' + b'\0' * 4 + b'\0\xff' + struct.pack('<H', 6) + b'BC' + struct.pack('<HH', 2, bsize) + data +
                   struct.pack('<II', zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    return b''.join(out)


@pytest.mark.parametrize('threads,slice_bytes', [(1, 8 << 20), (5, 3000), (4, 40000)])
def test_native_reader_bgzf_blocks_in_parallel(tmp_path, monkeypatch, threads, slice_bytes):
    """bgzip files: the blocks are inflated by the reader's thread pool window by window (records straddle block
    and window boundaries) and parsed by the same parallel parser; same records as Python's gzip module gives the
    pure-Python reader.  A flipped bit in a block is an error, not a short file."""
    import gzip
    rng = np.random.default_rng(8)
    letters = np.frombuffer(b'ACGTN', dtype=np.uint8)
    fq = b''.join(b'@read%d\n%s\n+\n%s\n' % (i, letters[rng.integers(0, 5, size=n)].tobytes(), b'@' * n)
                  for i, n in enumerate(rng.integers(0, 300, size=4000)))
    fa = b''.join(b'>seq%d\n' % i + b''.join(letters[rng.integers(0, 5, size=60)].tobytes() + b'\n' for _ in range(int(rng.integers(0, 400))))
                  for i in range(40))
    monkeypatch.setenv('KV_READER_THREADS', str(threads))
    monkeypatch.setenv('KV_READER_SLICE_BYTES', str(slice_bytes))
    for name, text in (('a.fq.gz', fq), ('b.fa.gz', fa), ('c.fq.gz', fq[:20000] + fa[:30000]), ('d.txt.gz', b'junk\n' + fq[:5000])):
        path = tmp_path / name
        path.write_bytes(_bgzf(text))
        assert gzip.open(str(path)).read() == text
        want = [(r.name, r.sequence, r.quality) for r in fastx.FastxReader(str(path))]
        got = [(r.name, r.sequence, r.quality) for r in fastx.NativeFastxReader(str(path))]
        assert got == want and len(want) > 5, name
        reader = fastx.NativeFastxReader(str(path))
        seqs = []
        for batch in reader.batches(20000):
            seqs.extend(batch.bases[int(batch.offsets[i]):int(batch.offsets[i + 1])].tobytes().decode() for i in range(len(batch)))
        assert seqs == [w[1] for w in want], name
    good = _bgzf(fq)
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0x10
    (tmp_path / 'bad.fq.gz').write_bytes(bytes(bad))
    with pytest.raises(OSError):
        for _ in fastx.NativeFastxReader(str(tmp_path / 'bad.fq.gz')).batches(1 << 20):
            pass


def test_native_reader_agrees_with_python_reader_on_line_soup(tmp_path, monkeypatch):
    """Property test: whatever lines a file is made of -- headers, '+' lines, quality strings that look like
    headers, blank lines, CRLF, junk, FASTA in the middle, no final newline -- the native reader (mapped, sliced
    every few hundred bytes over 3 threads, and as BGZF blocks) returns the records the pure-Python reader returns,
    whose rules are the reference parser's as far as the reference's files pin them.  (3000 examples each with
    3 threads x 200-byte, 5 x 64 and 2 x 1000 slices passed when this was written; 150 run here.)"""
    from hypothesis import given, settings, strategies as st, HealthCheck
    monkeypatch.setenv('KV_READER_THREADS', '3')
    monkeypatch.setenv('KV_READER_SLICE_BYTES', '200')
    seq = st.text(alphabet='ACGTNacgt', min_size=0, max_size=40)
    qual = st.text(alphabet='@>+I#!5', min_size=0, max_size=40)
    line = st.one_of(seq, qual, st.just(''), st.just('+'), st.text(alphabet='@>r1 x', min_size=1, max_size=8), st.just('junk'))
    fastq = st.tuples(st.text(alphabet='r12 /', min_size=0, max_size=6), seq).map(
        lambda t: '@{}\n{}\n+\n{}'.format(t[0], t[1], 'I' * len(t[1])))
    fasta = st.tuples(st.text(alphabet='c12', min_size=0, max_size=4), st.lists(seq, max_size=3)).map(
        lambda t: '>{}\n{}'.format(t[0], '\n'.join(t[1])))
    files = st.tuples(st.lists(st.one_of(fastq, fastq, fasta, line), min_size=0, max_size=40), st.sampled_from(['\n', '\r\n']),
                      st.booleans())
    counter = [0]

    @settings(max_examples=150, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
    @given(files)
    def check(spec):
        chunks, eol, final_newline = spec
        text = eol.join(c.replace('\n', eol) for c in chunks) + (eol if final_newline else '')
        counter[0] += 1
        path = tmp_path / 'soup{}.txt'.format(counter[0] % 4)
        path.write_bytes(text.encode())
        want = [(r.name, r.sequence, r.quality) for r in fastx.FastxReader(str(path))]
        got = [(r.name, r.sequence, r.quality) for r in fastx.NativeFastxReader(str(path))]
        assert got == want, text
        if text:
            gz = tmp_path / 'soup{}.gz'.format(counter[0] % 4)
            gz.write_bytes(_bgzf(text.encode(), block=97))
            got = [(r.name, r.sequence, r.quality) for r in fastx.NativeFastxReader(str(gz))]
            assert got == want, text
    check()


def test_fastx_reader_shared_by_threads():
    """kevlar/count.py:40-77: several consumers drain one parser; every read exactly once."""
    reader = kv.khmer.ReadParser(golden_data('trio1/case1.fq.gz'))
    got, lock = [], threading.Lock()

    def work():
        for batch in reader.batches(20000, keep_text=True):
            with lock:
                got.extend(batch.names)
    threads = [threading.Thread(target=work) for _ in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert len(got) == 12000 == len(set(got)) == reader.num_reads
    assert len(list(kv.multi_file_iter_khmer([golden_data('bogus-genome/mask-chr1.fa'),
                                              golden_data('bogus-genome/mask-chr2.fa')]))) == 4


# ------------------------------------------------------------------ timers / progress

def test_progress_indicator_batched_equals_stepwise(capsys):
    class Stepwise(object):   # the reference's one-record-at-a-time rule (kevlar/progress.py:31-42)
        def __init__(self, interval, breaks):
            self.counter, self.interval, self.nextupdate, self.breaks, self.out = 0, interval, interval, breaks, []

        def update(self):
            if self.counter in self.breaks:
                self.interval = self.counter
            if self.counter >= self.nextupdate:
                self.nextupdate += self.interval
                self.out.append(self.counter)
            self.counter += 1
    rng = np.random.default_rng(0)
    for interval, breaks in [(10, [100, 1000, 10000]), (1e2, [1e3, 1e4]), (7, [50, 51, 400])]:
        ref = Stepwise(interval, breaks)
        ind = kv.ProgressIndicator('{counter}', interval=interval, breaks=breaks)
        kv.logstream, saved = io.StringIO(), kv.logstream
        try:
            for n in rng.integers(1, 900, size=60):
                ind.update(int(n))
                for _ in range(int(n)):
                    ref.update()
            text = kv.logstream.getvalue()
        finally:
            kv.logstream = saved
        assert [float(x) for x in text.split()] == [float(x) for x in ref.out]
        assert ind.counter == ref.counter


def test_timer():
    timer = kv.Timer()
    timer.start()
    timer.start('x')
    with pytest.raises(ValueError, match='already started'):
        timer.start('x')
    with pytest.raises(ValueError, match='No timer started'):
        timer.stop('y')
    assert timer.probe('x') >= 0 and timer.stop('x') >= 0 and timer.stop() >= 0


# ------------------------------------------------------------------ unband

def test_unband_merges_band_outputs(tmp_path):
    """kevlar/unband.py:26-78: one record per read with the union of annotations, sorted by
    offset; deterministic across runs."""
    seq = 'TTAACTCTAGATTAGGGGCGTGACTTAATAAGGTGTGGGCCTAAGCGTCT'
    band1 = [Record('readA', seq, 'I' * 50, [KmerOfInterest(19, 15, (20, 0, 1))]),
             Record('readB', seq, 'I' * 50, [KmerOfInterest(19, 3, (9, 0, 0))])]
    band2 = [Record('readA', seq, 'I' * 50, [KmerOfInterest(19, 13, (12, 1, 1))]),
             Record('readC', seq, 'I' * 50, [KmerOfInterest(19, 1, (8, 1, 0))])]
    files = []
    for i, recs in enumerate((band1, band2)):
        path = str(tmp_path / 'band{}.augfastq'.format(i))
        with open(path, 'w') as fh:
            for rec in recs:
                kv.print_augmented_fastx(rec, fh)
        files.append(path)
    kv.logstream, saved = io.StringIO(), kv.logstream
    try:
        runs = [[(r.name, [a.offset for a in r.annotations]) for r in kv.unband.unband(kv.unband.afxstream(files), 4)]
                for _ in range(2)]
    finally:
        kv.logstream = saved
    assert runs[0] == runs[1]
    assert sorted(runs[0]) == [('readA', [13, 15]), ('readB', [3]), ('readC', [1])]


# ------------------------------------------------------------------ multi-GPU host logic

def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 300000, 300001):
        for world in (1, 2, 3, 4, 8):
            cuts = [multigpu.shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    for nbytes in (256, 4096, 64001024):
        for world in (2, 8):
            cuts = [multigpu.slice_bounds(nbytes, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == nbytes - nbytes % 256
            assert all(lo % 256 == 0 and hi % 256 == 0 for lo, hi in cuts)


def test_shard_batch_covers_every_read(oracle):
    seqs = [b'ACGT' * n for n in (3, 0, 10, 1, 25, 7, 7, 2, 40)]
    bases, offs = oracle.reads_to_batch(seqs)
    for world in (1, 2, 4, 9, 12):
        got = []
        for r in range(world):
            b, o = multigpu.shard_batch(bases, offs, r, world)
            assert o[0] == 0 and o[-1] == len(b)
            got += [b[int(o[i]):int(o[i + 1])].tobytes() for i in range(len(o) - 1)]
        assert got == seqs


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return str(s.getsockname()[1])


def test_gloo_two_rank_merge_equals_single_process(oracle, tmp_path):
    """world_size 2 over gloo: shard the reads, count per rank, merge with the all-reduce
    choreography; result must be byte-identical to one process counting everything."""
    port = _free_port()
    worker = os.path.join(REPO, 'tests', '_gloo_worker.py')
    procs = [subprocess.Popen([sys.executable, worker, str(r), '2', port, str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    res = [np.load(str(tmp_path / 'rank{}.npy'.format(r)), allow_pickle=True)[0] for r in range(2)]
    rng = np.random.default_rng(1234)
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)
    genome = letters[rng.integers(0, 4, size=3000)]
    seqs = []
    for _ in range(4000):
        s = int(rng.integers(0, 2900))
        seqs.append(genome[s:s + int(rng.integers(25, 100))].tobytes())
    bases, offs = oracle.reads_to_batch(seqs)
    for name in ('Counttable', 'SmallCounttable', 'Nodetable'):
        sk = getattr(oracle, name)(21, 900, 4)
        sk.consume_batch(bases, offs)
        want = [sk.table_bytes(t) for t in range(4)]
        assert res[0][name] == want and res[1][name] == want, name
        if name == 'Counttable':
            assert max(max(t) for t in want) == 255      # saturation really happened
    assert res[0]['hit_reads'] == [0, 1, 2, 2000, 2001, 2002] == res[1]['hit_reads']


# ------------------------------------------------------------------ kevlar dist host logic

DIST_ABUND = {10: 6, 11: 10, 12: 12, 13: 18, 14: 16, 15: 11, 16: 9, 17: 9, 18: 11, 19: 8, 20: 9, 21: 7, 22: 3}


def test_dist_mu_sigma():
    """kevlar/tests/test_dist.py:46-56."""
    from kevlar_b200.dist import calc_mu_sigma, KevlarZeroAbundanceDistError
    mu, sigma = calc_mu_sigma(DIST_ABUND)
    assert mu == pytest.approx(15.32558, abs=1e-5)
    assert sigma == pytest.approx(3.280581, abs=1e-5)
    with pytest.raises(KevlarZeroAbundanceDistError, match='all k-mer abundances are 0'):
        calc_mu_sigma({})


def test_dist_table_matches_shipped_tsv(tmp_path):
    """kevlar/tests/test_dist.py:59-66 and the byte layout of minitrio/trio-proband-dist.tsv."""
    from kevlar_b200.dist import compute_dist
    data = compute_dist(DIST_ABUND)
    assert list(data['Count'][:5]) == [6.0, 10.0, 12.0, 18.0, 16.0]
    assert list(data['CumulativeCount'][:5]) == [6.0, 16.0, 28.0, 46.0, 62.0]
    shipped = open(golden_data('minitrio/trio-proband-dist.tsv')).read()
    rows = [line.split('\t') for line in shipped.splitlines()[1:]]
    abundance = {int(float(r[0])): int(float(r[1])) for r in rows}
    out = str(tmp_path / 'dist.tsv')
    compute_dist(abundance).to_csv(out, sep='\t', index=False)
    assert open(out).read() == shipped


def test_dist_cli_arguments():
    """kevlar/cli/dist.py:14-46: flags and defaults."""
    import kevlar_b200
    args = kevlar_b200.cli.parser().parse_args(['dist', 'mask.nt', 'a.fq', 'b.fq'])
    assert (args.cmd, args.mask, args.infiles) == ('dist', 'mask.nt', ['a.fq', 'b.fq'])
    assert (args.ksize, args.memory, args.threads, args.plot, args.tsv, tuple(args.plot_xlim)) == \
        (31, 1e6, 1, None, None, (0, 100))
    args = kevlar_b200.cli.parser().parse_args(['dist', '-k', '25', '-M', '4M', '--tsv', 'o.tsv', '--plot-xlim', '0', '50',
                                                'mask.nt', 'a.fq'])
    assert (args.ksize, args.memory, args.tsv, list(args.plot_xlim)) == (25, 4e6, 'o.tsv', [0, 50])


# ------------------------------------------------------------------ simlike host logic (no GPU)

class _OracleQueries(object):
    """Gives an oracle sketch the batched query method the product's sketches have, so the host
    side of kevlar_b200.simlike can run on the CPU."""

    def __init__(self, sketch):
        self.sketch = sketch

    def ksize(self):
        return self.sketch.ksize()

    def get_kmer_counts_many(self, sequences):
        k = self.sketch.ksize()
        return [np.array(self.sketch.get_kmer_counts(s) if len(s) >= k else [], dtype=np.uint8) for s in sequences]


def test_simlike_host_logic_against_reference_outputs(oracle):
    """kevlar/simlike.py:22-96: filtering by reference-genome abundance, outlier dropping, SNV vs
    indel pairing -- the 52 windows the reference's own function produced over the oracle
    (tests/golden/gen/simlike_spanning.json), batched and one by one."""
    import json
    from conftest import golden_gen
    from kevlar_b200.simlike import (spanning_kmer_abundances, spanning_kmer_abundances_many, discard_nonunique_kmers,
                                     discard_outlier_abunds)
    kid, mom, dad = (oracle.Counttable(31, 1e6, 4) for _ in range(3))
    ref = oracle.SmallCounttable(31, 125000, 4)
    for sk, fn in ((kid, 'trio-proband.fq.gz'), (mom, 'trio-mother.fq.gz'), (dad, 'trio-father.fq.gz'), (ref, 'refr.fa.gz')):
        sk.consume_seqfile(oracle.ReadParser(golden_data('minitrio/' + fn)))
    kid, mom, dad, ref = (_OracleQueries(s) for s in (kid, mom, dad, ref))
    cases = json.load(open(golden_gen('simlike_spanning.json')))
    for drop in (False, True):
        subset = [c for c in cases if c['dropoutliers'] == drop]
        got = spanning_kmer_abundances_many([(c['alt'], c['refr']) for c in subset], kid, (mom, dad), ref, dropoutliers=drop)
        assert [list(g) for g in got] == [[c['abundances'], c['refr_abunds'], c['ndropped']] for c in subset]
    first = cases[0]
    assert spanning_kmer_abundances(first['alt'], first['refr'], kid, (mom, dad), ref) == \
        (first['abundances'], first['refr_abunds'], 3)
    case_counts, ctrl_counts, alt_in_refr = discard_nonunique_kmers(first['alt'], kid, (mom, dad), ref)
    assert case_counts == first['abundances'][0] and ctrl_counts == first['abundances'][1:]
    assert len(alt_in_refr) == 31 and sum(1 for r in alt_in_refr if r) == 3
    # the reference's sketch + VCF fixtures (k = 49 and 31; 4-bit reference sketches; a one-bucket table)
    fixture = json.load(open(golden_gen('simlike_fixture_windows.json')))
    assert len(fixture) == 105
    loaded = {}
    for c in fixture:
        key = (c['set'],) + tuple(c['sketches'])
        if key not in loaded:
            loaded[key] = [_OracleQueries((oracle.SmallCounttable if f.endswith('.sct') else oracle.Counttable)
                                          .load(golden_data(c['set'] + '/' + f))) for f in c['sketches']]
        sk = loaded[key]
        got = spanning_kmer_abundances(c['alt'], c['refr'], sk[0], sk[1:-1], sk[-1], dropoutliers=c['dropoutliers'])
        assert got == (c['abundances'], c['refr_abunds'], c['ndropped'])
    assert discard_outlier_abunds([10, 11, 12, 90], [[1, 1, 1, 1], [0, 40, 0, 0]]) == ([11, 12], [[1, 1, 1, 1], [0, 0, 0]])


def test_band_assignment_covers_every_band_once():
    """Config 5 driver: band b runs on rank (b-1) mod world -- every band exactly once for any world size."""
    from kevlar_b200 import bands
    for num_bands in (1, 2, 8, 16, 23):
        for world in (1, 2, 3, 8):
            got = sorted(b for r in range(world) for b in bands.band_of_rank(num_bands, r, world))
            assert got == list(range(1, num_bands + 1))
    ns = bands.parser().parse_args(['--case', 'a.fq', 'b.fq', '--control', 'c.fq', '--num-bands', '4', '--out-prefix', 'x'])
    argv = bands.novel_band_args(ns, 3, 'x.band3.augfastq')
    assert argv[:1] == ['novel'] and argv[argv.index('--band') + 1] == '3' and argv[argv.index('--case') + 1:][:2] == ['a.fq', 'b.fq']


def test_simtrio_piecewise_haplotypes_equal_sequential():
    """The measurement fixture builds haplotypes from pieces in one pass; for the benchmark trio (1 Mbp) the
    result must be what the original edit-by-edit construction gave, so the C2 workload is unchanged."""
    from kevlar_b200 import simtrio
    new = simtrio.trio_haplotypes(1000000)
    piecewise = simtrio._apply
    simtrio._apply = simtrio._apply_sequential
    try:
        old = simtrio.trio_haplotypes(1000000)
    finally:
        simtrio._apply = piecewise
    for a, b in zip(new, old):
        for x, y in zip(a, b):
            assert len(x) == len(y) and (x == y).all()


def test_native_reader_reports_damaged_gzip_and_keeps_prefetched_batches(tmp_path):
    """(a) a truncated .gz must raise OSError like khmer's ReadParser instead of ending the file
    silently; (b) a batches() generator abandoned early must not lose the batch it had parsed ahead."""
    import gzip
    from kevlar_b200 import fastx
    reads = ['@r{}\nACGTACGTACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n'.format(i) for i in range(20000)]
    good = tmp_path / 'reads.fq.gz'
    with gzip.open(str(good), 'wt') as fh:
        fh.write(''.join(reads))
    raw = open(str(good), 'rb').read()
    bad = tmp_path / 'cut.fq.gz'
    open(str(bad), 'wb').write(raw[:len(raw) // 2])
    with pytest.raises(OSError):
        for _ in fastx.NativeFastxReader(str(bad)).batches(64 << 10):
            pass
    parser = fastx.NativeFastxReader(str(good))
    gen = parser.batches(100 * 32)          # 100 reads per batch, one batch parsed ahead
    first = next(gen)
    gen.close()
    rest = sum(len(b) for b in parser.batches(100 * 32))
    assert len(first) + rest == 20000 == parser.num_reads
