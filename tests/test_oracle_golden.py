"""Pins the CPU oracle (oracle/) to the reference's golden vectors -- no GPU needed.

The oracle is a restatement of the un-vendored khmer dependency; these tests are what makes
it trustworthy as the parity checker for the CUDA path (prompt section 3).
"""
import filecmp
import gzip
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_data, golden_gen


@pytest.mark.parametrize('infile,golden,numbands,band,n_unique,n_occupied', [
    ('case', 'case', 0, 0, 973, 801),
    ('ctrl1', 'ctrl1', 0, 0, 973, 791),
    ('ctrl2', 'ctrl2', 0, 0, 966, 800),
    ('case', 'case-band-2-1', 2, 1, 501, 444),
    ('case', 'case-band-16-7', 16, 7, 68, 67),
])
def test_count_golden_sketches(oracle, tmp_path, infile, golden, numbands, band, n_unique, n_occupied):
    """kevlar/tests/test_count.py:45-68: byte-for-byte."""
    sk = oracle.Counttable(25, 10e3 / 4, 4)
    assert sk.hashsizes() == [2477, 2473, 2467, 2459]
    parser = oracle.ReadParser(golden_data('simple-genome-{}-reads.fa.gz'.format(infile)))
    if numbands:
        nreads, _ = sk.consume_seqfile_banding(parser, numbands, band - 1)
    else:
        nreads, _ = sk.consume_seqfile(parser)
    assert nreads == 600
    assert sk.n_unique_kmers() == n_unique
    assert sk.n_occupied() == n_occupied
    out = str(tmp_path / 'o.ct')
    sk.save(out)
    assert filecmp.cmp(out, golden_data('simple-genome-{}.ct'.format(golden)), shallow=False)


@pytest.mark.parametrize('filename,cls,testkmer', [
    ('test.countgraph', 'Countgraph', 'TGGAACCGGCAACGACGAAAA'),
    ('test.smallcountgraph', 'SmallCountgraph', 'CTGTACTACAGCTACTACAGT'),
    ('test.counttable', 'Counttable', 'CCTGATATCCGGAATCTTAGC'),
    ('test.smallcounttable', 'SmallCounttable', 'GGGCCCCCATCTCTATCTTGC'),
    ('test.nodegraph', 'Nodegraph', 'GGGAACTTACCTGGGGGTGCG'),
    ('test.nodetable', 'Nodetable', 'CTGTTCGATATGAGGAATCTG'),
])
def test_sketch_files(oracle, tmp_path, filename, cls, testkmer):
    """kevlar/tests/test_sketch.py:17-29 + header/occupancy consistency + save round trip."""
    sk = getattr(oracle, cls).load(golden_data(filename))
    assert sk.ksize() == 21
    assert sk.get(testkmer) > 0
    assert sk.get('GATTACA' * 3) == 0
    assert sk.n_occupied() == oracle._lib.ko_count_occupied(sk._h)
    out = str(tmp_path / filename)
    sk.save(out)
    assert filecmp.cmp(out, golden_data(filename), shallow=False)


def test_primes(oracle):
    """SURVEY App. A.1 vectors."""
    assert oracle.primes_below(2500, 4) == [2477, 2473, 2467, 2459]
    assert oracle.primes_below(250000, 4) == [249989, 249973, 249971, 249967]
    assert oracle.primes_below(125, 4) == [113, 109, 107, 103]
    assert oracle.primes_below(1e4, 4) == [9973, 9967, 9949, 9941]
    assert oracle.primes_below(100, 4) == [97, 89, 83, 79]


def test_band_intervals(oracle):
    lo, hi = oracle.band_interval(2, 0)
    assert lo == 0 and hi == (2 ** 64 - 1) // 2
    lo, hi = oracle.band_interval(16, 15)
    assert lo == 15 * ((2 ** 64 - 1) // 16) and hi == 2 ** 64 - 1
    with pytest.raises(ValueError):
        oracle.band_interval(4, 4)


def test_murmur_known_answers(oracle):
    """MurmurHash3_x64_128 low word, seed 0 (published smhasher behaviour)."""
    assert oracle.murmur3_lo(b'') == 0
    assert oracle.murmur3_lo(b'hello') == 0xcbd8a7b341bd9b02
    assert oracle.murmur3_lo(b'The quick brown fox jumps over the lazy dog') == 0xe34bbc7bbc071b6c


def test_strand_symmetry(oracle):
    """kevlar/tests/test_novel.py:68-77."""
    ct = oracle.Counttable(27, 1e5, 2)
    cg = oracle.Countgraph(27, 1e5, 2)
    comp = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}
    for kmer in ('ACCGTACAA' * 3, 'TTATAATAG' * 3, 'CGAAAAATT' * 3):
        rc = ''.join(comp[c] for c in reversed(kmer))
        assert ct.hash(kmer) == ct.hash(rc)
        assert cg.hash(kmer) == cg.hash(rc)
        assert cg.reverse_hash(cg.hash(kmer)) in (kmer, rc)
    with pytest.raises(ValueError, match='not implemented'):
        ct.reverse_hash(5)


def test_masked_count_pin(oracle):
    """kevlar/tests/test_count.py:153-166: 36898 distinct k-mers with the mask polarity of App. A.8."""
    mask = oracle.Nodetable(21, 1e4, 4)
    mask.consume('CACCAATCCGTACGGAGAGCCGTATATATAGACTGCTATACTATTGGATCGTACGGGGC')
    sk = oracle.Counttable(21, 1e6 / 4, 4)
    sk.consume_seqfile_with_mask(oracle.ReadParser(golden_data('bogus-genome/refr.fa')), mask, threshold=0,
                                 consume_masked=False)
    assert sk.n_unique_kmers() == 36898
    sk2 = oracle.Counttable(21, 1e6 / 4, 4)
    sk2.consume_seqfile_with_mask(oracle.ReadParser(golden_data('bogus-genome/refr.fa')), mask, threshold=1,
                                  consume_masked=True)
    assert sk2.n_unique_kmers() == 39
    assert sk2.get('CACCAATCCGTACGGAGAGCC') > 0 and sk2.get('GAATCGGTGGCTGGTTGCCGT') == 0


def test_novel_scan_pins(oracle):
    """kevlar/tests/test_novel.py:179-194 ('29 unique novel kmers in 14 reads' after skipping
    1001 reads) and the shipped microtrio output, through the oracle's batch scan."""
    comp = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}
    sketches = []
    for name in ('case1', 'ctrl1', 'ctrl2'):
        sk = oracle.Counttable(31, 1e6 / 4, 4)
        sk.consume_seqfile(oracle.ReadParser(golden_data('trio1/{}.fq.gz'.format(name))))
        sketches.append(sk)
    reads = list(oracle.ReadParser(golden_data('trio1/case1.fq.gz')))
    names = [r.name for r in reads]
    at = names.index('bogus-genome-chr1_115_449_0:0:0_0:0:0_1f4/1')
    assert at + 1 == 1001
    seqs = [r.sequence for r in reads[at + 1:]]
    bases, offs = oracle.reads_to_batch(seqs)
    hits, flags = oracle.novel_batch(sketches[:1], sketches[1:], bases, offs, 6, 0)
    kmers = set()
    for h in hits:
        km = seqs[int(h['read'])][int(h['offset']):int(h['offset']) + 31]
        rc = ''.join(comp[c] for c in reversed(km))
        kmers.add(min(km, rc))
    assert len(set(hits['read'].tolist())) == 14 and len(kmers) == 29 and len(hits) == 158

    # shipped `kevlar novel` output: microtrios/novel-na.augfastq.gz (k=31, case-min 5, ctrl-max 1, 500K)
    sk = []
    for who in ('proband', 'mother', 'father'):
        s = oracle.Counttable(31, 5e5 / 4, 4)
        s.consume_seqfile(oracle.ReadParser(golden_data('microtrios/trio-na-{}.fq.gz'.format(who))))
        sk.append(s)
    reads = list(oracle.ReadParser(golden_data('microtrios/trio-na-proband.fq.gz')))
    bases, offs = oracle.reads_to_batch([r.sequence for r in reads])
    hits, _ = oracle.novel_batch(sk[:1], sk[1:], bases, offs, 5, 1)
    lines = []
    last = None
    for h in hits:
        r = reads[int(h['read'])]
        if h['read'] != last:
            lines += ['@' + r.name, r.sequence, '+', r.quality]
            last = h['read']
        o = int(h['offset'])
        lines.append(' ' * o + r.sequence[o:o + 31] + ' ' * 10 + ' '.join(str(int(a)) for a in h['abund'][:3]) + '#')
    shipped = gzip.open(golden_data('microtrios/novel-na.augfastq.gz'), 'rt').read().splitlines()
    shipped = [ln for ln in shipped if not ln.startswith('#mateseq=')]
    assert lines == shipped
    assert '\n'.join(lines) + '\n' == open(golden_gen('novel_microtrio_na.out')).read()


def test_threaded_consume_equals_serial(oracle):
    rng = np.random.default_rng(3)
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)
    seqs = [letters[rng.integers(0, 4, size=int(rng.integers(10, 200)))].tobytes() for _ in range(3000)]
    bases, offs = oracle.reads_to_batch(seqs)
    for cls in ('Counttable', 'SmallCounttable', 'Nodetable', 'Countgraph'):
        a = getattr(oracle, cls)(21, 2000, 4)
        b = getattr(oracle, cls)(21, 2000, 4)
        assert a.consume_batch(bases, offs, threads=1) == b.consume_batch(bases, offs, threads=4)
        for t in range(4):
            assert a.table_bytes(t) == b.table_bytes(t)
        assert a.n_occupied() == b.n_occupied()


def test_manifest_and_reference_run():
    """The committed fixtures are the ones make_golden.py recorded, and the reference's own
    test files passed on top of the oracle when they were generated."""
    import hashlib
    manifest = json.load(open(os.path.join(GOLDEN, 'MANIFEST.json')))
    checked = 0
    for rel, meta in manifest.items():
        path = os.path.join(GOLDEN, rel if rel.startswith('gen/') else os.path.join('data', rel))
        if meta.get('stored') is False or 'sha256' not in meta:
            continue
        assert hashlib.sha256(open(path, 'rb').read()).hexdigest() == meta['sha256'], rel
        checked += 1
    assert checked > 40
    log = open(os.path.join(GOLDEN, 'reference_tests_over_oracle.log')).read()
    assert '130 passed' in log and 'failed' not in log


DIST_ABUND_10K = {10: 6, 11: 10, 12: 12, 13: 18, 14: 16, 15: 11, 16: 9, 17: 9, 18: 11, 19: 8, 20: 9, 21: 7, 22: 3}


def test_dist_first_pass_golden(oracle, tmp_path):
    """kevlar/tests/test_dist.py:25-33: masked count of the minitrio proband, byte-for-byte."""
    mask = oracle.Nodetable.load(golden_data('minitrio/mask.nt'))
    counts = oracle.Counttable(31, 1e4, 4)
    counts.consume_seqfile_with_mask(oracle.ReadParser(golden_data('minitrio/trio-proband.fq.gz')), mask,
                                     threshold=1, consume_masked=True)
    out = str(tmp_path / 'first.ct')
    counts.save(out)
    assert filecmp.cmp(out, golden_data('minitrio/trio-proband-mask-counts.ct'), shallow=False)


def test_dist_second_pass_golden(oracle):
    """kevlar/tests/test_dist.py:36-43: abundance_distribution with a tracking Nodetable built on
    the counts' table sizes."""
    counts = oracle.Counttable.load(golden_data('minitrio/trio-proband-mask-counts.ct'))
    tracking = oracle.Nodetable(31, 1, 1, primes=counts.hashsizes())
    assert tracking.hashsizes() == counts.hashsizes()
    dist = counts.abundance_distribution(oracle.ReadParser(golden_data('minitrio/trio-proband.fq.gz')), tracking)
    assert len(dist) == 65536
    assert {a: c for a, c in enumerate(dist) if a > 0 and c > 0} == DIST_ABUND_10K


def test_dist_default_memory_matches_shipped_tsv(oracle):
    """kevlar/tests/test_dist.py:106-122 + the shipped minitrio/trio-proband-dist.tsv: both passes
    at the CLI's default --memory 1M."""
    mask = oracle.Nodetable.load(golden_data('minitrio/mask.nt'))
    reads = golden_data('minitrio/trio-proband.fq.gz')
    counts = oracle.Counttable(31, 1e6 / 4, 4)
    counts.consume_seqfile_with_mask(oracle.ReadParser(reads), mask, threshold=1, consume_masked=True)
    tracking = oracle.Nodetable(31, 1, 1, primes=counts.hashsizes())
    dist = counts.abundance_distribution(oracle.ReadParser(reads), tracking)
    rows = [line.split('\t') for line in open(golden_data('minitrio/trio-proband-dist.tsv')).read().splitlines()[1:]]
    assert {a: c for a, c in enumerate(dist) if a > 0 and c > 0} == {int(float(r[0])): int(float(r[1])) for r in rows}


def test_simlike_spanning_abundances_literal(oracle):
    """kevlar/tests/test_simlike.py:82-107: the literal abundance lists of the minitrio window, as
    produced by the reference's own spanning_kmer_abundances over the oracle (gen/simlike_spanning.json)
    and directly through the oracle's get_kmer_counts here."""
    cases = json.load(open(golden_gen('simlike_spanning.json')))
    first = cases[0]
    assert first['ndropped'] == 3 and not first['dropoutliers']
    assert first['abundances'] == [
        [7, 6, 6, 6, 6, 6, 6, 6, 6, 6, 7, 9, 8, 8, 9, 9, 9, 7, 7, 8, 8, 8, 7, 7, 7, 7, 7, 7],
        [1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1],
        [0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    ]
    assert first['refr_abunds'] == [2, 2, 1, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 1, 1]
    indel = cases[2]
    assert indel['ndropped'] == 3 and indel['refr_abunds'] == [None] * 28
    kid = oracle.Counttable(31, 1e6, 4)
    kid.consume_seqfile(oracle.ReadParser(golden_data('minitrio/trio-proband.fq.gz')))
    ref = oracle.SmallCounttable(31, 125000, 4)
    ref.consume_seqfile(oracle.ReadParser(golden_data('minitrio/refr.fa.gz')))
    keep = [r == 0 for r in ref.get_kmer_counts(first['alt'])]
    assert [c for c, ok in zip(kid.get_kmer_counts(first['alt']), keep) if ok] == first['abundances'][0]


SIMLIKE_SKETCHES = [
    'case-low-abund/dad.ct', 'case-low-abund/kid.ct', 'case-low-abund/mom.ct', 'case-low-abund/refr.sct',
    'ctrl-high-abund/cc57120.dad.sct', 'ctrl-high-abund/cc57120.kid.sct', 'ctrl-high-abund/cc57120.mom.sct',
    'ctrl-high-abund/cc57120.refr.sct', 'term-high-abund/father.ct', 'term-high-abund/mother.ct',
    'term-high-abund/proband.ct', 'term-high-abund/reference.sct', 'partscore/partscore-father.ct',
    'partscore/partscore-mother.ct', 'partscore/partscore-proband.ct', 'partscore/partscore-refr.sct',
    'simlike-fast-mode/cc27.dad.ct', 'simlike-fast-mode/cc27.kid.ct', 'simlike-fast-mode/cc27.mom.ct',
    'simlike-fast-mode/cc27.refr.sct', 'homopolymer/12175-dad.sct', 'homopolymer/12175-kid.sct',
    'homopolymer/12175-mom.sct', 'homopolymer/12175-refr.sct',
]


@pytest.mark.parametrize('rel', SIMLIKE_SKETCHES)
def test_more_reference_sketch_files(oracle, tmp_path, rel):
    """The 24 small sketches of the reference's simlike fixtures (8-bit .ct and 4-bit .sct, k = 31
    and 49, two of them with a single table): the occupancy recorded in the header equals a
    recount and a re-save is byte-identical."""
    cls = oracle.SmallCounttable if rel.endswith('.sct') else oracle.Counttable
    sk = cls.load(golden_data(rel))
    assert sk.ksize() in (31, 49)
    assert sk.n_occupied() == oracle._lib.ko_count_occupied(sk._h)
    out = str(tmp_path / 'resaved')
    sk.save(out)
    assert filecmp.cmp(out, golden_data(rel), shallow=False)


def test_single_bucket_table(oracle):
    """khmer sizes one table "near 1" as a table of one bucket: the reference fixture
    term-high-abund/reference.sct is SmallCounttable(31, 1, 1) and kevlar/tests/test_simlike.py:69
    builds Nodetable(31, 1, 1)."""
    assert oracle.primes_below(1, 1) == [1]
    shipped = oracle.SmallCounttable.load(golden_data('term-high-abund/reference.sct'))
    assert shipped.hashsizes() == [1] and shipped.n_occupied() == 0
    node = oracle.Nodetable(31, 1, 1)
    assert node.hashsizes() == [1] and node.get('ACGT' * 7 + 'ACG') == 0
    node.consume('ACGT' * 10)
    assert node.get('TTTT' * 7 + 'TTT') == 1 and node.n_occupied() == 1   # every k-mer shares the one bucket
    with pytest.raises(ValueError):
        oracle.primes_below(1, 2)
