"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden vectors.

Everything here is integer/byte work: the bar is bit-exact.  Run on the B200 box with
``pytest -m gpu``.  Nothing in this file reads /root/reference.
"""
import filecmp
import gzip
import io
import os
import re

import numpy as np
import pytest

from conftest import golden_data, golden_gen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def kv():
    import kevlar_b200
    if kevlar_b200._lib.device_count() < 1:
        pytest.fail('no CUDA device: the gpu tests must run on the B200 box')
    return kevlar_b200


LETTERS = np.frombuffer(b'ACGT', dtype=np.uint8)


def random_reads(seed, n, lo=20, hi=160, alphabet=b'ACGT', genome=None):
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(alphabet, dtype=np.uint8)
    out = []
    for _ in range(n):
        length = int(rng.integers(lo, hi + 1))
        if genome is not None:
            start = int(rng.integers(0, len(genome) - length))
            out.append(genome[start:start + length].tobytes())
        else:
            out.append(letters[rng.integers(0, len(letters), size=length)].tobytes())
    return out


def assert_same_sketch(gpu, cpu, check_unique=True):
    assert gpu.hashsizes() == cpu.hashsizes()
    for t in range(len(cpu.hashsizes())):
        assert gpu.table_bytes(t) == cpu.table_bytes(t), 'table {} differs'.format(t)
    assert gpu.n_occupied() == cpu.n_occupied()
    if check_unique:
        assert gpu.n_unique_kmers() == cpu.n_unique_kmers()


CLASSES = ['Counttable', 'SmallCounttable', 'Nodetable', 'Countgraph', 'SmallCountgraph', 'Nodegraph']


# ------------------------------------------------------------------ hashing

@pytest.mark.parametrize('k', [1, 4, 13, 15, 16, 17, 19, 21, 25, 27, 31, 32, 33, 35, 47, 48, 49, 63, 64])
def test_murmur_hash_matches_oracle(kv, oracle, k):
    rng = np.random.default_rng(k)
    kmers = [LETTERS[rng.integers(0, 4, size=k)].tobytes().decode() for _ in range(300)]
    g = kv.khmer.Counttable(k, 1000, 2)
    c = oracle.Counttable(k, 1000, 2)
    got = g.hash_many(kmers)
    want = np.array([c.hash(km) for km in kmers], dtype=np.uint64)
    assert (got == want).all()
    # strand symmetry (kevlar/tests/test_novel.py:68-77)
    assert g.hash(kmers[0]) == g.hash(kv.revcom(kmers[0]))


@pytest.mark.parametrize('k', [1, 5, 15, 16, 17, 21, 31, 32])
def test_twobit_hash_matches_oracle(kv, oracle, k):
    rng = np.random.default_rng(100 + k)
    kmers = [LETTERS[rng.integers(0, 4, size=k)].tobytes().decode() for _ in range(300)]
    g = kv.khmer.Countgraph(k, 1000, 2)
    c = oracle.Countgraph(k, 1000, 2)
    got = g.hash_many(kmers)
    want = np.array([c.hash(km) for km in kmers], dtype=np.uint64)
    assert (got == want).all()
    assert kv.same_seq(g.reverse_hash(g.hash(kmers[3])), kmers[3])


def test_hash_rejects_non_acgt(kv):
    g = kv.khmer.Counttable(21, 1000, 2)
    with pytest.raises(ValueError):
        g.hash('ACGTACGTACNTACGTACGTA')
    with pytest.raises(ValueError):
        g.hash('ACGT')
    t = kv.khmer.Counttable(35, 1e4, 4)
    with pytest.raises(ValueError, match=r'not implemented'):
        t.reverse_hash(t.hash('CTATGGCGGAAGGGCACACCTAACCGCGATGACGG'))


def test_get_kmer_hashes_and_counts(kv, oracle):
    seq = 'TGCCACGATCCGGCTATGGCGGAAGGGCACACCTAACCGCGATGACGGAGTAACTCGCAGCA'
    for name in ('Counttable', 'Countgraph', 'SmallCounttable', 'Nodegraph'):
        g = getattr(kv.khmer, name)(21, 1e4, 4)
        c = getattr(oracle, name)(21, 1e4, 4)
        g.consume(seq)
        g.consume(seq[10:50])
        c.consume(seq)
        c.consume(seq[10:50])
        assert g.get_kmer_hashes(seq) == c.get_kmer_hashes(seq)
        assert g.get_kmer_counts(seq) == c.get_kmer_counts(seq)
        assert g.get_kmers(seq) == c.get_kmers(seq)
        assert g.get(seq[:21]) == c.get(seq[:21]) > 0
        assert g.get('GATTACA' * 3) == 0


# ------------------------------------------------------------------ golden sketches

@pytest.mark.parametrize('infile,testout,numbands,band,kmers_stored', [
    ('case', 'case', 0, 0, 973),
    ('ctrl1', 'ctrl1', 0, 0, 973),
    ('ctrl2', 'ctrl2', 0, 0, 966),
    ('case', 'case-band-2-1', 2, 1, 501),
    ('case', 'case-band-16-7', 16, 7, 68),
])
def test_count_simple_golden(kv, tmp_path, capsys, infile, testout, numbands, band, kmers_stored):
    """kevlar/tests/test_count.py:45-68 -- `kevlar count` output byte-identical to the reference's."""
    out = str(tmp_path / 'out')
    arglist = ['count', '--ksize', '25', '--memory', '10K', '--num-bands', str(numbands), '--band', str(band),
               out, golden_data('simple-genome-{}-reads.fa.gz'.format(infile))]
    args = kv.cli.parser().parse_args(arglist)
    kv.count.main(args)
    err = capsys.readouterr().err
    assert '600 reads processed' in err
    assert '{:d} distinct k-mers stored'.format(kmers_stored) in err
    assert filecmp.cmp(out + '.counttable', golden_data('simple-genome-{}.ct'.format(testout)), shallow=False)


@pytest.mark.parametrize('filename,testkmer', [
    ('test.countgraph', 'TGGAACCGGCAACGACGAAAA'),
    ('test.smallcountgraph', 'CTGTACTACAGCTACTACAGT'),
    ('test.counttable', 'CCTGATATCCGGAATCTTAGC'),
    ('test.smallcounttable', 'GGGCCCCCATCTCTATCTTGC'),
    ('test.nodegraph', 'GGGAACTTACCTGGGGGTGCG'),
    ('test.nodetable', 'CTGTTCGATATGAGGAATCTG'),
])
def test_sketch_load_golden(kv, tmp_path, filename, testkmer):
    """kevlar/tests/test_sketch.py:17-29, plus a byte-exact save round trip."""
    sketch = kv.sketch.load(golden_data(filename))
    assert sketch.get(testkmer) > 0
    assert sketch.get('GATTACA' * 3) == 0
    out = str(tmp_path / filename)
    sketch.save(out)
    assert filecmp.cmp(out, golden_data(filename), shallow=False)


def test_sketch_load_errors(kv, tmp_path):
    with pytest.raises(kv.sketch.KevlarSketchTypeError, match='sketch type from filename'):
        kv.sketch.load(golden_data('test.notasketchtype'))
    bad = tmp_path / 'bad.ct'
    bad.write_bytes(b'NOPE' + b'\0' * 40)
    with pytest.raises(OSError):
        kv.sketch.load(str(bad))
    with pytest.raises(OSError):
        kv.sketch.load(str(tmp_path / 'missing.ct'))
    with pytest.raises(OSError):   # a 1-bit file under an 8-bit extension
        kv.khmer.Counttable.load(golden_data('test.nodetable'))


def test_load_sketchfiles_fpr_gate(kv):
    sketches = kv.sketch.load_sketchfiles([golden_data('test.counttable')], maxfpr=0.5)
    assert sketches[0].get('CCTGATATCCGGAATCTTAGC') > 0
    with pytest.raises(kv.sketch.KevlarUnsuitableFPRError, match='FPR too high, bailing out!!!'):
        kv.sketch.load_sketchfiles([golden_data('test.counttable')], maxfpr=0.001)


# ------------------------------------------------------------------ consume vs oracle

@pytest.mark.parametrize('name', CLASSES)
@pytest.mark.parametrize('k', [13, 21, 31])
def test_consume_random_ragged(kv, oracle, name, k):
    """Ragged reads (shorter than k, exactly k, long), tiny tables so saturation and collisions
    are everywhere, reads crossing many 1024-base tiles."""
    reads = random_reads(k * 7 + len(name), 700, lo=5, hi=300)
    reads += [b'', b'A' * k, b'ACGT' * 700, b'T' * 5000, b'C' * (k - 1)]
    bases, offs = oracle.reads_to_batch(reads)
    g = getattr(kv.khmer, name)(k, 3000, 4)
    c = getattr(oracle, name)(k, 3000, 4)
    ng = g.consume_batch(bases, offs)
    nc = c.consume_batch(bases, offs)
    assert ng == nc
    assert_same_sketch(g, c)
    # second batch on top of the first: order-dependent n_unique must still agree
    reads2 = random_reads(99, 300, lo=k, hi=120)
    bases2, offs2 = oracle.reads_to_batch(reads2)
    assert g.consume_batch(bases2, offs2) == c.consume_batch(bases2, offs2)
    assert_same_sketch(g, c)


@pytest.mark.parametrize('name', ['Counttable', 'SmallCounttable', 'Countgraph'])
def test_consume_saturation(kv, oracle, name):
    """Thousands of copies of the same k-mers: every counter must stop exactly at 255 / 15,
    and the neighbouring counters in the same 32-bit word must be untouched."""
    reads = [b'ACGTTGCAAGGCTTAACCGGTTAAACCCGGGTTT' * 3] * 400 + random_reads(5, 50, 40, 60)
    bases, offs = oracle.reads_to_batch(reads)
    g = getattr(kv.khmer, name)(21, 500, 4)
    c = getattr(oracle, name)(21, 500, 4)
    redone = kv._lib.redo_count()
    assert g.consume_batch(bases, offs) == c.consume_batch(bases, offs)
    assert_same_sketch(g, c)
    top = 15 if name.startswith('Small') else 255
    assert g.get('ACGTTGCAAGGCTTAACCGGT') == top
    # > 1000 concurrent adds per bucket overflow the speculative path: the chunk must have been
    # rolled back and redone exactly (this is what makes the optimistic update safe).  The tiled path
    # (KV_UPDATE_PATH=tile) owns its regions and never speculates: there the same input overflows the
    # slabs of a few regions instead, which exercises the producer's in-place fallback.
    if os.environ.get('KV_UPDATE_PATH') != 'tile':
        assert kv._lib.redo_count() > redone


@pytest.mark.parametrize('name', ['Counttable', 'Nodegraph'])
def test_consume_cleaning(kv, oracle, name):
    """Lower case is upper-cased and anything outside ACGT counts as 'A' in the count path
    (unpinned by the reference, pinned here against the oracle)."""
    reads = random_reads(3, 200, 30, 90, alphabet=b'ACGTNacgtnRY-')
    bases, offs = oracle.reads_to_batch(reads)
    g = getattr(kv.khmer, name)(19, 5000, 4)
    c = getattr(oracle, name)(19, 5000, 4)
    assert g.consume_batch(bases, offs) == c.consume_batch(bases, offs)
    assert_same_sketch(g, c)


@pytest.mark.parametrize('name', ['Counttable', 'SmallCounttable', 'Nodetable', 'Countgraph'])
@pytest.mark.parametrize('num_bands,band', [(2, 0), (2, 1), (9, 2), (23, 19), (16, 15)])
def test_consume_banding(kv, oracle, name, num_bands, band):
    reads = random_reads(num_bands * 31 + band, 400, 30, 150)
    bases, offs = oracle.reads_to_batch(reads)
    g = getattr(kv.khmer, name)(19, 20000, 4)
    c = getattr(oracle, name)(19, 20000, 4)
    ng = g.consume_batch(bases, offs, num_bands=num_bands, band=band)
    nc = c.consume_batch(bases, offs, num_bands=num_bands, band=band)
    assert ng == nc
    assert_same_sketch(g, c)


def test_banding_partitions_the_kmers(kv, oracle):
    """Size-independent property: the bands partition the k-mer occurrences, so the per-band
    k-mer counts add up to the unbanded count."""
    reads = random_reads(11, 500, 50, 150)
    bases, offs = oracle.reads_to_batch(reads)
    full = kv.khmer.Counttable(25, 50000, 4).consume_batch(bases, offs)
    parts = [kv.khmer.Counttable(25, 50000, 4).consume_batch(bases, offs, num_bands=8, band=b) for b in range(8)]
    assert sum(parts) == full
    with pytest.raises(ValueError):
        kv.khmer.Counttable(25, 50000, 4).consume_batch(bases, offs, num_bands=4, band=4)


@pytest.mark.parametrize('name', ['Counttable', 'SmallCounttable', 'Nodetable'])
@pytest.mark.parametrize('consume_masked', [False, True])
@pytest.mark.parametrize('banded', [False, True])
def test_consume_with_mask(kv, oracle, name, consume_masked, banded):
    genome = LETTERS[np.random.default_rng(1).integers(0, 4, size=6000)]
    reads = random_reads(21, 400, 40, 120, genome=genome)
    maskreads = random_reads(22, 60, 40, 120, genome=genome)
    bases, offs = oracle.reads_to_batch(reads)
    mb, mo = oracle.reads_to_batch(maskreads)
    gm = kv.khmer.Nodetable(21, 1e4, 4)
    cm = oracle.Nodetable(21, 1e4, 4)
    gm.consume_batch(mb, mo)
    cm.consume_batch(mb, mo)
    g = getattr(kv.khmer, name)(21, 1e4, 4)
    c = getattr(oracle, name)(21, 1e4, 4)
    kw = dict(mask=None, threshold=1 if consume_masked else 0, consume_masked=consume_masked)
    if banded:
        kw.update(num_bands=3, band=1)
    ng = g.consume_batch(bases, offs, **dict(kw, mask=gm))
    nc = c.consume_batch(bases, offs, **dict(kw, mask=cm))
    assert ng == nc and ng > 0
    assert_same_sketch(g, c)


def test_count_cli_with_mask_golden(kv, tmp_path, capsys):
    """kevlar/tests/test_count.py:153-166: `36898 distinct k-mers stored`."""
    mask = kv.khmer.Nodetable(21, 1e4, 4)
    mask.consume('CACCAATCCGTACGGAGAGCCGTATATATAGACTGCTATACTATTGGATCGTACGGGGC')
    maskfile = str(tmp_path / 'mask.nt')
    mask.save(maskfile)
    arglist = ['count', '--ksize', '21', '--mask', maskfile, '--memory', '1M', str(tmp_path / 'out.sct'),
               golden_data('bogus-genome/refr.fa')]
    kv.count.main(kv.cli.parser().parse_args(arglist))
    assert '36898 distinct k-mers stored' in capsys.readouterr().err
    assert os.path.exists(str(tmp_path / 'out.sct.counttable'))


@pytest.mark.parametrize('count,smallcount,count_masked,kpresent,kabsent', [
    (True, True, True, 'CACCAATCCGTACGGAGAGCC', 'GAATCGGTGGCTGGTTGCCGT'),
    (True, False, True, 'CACCAATCCGTACGGAGAGCC', 'GAATCGGTGGCTGGTTGCCGT'),
    (False, False, True, 'CACCAATCCGTACGGAGAGCC', 'GAATCGGTGGCTGGTTGCCGT'),
    (True, True, False, 'GAATCGGTGGCTGGTTGCCGT', 'CACCAATCCGTACGGAGAGCC'),
    (True, False, False, 'GAATCGGTGGCTGGTTGCCGT', 'CACCAATCCGTACGGAGAGCC'),
    (False, False, False, 'GAATCGGTGGCTGGTTGCCGT', 'CACCAATCCGTACGGAGAGCC'),
])
def test_load_sample_seqfile_withmask(kv, count, smallcount, count_masked, kpresent, kabsent):
    """kevlar/tests/test_count.py:130-150."""
    mask = kv.khmer.Nodetable(21, 1e4, 4)
    mask.consume('CACCAATCCGTACGGAGAGCCGTATATATAGACTGCTATACTATTGGATCGTACGGGGC')
    sketch = kv.count.load_sample_seqfile([golden_data('bogus-genome/refr.fa')], 21, 1e6, mask=mask,
                                          consume_masked=count_masked, count=count, smallcount=smallcount)
    assert sketch.get(kpresent) > 0
    assert sketch.get(kabsent) == 0
    assert sketch.get('GATTACAGATTACAGATTACA') == 0


@pytest.mark.parametrize('cmd,name,outname', [
    (['--ksize', '21', '--memory', '200K', '-c', '4'], 'count_refr_small', 'o.sct'),
    (['--ksize', '21', '--memory', '100K', '-c', '1', '--num-bands', '3', '--band', '2'], 'count_refr_node_band', 'o.nt'),
])
def test_count_cli_generated_golden(kv, tmp_path, capsys, cmd, name, outname):
    """Outputs of the reference's own count.py over the oracle (tests/golden/make_golden.py)."""
    out = str(tmp_path / outname)
    kv.count.main(kv.cli.parser().parse_args(['count'] + cmd + [out, golden_data('bogus-genome/refr.fa')]))
    golden = golden_gen(name + ('.sct' if outname.endswith('.sct') else '.nt'))
    assert filecmp.cmp(out, golden, shallow=False)
    want = open(golden_gen(name + '.log')).read()
    got = capsys.readouterr().err
    for line in re.findall(r'\d+ reads processed, \d+ distinct k-mers stored', want):
        assert line in got


def test_count_problematic(kv):
    """kevlar/tests/test_count.py:83-99."""
    args = kv.cli.parser().parse_args(['count', '--ksize', '21', '--memory', '200K', '--band', '2', 'bogusoutput',
                                       golden_data('trio1/ctrl1.fq.gz')])
    with pytest.raises(ValueError, match=r'Must specify --num-bands and --band together'):
        kv.count.main(args)
    args = kv.cli.parser().parse_args(['count', '--ksize', '21', '--memory', '97', 'bogusoutput',
                                       golden_data('trio1/ctrl1.fq.gz')])
    with pytest.raises(kv.sketch.KevlarUnsuitableFPRError):
        kv.count.main(args)


@pytest.mark.parametrize('mask,numbands,band', [(False, None, None), (False, 9, 2), (True, None, None), (True, 23, 19)])
def test_load_threading(kv, oracle, mask, numbands, band):
    """kevlar/tests/test_count.py:31-42, and the 2-thread result must equal the oracle's
    (saturating increments commute)."""
    def build(mod):
        m = None
        if mask:
            m = mod.Counttable(19, 1e4, 4)
            m.consume('TGAGGGGACTAGGTGATCAGGTGAGGGTTTCCCAGTTCCCGAAGATGACT')
            m.consume('GATCTTTCGCTCCCTGTCATCAAGGAGTGATACGCGAAGTGCGTCCCCTT')
        return m
    gpu = kv.count.load_sample_seqfile([golden_data('trio1/case1.fq.gz')], 19, 1e7, mask=build(kv.khmer),
                                       numbands=numbands, band=band, numthreads=2)
    cpu = oracle.Counttable(19, 1e7 / 4, 4)
    kw = {}
    parser = oracle.ReadParser(golden_data('trio1/case1.fq.gz'))
    m = build(oracle)
    if m is not None and numbands:
        cpu.consume_seqfile_banding_with_mask(parser, numbands, band, m, threshold=0, consume_masked=False)
    elif m is not None:
        cpu.consume_seqfile_with_mask(parser, m, threshold=0, consume_masked=False)
    elif numbands:
        cpu.consume_seqfile_banding(parser, numbands, band)
    else:
        cpu.consume_seqfile(parser)
    assert_same_sketch(gpu, cpu, check_unique=False)


def test_memory_sizing(kv):
    """kevlar/tests/test_count.py:169-182."""
    for count, smallcount, kind in [(False, False, 'nodegraph'), (True, False, 'countgraph'), (True, True, 'smallcountgraph')]:
        sketch = kv.count.load_sample_seqfile([golden_data('bogus-genome/refr.fa')], 21, 2e6, count=count,
                                              smallcount=smallcount)
        actual = sum(sketch.hashsizes()) / kv.khmer._buckets_per_byte[kind]
        assert actual / 2e6 == pytest.approx(1.0, rel=1e-4)


def test_chunked_consume_equals_unchunked(kv, oracle, monkeypatch):
    """A batch larger than the scratch chunk is processed in several chunks: same result."""
    reads = random_reads(8, 3000, 80, 120)
    bases, offs = oracle.reads_to_batch(reads)
    c = oracle.Counttable(31, 1e5, 4)
    c.consume_batch(bases, offs)
    g = kv.khmer.Counttable(31, 1e5, 4)
    half = len(reads) // 2
    cut = int(offs[half])
    g.consume_batch(bases[:cut], offs[:half + 1])
    g.consume_batch(bases[cut:], offs[half:] - offs[half])
    assert_same_sketch(g, c)


# ------------------------------------------------------------------ add / get lists

@pytest.mark.parametrize('name', CLASSES)
def test_add_get_lists(kv, oracle, name):
    rng = np.random.default_rng(12)
    kmers = [LETTERS[rng.integers(0, 4, size=21)].tobytes().decode() for _ in range(400)]
    kmers = kmers + kmers[:150] + kmers[:40] * 20
    g = getattr(kv.khmer, name)(21, 700, 4)
    c = getattr(oracle, name)(21, 700, 4)
    g.add_many(kmers)
    for km in kmers:
        c.add(km)
    assert_same_sketch(g, c)
    assert list(g.get_many(kmers[:400])) == [c.get(km) for km in kmers[:400]]
    g.add(kmers[0])
    c.add(kmers[0])
    assert g.get(kmers[0]) == c.get(kmers[0])
    assert g.get(g.hash(kmers[1])) == c.get(kmers[1])


# ------------------------------------------------------------------ novel

def _novel_inputs(oracle, kv, seed=5, k=25, mem=2e5):
    rng = np.random.default_rng(seed)
    genome = LETTERS[rng.integers(0, 4, size=8000)]
    child = genome.copy()
    for pos in (2000, 5000, 5003):
        child[pos] = LETTERS[(np.where(LETTERS == child[pos])[0][0] + 1) % 4]
    samples = [random_reads(seed + 1, 2500, 60, 110, genome=child),
               random_reads(seed + 2, 2500, 60, 110, genome=genome),
               random_reads(seed + 3, 2500, 60, 110, genome=genome)]
    samples[0] += [b'ACGTNACGT' * 10, b'acgt' * 20, b'AC', b'']
    gpu, cpu = [], []
    for seqs in samples:
        bases, offs = oracle.reads_to_batch(seqs)
        g, c = kv.khmer.Counttable(k, mem / 4, 4), oracle.Counttable(k, mem / 4, 4)
        g.consume_batch(bases, offs)
        c.consume_batch(bases, offs)
        gpu.append(g)
        cpu.append(c)
    return samples, gpu, cpu


@pytest.mark.parametrize('case_min,ctrl_max,screen,numbands,bm1', [
    (6, 1, None, None, 0), (5, 0, None, None, 0), (8, 1, 3, None, 0), (6, 1, 6, None, 0),
    (6, 1, None, 2, 0), (6, 1, None, 4, 2), (6, 1, None, 8, -1), (3, 2, 2, 4, 1),
])
def test_novel_batch_matches_oracle(kv, oracle, case_min, ctrl_max, screen, numbands, bm1):
    samples, gpu, cpu = _novel_inputs(oracle, kv)
    bases, offs = oracle.reads_to_batch(samples[0])
    hits, flags, _ = kv.khmer.novel_batch(gpu[:1], gpu[1:], bases, offs, case_min, ctrl_max, screen=screen,
                                          num_bands=numbands, band_minus_1=bm1)
    ohits, oflags = oracle.novel_batch(cpu[:1], cpu[1:], bases, offs, case_min, ctrl_max, screen=screen,
                                       numbands=numbands, band_minus_1=bm1)
    assert len(hits) == len(ohits)
    assert (hits['read'] == ohits['read']).all()
    assert (hits['offset'] == ohits['offset']).all()
    assert (hits['abund'][:, :3] == ohits['abund'][:, :3]).all()
    assert (flags == oflags).all()
    if numbands is None and screen is None:
        assert len(hits) > 0


def test_novel_two_cases_mixed_types(kv, oracle):
    """Two case sketches and controls of different counter widths in one scan."""
    samples, gpu, cpu = _novel_inputs(oracle, kv, seed=9, k=21)
    bases, offs = oracle.reads_to_batch(samples[0])
    g2, c2 = kv.khmer.SmallCounttable(21, 40000, 4), oracle.SmallCounttable(21, 40000, 4)
    g3, c3 = kv.khmer.Nodetable(21, 90000, 4), oracle.Nodetable(21, 90000, 4)
    for s, (gx, cx) in zip((samples[0], samples[1]), ((g2, c2), (g3, c3))):
        b, o = oracle.reads_to_batch(s)
        gx.consume_batch(b, o)
        cx.consume_batch(b, o)
    hits, flags, _ = kv.khmer.novel_batch([gpu[0], g2], [gpu[1], gpu[2], g3], bases, offs, 5, 1)
    ohits, oflags = oracle.novel_batch([cpu[0], c2], [cpu[1], cpu[2], c3], bases, offs, 5, 1)
    assert len(hits) == len(ohits)
    assert (hits['read'] == ohits['read']).all() and (hits['offset'] == ohits['offset']).all()
    assert (hits['abund'][:, :5] == ohits['abund'][:, :5]).all()
    assert (flags == oflags).all()


def _run_cli(kv, arglist):
    args = kv.cli.parser().parse_args(arglist)
    log, out = io.StringIO(), io.StringIO()
    saved = kv.logstream
    kv.logstream = log
    import contextlib
    try:
        with contextlib.redirect_stdout(out):
            kv.cli.mains[args.cmd](args)
    finally:
        kv.logstream = saved
    logtext = re.sub(r'\d+\.\d\d sec(onds)?', 'T sec', log.getvalue())
    return out.getvalue(), logtext.replace(golden_data(''), 'DATA/')


def _same_log(got, want):
    """Log lines equal up to the input file suffix (trio1 fixtures are stored gzipped)."""
    got = got.replace('.fq.gz"', '.fq"')
    want = want.replace('.fq.gz"', '.fq"')
    assert got == want


NA = ['microtrios/trio-na-{}.fq.gz'.format(w) for w in ('proband', 'mother', 'father')]


def test_novel_cli_microtrio_golden(kv):
    """End to end: byte-identical to the reference's novel.py over the oracle, which in turn
    equals the novel output shipped with the reference (microtrios/novel-na.augfastq.gz)."""
    out, log = _run_cli(kv, ['novel', '-k', '31', '--case-min', '5', '--ctrl-max', '1', '--memory', '500K',
                             '--case', golden_data(NA[0]), '--control', golden_data(NA[1]),
                             '--control', golden_data(NA[2])])
    assert out == open(golden_gen('novel_microtrio_na.out')).read()
    _same_log(log, open(golden_gen('novel_microtrio_na.log')).read())
    shipped = gzip.open(golden_data('microtrios/novel-na.augfastq.gz'), 'rt').read()
    shipped = ''.join(line for line in shipped.splitlines(True) if not line.startswith('#mateseq='))
    assert out == shipped


@pytest.mark.parametrize('nb,b,name', [('2', '2', 'novel_microtrio_na_band2of2'), ('4', '3', 'novel_microtrio_na_band3of4')])
def test_novel_cli_banded_golden(kv, nb, b, name):
    """kevlar/tests/test_novel.py:80-105 configuration; checks the band quirk end to end."""
    out, log = _run_cli(kv, ['novel', '--case', golden_data(NA[0]), '--ksize', '25', '--case-min', '7',
                             '--control', golden_data(NA[2]), '--control', golden_data(NA[1]),
                             '--num-bands', nb, '--band', b, '--ctrl-max', '0', '--memory', '500K'])
    assert out == open(golden_gen(name + '.out')).read()
    _same_log(log, open(golden_gen(name + '.log')).read())
    for line in out.split('\n'):
        if line.endswith('#'):
            m = re.search(r'(\d+) (\d+) (\d+)#$', line)
            assert int(m.group(1)) >= 7 and m.group(2) == '0' and m.group(3) == '0'


@pytest.mark.parametrize('skip', [True, False])
def test_novel_cli_trio1_golden(kv, skip):
    """kevlar/tests/test_novel.py:179-207: '(skipped 1001 reads)', '29 unique novel kmers in 14 reads'."""
    cmd = ['novel', '--ctrl-max', '0', '--case-min', '6', '--case', golden_data('trio1/case1.fq.gz'),
           '--control', golden_data('trio1/ctrl1.fq.gz'), '--control', golden_data('trio1/ctrl2.fq.gz')]
    name = 'novel_trio1'
    if skip:
        cmd += ['--skip-until', 'bogus-genome-chr1_115_449_0:0:0_0:0:0_1f4/1']
        name += '_skipuntil'
    out, log = _run_cli(kv, cmd)
    assert out == open(golden_gen(name + '.out')).read()
    _same_log(log, open(golden_gen(name + '.log')).read())
    if skip:
        assert 'Found read bogus-genome-chr1_115_449_0:0:0_0:0:0_1f4/1 (skipped 1001 reads)' in log
        assert '29 unique novel kmers in 14 reads' in log


def test_novel_skip_until_missing(kv):
    out, log = _run_cli(kv, ['novel', '--ctrl-max', '0', '--case-min', '6', '--case', golden_data('trio1/case1.fq.gz'),
                             '--control', golden_data('trio1/ctrl1.fq.gz'), '--control',
                             golden_data('trio1/ctrl2.fq.gz'), '--skip-until', 'BOGUSREADNAME'])
    assert 'Found read' not in log and '(skipped ' not in log
    assert 'Found 0 instances of 0 unique novel kmers in 0 reads' in log
    assert out == ''


@pytest.mark.parametrize('screen,name', [(['--abund-screen', '3'], 'novel_abund_screen'), ([], 'novel_no_abund_screen')])
def test_novel_cli_abund_screen_golden(kv, screen, name):
    """kevlar/tests/test_novel.py:167-176."""
    out, log = _run_cli(kv, ['novel', '--ksize', '25', '--ctrl-max', '1', '--case-min', '8',
                             '--case', golden_data('screen-case.fa'), '--control', golden_data('screen-ctrl.fa')] + screen)
    assert out == open(golden_gen(name + '.out')).read()
    _same_log(log, open(golden_gen(name + '.log')).read())
    if screen:
        assert '>seq_error' not in out


def test_novel_cli_load_counts_golden(kv):
    """kevlar/tests/test_novel.py:268-282: pre-computed sketches + a read with an N."""
    out, log = _run_cli(kv, ['novel', '-k', '25', '--case', golden_data('simple-genome-case-reads.fa.gz'),
                             golden_data('ambig.fasta'), '--case-counts', golden_data('simple-genome-case.ct'),
                             '--control-counts', golden_data('simple-genome-ctrl1.ct'),
                             golden_data('simple-genome-ctrl2.ct')])
    assert 'counttables for 2 sample(s) provided' in log
    assert out == open(golden_gen('novel_load_counts.out')).read()
    _same_log(log, open(golden_gen('novel_load_counts.log')).read())


def test_novel_save_counts(kv, tmp_path):
    """kevlar/tests/test_novel.py:210-241: --save-*-counts files equal `kevlar count` outputs."""
    d = str(tmp_path)
    for ind, src in zip(('proband', 'mother', 'father'), NA):
        kv.count.main(kv.cli.parser().parse_args(['count', '--ksize', '27', '--memory', '500K',
                                                  '{}/{}.ct'.format(d, ind), golden_data(src)]))
    kv.novel.main(kv.cli.parser().parse_args([
        'novel', '--ksize', '27', '--out', d + '/novel.augfastq.gz', '--save-case-counts', d + '/kid.ct',
        '--save-ctrl-counts', d + '/mom.ct', d + '/dad.ct', '--case', golden_data(NA[0]),
        '--control', golden_data(NA[1]), '--control', golden_data(NA[2]), '--memory', '500K']))
    for a, b in (('father', 'dad'), ('mother', 'mom'), ('proband', 'kid')):
        assert filecmp.cmp('{}/{}.ct'.format(d, a), '{}/{}.ct'.format(d, b), shallow=False)
    import hashlib
    import json
    manifest = json.load(open(os.path.join(os.path.dirname(golden_gen('x')), '..', 'MANIFEST.json')))
    digest = hashlib.sha256(open(d + '/proband.ct', 'rb').read()).hexdigest()
    assert digest == manifest['gen/count_na_proband.ct']['sha256']


def test_novel_api_errors(kv):
    """kevlar/tests/test_novel.py:25-37."""
    with pytest.raises(ValueError, match=r'Must specify `numbands` and `band` together'):
        list(kv.novel.novel(None, [], [], numbands=4))
    with pytest.raises(ValueError, match=r'Must specify `numbands` and `band` together'):
        list(kv.novel.novel(None, [], [], band=0))
    with pytest.raises(ValueError, match=r'`band` must be a value between 0 and 3'):
        list(kv.novel.novel(None, [], [], numbands=4, band=-1))


def test_novel_generic_record_stream(kv, oracle):
    """`novel` accepts any iterable of records (the reference's Python-API use)."""
    samples, gpu, cpu = _novel_inputs(oracle, kv, seed=17)
    stream = [kv.sequence.Record('r{}'.format(i), s.decode(), 'I' * len(s)) for i, s in enumerate(samples[0])]
    recs = list(kv.novel.novel(stream, gpu[:1], gpu[1:], ksize=25, casemin=6, ctrlmax=1))
    bases, offs = oracle.reads_to_batch(samples[0])
    ohits, _ = oracle.novel_batch(cpu[:1], cpu[1:], bases, offs, 6, 1)
    assert sum(len(r.annotations) for r in recs) == len(ohits)
    assert [r.name for r in recs] == ['r{}'.format(i) for i in sorted(set(ohits['read'].tolist()))]
    for ikmer in recs[0].annotations:
        assert ikmer.abund[0] >= 6 and max(ikmer.abund[1:]) <= 1


def test_kmer_is_interesting(kv, oracle):
    samples, gpu, cpu = _novel_inputs(oracle, kv, seed=23)
    bases, offs = oracle.reads_to_batch(samples[0])
    ohits, _ = oracle.novel_batch(cpu[:1], cpu[1:], bases, offs, 6, 1)
    h = ohits[0]
    seq = samples[0][int(h['read'])].decode()
    kmer = seq[int(h['offset']):int(h['offset']) + 25]
    ok, discard, ca, co = kv.novel.kmer_is_interesting(kmer, gpu[:1], gpu[1:], case_min=6, ctrl_max=1)
    assert ok and not discard and ca == [int(h['abund'][0])] and co == [int(h['abund'][1]), int(h['abund'][2])]
    ok, discard, ca, co = kv.novel.kmer_is_interesting('GATTACA' * 3 + 'GATT', gpu[:1], gpu[1:], case_min=6,
                                                       ctrl_max=1, screen_thresh=2)
    assert not ok and discard and ca == [] and co == []


# ------------------------------------------------------------------ filter

def _filter_text(kv, readfile, **kw):
    out = io.StringIO()
    saved, kv.logstream = kv.logstream, io.StringIO()
    try:
        recs = list(kv.filter.filter(readfile, **kw))
    finally:
        kv.logstream = saved
    for rec in recs:
        kv.print_augmented_fastx(rec, out)
    return recs, out.getvalue()


def test_filter_alpha_golden(kv):
    """kevlar/tests/test_filter.py:27-41 (memory=500: collisions everywhere)."""
    recs, text = _filter_text(kv, golden_data('collect.alpha.txt'), memory=500)
    assert len(recs) == 8
    assert text == open(golden_gen('filter_alpha.out')).read()


def test_filter_worm_golden(kv):
    """kevlar/tests/test_filter.py:61-73."""
    recs, text = _filter_text(kv, golden_data('worm.augfasta'), memory=1000, casemin=5, ctrlmax=0)
    assert len(recs) == 5
    assert text == open(golden_gen('filter_worm.out')).read()


@pytest.mark.parametrize('maskkind,nkmers,ninst', [(None, 424, 5782), ('refr', 424, 5782), ('mask.nt', 13, 171)])
def test_filter_ctrl3(kv, maskkind, nkmers, ninst):
    """kevlar/tests/test_filter.py:44-58."""
    mask = None
    if maskkind == 'refr':
        mask = kv.count.load_sample_seqfile([golden_data('bogus-genome/refr.fa')], 13, 1e7, count=False)
    elif maskkind:
        mask = kv.sketch.load(golden_data('bogus-genome/mask.nt'))
    recs, _ = _filter_text(kv, golden_data('trio1/novel_3_1,2.txt'), memory=1e7, mask=mask)
    seen = {}
    for read in recs:
        for ikmer in read.annotations:
            key = kv.revcommin(read.ikmerseq(ikmer))
            seen[key] = seen.get(key, 0) + 1
    assert len(seen) == nkmers and sum(seen.values()) == ninst


def test_filter_main_golden(kv):
    """kevlar/tests/test_filter.py:76-87."""
    out, log = _run_cli(kv, ['filter', '--mask', golden_data('bogus-genome/mask.nt'), '--memory', '10M',
                             '--max-fpr', '0.001', '--case-min', '6', golden_data('trio1/novel_3_1,2.txt')])
    assert 'Processed 178 reads' in log and 'Validated 18 reads' in log
    assert out == open(golden_gen('filter_trio1_mask.out')).read()
    assert log == open(golden_gen('filter_trio1_mask.log')).read()


# ------------------------------------------------------------------ larger, property-based

def test_full_size_properties(kv, oracle):
    """BASELINE config 2 shape (300k reads x 100 bp, k=31, 64 MB sketch): too slow for the
    Python side of the oracle, so checked through size-independent properties plus a
    multi-threaded oracle run (saturating counts are order-independent)."""
    rng = np.random.default_rng(42)
    genome = LETTERS[rng.integers(0, 4, size=1000000)]
    n = 300000
    starts = rng.integers(0, len(genome) - 100, size=n)
    idx = starts[:, None] + np.arange(100)[None, :]
    bases = genome[idx].reshape(-1).copy()
    err = rng.random(len(bases)) < 0.005
    bases[err] = LETTERS[rng.integers(0, 4, size=int(err.sum()))]
    offs = (np.arange(n + 1) * 100).astype(np.uint64)
    g = kv.khmer.Counttable(31, 64e6 / 4, 4)
    nk = g.consume_batch(bases, offs)
    assert nk == n * 70
    c = oracle.Counttable(31, 64e6 / 4, 4)
    assert c.consume_batch(bases, offs, threads=8) == nk
    assert_same_sketch(g, c, check_unique=False)
    # linearity: consuming the batch in two halves into a fresh sketch gives the same tables
    g2 = kv.khmer.Counttable(31, 64e6 / 4, 4)
    half = n // 2
    g2.consume_batch(bases[:half * 100], offs[:half + 1])
    g2.consume_batch(bases[half * 100:], offs[half:] - offs[half])
    for t in range(4):
        assert g2.table_bytes(t) == g.table_bytes(t)
    assert g2.n_unique_kmers() == g.n_unique_kmers()
    # every k-mer of a read just counted is present at least once
    counts = g.get_kmer_counts(bases[:100].tobytes().decode())
    assert min(counts) >= 1


def _rerun_in_child(selection, **env_overrides):
    """Library tunables are read once per process, so variants run in a child pytest."""
    import subprocess
    import sys
    if os.environ.get('KV_TEST_CHILD'):
        pytest.skip('already inside a child run')
    env = dict(os.environ, KV_TEST_CHILD='1', **env_overrides)
    res = subprocess.run([sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-m', 'gpu', '-x', '-q', '-k', selection],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200)
    assert res.returncode == 0, res.stdout[-3000:]
    assert ' passed' in res.stdout


def test_partitioned_update_path_is_exact():
    """Sketches larger than L2 use the region-partitioned update kernels (hist / scan / scatter /
    apply).  The path is chosen from the sketch size, so a child process with the threshold
    forced to zero and tiny regions re-runs the count parity tests through it."""
    _rerun_in_child('consume or saturation or add_get or count_simple or full_size or abundance_distribution',
                    KV_PART_MIN_BYTES='0', KV_PART_REGION_LOG2='10')


def test_tiled_update_path_is_exact():
    """Sketches that do not fit L2 are updated region by region in shared memory (K3c: offsets filed by
    the hash kernel, one CTA per region).  The path is chosen from the sketch size, so child processes
    force it onto the parity tests: with 256-bucket regions (every table spans many regions, ragged last
    region), with 1024-bucket regions plus 4096-position chunks and the sparse-region shortcut switched
    off (slab overflow, many chunks, n_unique across chunk boundaries), and with the shortcut for all."""
    tests = 'consume or add_get or count_simple or full_size or abundance_distribution or load_threading or count_cli'
    _rerun_in_child(tests, KV_UPDATE_PATH='tile', KV_TILE_RB='8')
    _rerun_in_child(tests, KV_UPDATE_PATH='tile', KV_TILE_RB='10', KV_TILE_CHUNK_BASES='4096', KV_TILE_DIRECT_BELOW='0')
    _rerun_in_child('consume', KV_UPDATE_PATH='tile', KV_TILE_DIRECT_BELOW='1000000000')


def test_tiled_update_default_geometry(kv, oracle):
    """The tiled path with its default tunables at a size that takes it naturally (256 MB sketch, 32768-
    bucket regions): sketch bytes and n_unique_kmers equal the oracle's (8 threads; saturating counts are
    order-independent), for 8-bit and 4-bit counters; a repeated read overflows some slabs."""
    rng = np.random.default_rng(11)
    genome = LETTERS[rng.integers(0, 4, size=2000000)]
    n = 200000
    starts = rng.integers(0, len(genome) - 100, size=n)
    bases = genome[starts[:, None] + np.arange(100)[None, :]].reshape(-1).copy()
    bases[:100 * 3000] = np.tile(bases[:100], 3000)      # 3000 copies of one read: counters saturate, slabs overflow
    offs = (np.arange(n + 1) * 100).astype(np.uint64)
    for name, mem in (('Counttable', 256e6), ('SmallCounttable', 160e6)):
        buckets = mem / 4 * (2 if name.startswith('Small') else 1)
        g = getattr(kv.khmer, name)(31, buckets, 4)
        c = getattr(oracle, name)(31, buckets, 4)
        assert g.consume_batch(bases, offs) == n * 70
        c.consume_batch(bases, offs, threads=8)
        for t in range(4):
            assert g.table_bytes(t) == c.table_bytes(t), (name, t)
        assert g.n_occupied() == c.n_occupied()
        g2 = getattr(kv.khmer, name)(31, buckets, 4)      # two halves == one batch, n_unique included
        half = n // 2
        g2.consume_batch(bases[:half * 100], offs[:half + 1])
        g2.consume_batch(bases[half * 100:], offs[half:] - offs[half])
        for t in range(4):
            assert g2.table_bytes(t) == g.table_bytes(t)
        assert g2.n_unique_kmers() == g.n_unique_kmers()
        del g, g2, c


def test_multi_chunk_batches_are_exact():
    """A batch larger than the per-chunk scratch is hashed and applied chunk by chunk; the
    order-dependent n_unique_kmers must survive the chunk boundaries.  Child process with a
    2048-position chunk (a few reads per chunk)."""
    # (the saturation test is left out: it asserts that the overflow redo fires, which needs big chunks)
    _rerun_in_child('(consume and not saturation) or count_simple or novel_cli_microtrio or count_cli_with_mask '
                    'or abundance_distribution or dist_passes or unique_across_shards',
                    KV_CHUNK_BASES='2048')


def test_ranged_first_touch_passes_are_exact():
    """Tables with more buckets than the n_unique scratch covers are processed in bucket ranges.
    Child process with a 1024-bucket range: every table of the parity tests spans several ranges,
    and n_unique_kmers / the abundance distribution must not change."""
    _rerun_in_child('(consume and not saturation) or count_simple or abundance_distribution or dist_passes',
                    KV_FIRST_RANGE_LOG2='10')


def test_unfused_first_touch_passes_are_exact():
    """n_unique_kmers without the fused table-0 pass / compact list (the route taken by add(), by the
    abundance distribution and by tables larger than the scratch): same numbers."""
    _rerun_in_child('(consume and not saturation) or count_simple or full_size', KV_NO_CLASSIFY='1')


def test_khmer_namespace_drop_in(kv, tmp_path):
    """INTEGRATION.md route 1: code written against the `khmer` namespace the way the reference
    uses it -- threads sharing one ReadParser into consume_seqfile, then one get() per k-mer per
    sample (kevlar/count.py:40-77, kevlar/novel.py:134-162) -- runs on kevlar_b200.khmer and
    reproduces the golden sketch and the golden novel output."""
    import sys
    import threading
    saved = {name: sys.modules.get(name) for name in ('khmer', 'khmer.khmer_args')}
    sys.modules['khmer'] = kv.khmer
    sys.modules['khmer.khmer_args'] = kv.khmer.khmer_args
    try:
        import khmer
        from khmer import khmer_args
        assert khmer_args.memory_setting('10K') == 1e4
        # --- count, reference style
        tablesize = khmer_args.memory_setting('10K') / 4 * khmer._buckets_per_byte['countgraph']
        sketch = khmer.Counttable(25, tablesize, 4)
        parser = khmer.ReadParser(golden_data('simple-genome-case-reads.fa.gz'))
        workers = [threading.Thread(target=sketch.consume_seqfile, args=(parser,)) for _ in range(2)]
        [w.start() for w in workers]
        [w.join() for w in workers]
        assert parser.num_reads == 600
        out = str(tmp_path / 'case.ct')
        sketch.save(out)
        assert filecmp.cmp(out, golden_data('simple-genome-case.ct'), shallow=False)
        fpr = (sketch.n_occupied() / min(sketch.hashsizes())) ** len(sketch.hashsizes())
        assert '{:1.3f}'.format(fpr) == '0.011'
        # --- novel, reference style: per read, per k-mer, per sample
        cases = [khmer.Counttable.load(golden_data('simple-genome-case.ct'))]
        ctrls = [khmer.Counttable.load(golden_data('simple-genome-ctrl{}.ct'.format(i))) for i in (1, 2)]
        lines = []
        for fn in ('simple-genome-case-reads.fa.gz', 'ambig.fasta'):
            for record in khmer.ReadParser(golden_data(fn)):
                if len(record.sequence) < 25 or re.search('[^ACGT]', record.sequence):
                    continue
                annots = []
                for i, kmer in enumerate(cases[0].get_kmers(record.sequence)):
                    abunds = [ct.get(kmer) for ct in cases]
                    if min(abunds) < 6:
                        continue
                    cab = [ct.get(kmer) for ct in ctrls]
                    if max(cab) > 1:
                        continue
                    annots.append(' ' * i + kmer + ' ' * 10 + ' '.join(str(a) for a in abunds + cab) + '#')
                if annots:
                    lines += ['>' + record.name, record.sequence] + annots
        assert '\n'.join(lines) + '\n' == open(golden_gen('novel_load_counts.out')).read()
    finally:
        for name, mod in saved.items():
            if mod is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = mod


def test_device_resident_batches(kv, oracle):
    """KV_MEM_DEVICE inputs (reads already in HBM, e.g. torch tensors): same sketches and hits as
    the host-buffer path and the oracle; ragged tail not a multiple of 4 bytes."""
    torch = pytest.importorskip('torch')
    samples, gpu_host, cpu = _novel_inputs(oracle, kv, seed=31)
    dev = torch.device('cuda', kv._lib.current_device())
    gpu_dev = []
    for seqs, ref in zip(samples, cpu):
        bases, offs = oracle.reads_to_batch(seqs)
        assert len(bases) % 4 != 0 or True
        b = torch.from_numpy(bases).to(dev)
        o = torch.from_numpy(offs.view(np.int64)).to(dev)
        torch.cuda.synchronize()
        g = kv.khmer.Counttable(25, 2e5 / 4, 4)
        n = g.consume_batch(b.data_ptr(), (o.data_ptr(), o.numel() - 1), where=kv.khmer.MEM_DEVICE)
        assert n == sum(max(0, len(s) - 24) for s in seqs)
        assert_same_sketch(g, ref)
        gpu_dev.append(g)
    bases, offs = oracle.reads_to_batch(samples[0])
    b = torch.from_numpy(bases).to(dev)
    o = torch.from_numpy(offs.view(np.int64)).to(dev)
    torch.cuda.synchronize()
    hits, flags, _ = kv.khmer.novel_batch(gpu_dev[:1], gpu_dev[1:], b.data_ptr(), (o.data_ptr(), o.numel() - 1, b.numel()),
                                          6, 1, where=kv.khmer.MEM_DEVICE)
    ohits, oflags = oracle.novel_batch(cpu[:1], cpu[1:], bases, offs, 6, 1)
    assert len(hits) == len(ohits) > 0
    assert (hits['read'] == ohits['read']).all() and (hits['offset'] == ohits['offset']).all()
    assert (hits['abund'][:, :3] == ohits['abund'][:, :3]).all()
    # non-ACGT reads and reads shorter than k are flagged on the device, whatever memory the batch is in
    assert (np.diff(offs.astype(np.int64)) < 25).any()
    assert (flags == oflags).all()


def test_get_kmer_counts_many(kv, oracle):
    """Batched form of get_kmer_counts (kevlar/simlike.py:23-29 queries a sketch per call window)."""
    seqs = [s.decode() for s in random_reads(77, 40, 30, 200)]
    g, c = kv.khmer.Counttable(21, 5e4, 4), oracle.Counttable(21, 5e4, 4)
    bases, offs = oracle.reads_to_batch(seqs[:25])
    g.consume_batch(bases, offs)
    c.consume_batch(bases, offs)
    got = g.get_kmer_counts_many(seqs)
    assert len(got) == len(seqs)
    for seq, counts in zip(seqs, got):
        assert list(counts) == c.get_kmer_counts(seq)
    with pytest.raises(ValueError):
        g.get_kmer_counts_many(['ACGT' * 10, 'ACGTN' * 10])


# ------------------------------------------------------------------ kevlar dist

DIST_ABUND_10K = {10: 6, 11: 10, 12: 12, 13: 18, 14: 16, 15: 11, 16: 9, 17: 9, 18: 11, 19: 8, 20: 9, 21: 7, 22: 3}


def test_dist_passes_golden(kv, tmp_path):
    """kevlar/tests/test_dist.py:25-43: the masked first pass byte-for-byte, then the second pass's
    abundance dictionary."""
    from kevlar_b200.dist import count_first_pass, count_second_pass
    mask = kv.khmer.Nodetable.load(golden_data('minitrio/mask.nt'))
    counts = kv.khmer.Counttable(31, 1e4, 4)
    reads = golden_data('minitrio/trio-proband.fq.gz')
    count_first_pass([reads], counts, mask)
    out = str(tmp_path / 'first.ct')
    counts.save(out)
    assert filecmp.cmp(out, golden_data('minitrio/trio-proband-mask-counts.ct'), shallow=False)
    loaded = kv.khmer.Counttable.load(golden_data('minitrio/trio-proband-mask-counts.ct'))
    assert count_second_pass([reads], loaded) == DIST_ABUND_10K


def test_dist_cli_golden(kv, tmp_path):
    """kevlar/tests/test_dist.py:69-122: mu/sigma on stdout; the TSV equals the file shipped with
    the reference (minitrio/trio-proband-dist.tsv) byte for byte."""
    import json
    tsv = str(tmp_path / 'dist.tsv')
    out, _ = _run_cli(kv, ['dist', '--tsv', tsv, golden_data('minitrio/mask.nt'),
                           golden_data('minitrio/trio-proband.fq.gz')])
    js = json.loads(out)
    assert open(tsv).read() == open(golden_data('minitrio/trio-proband-dist.tsv')).read()
    from kevlar_b200.dist import dist
    mask = kv.khmer.Nodetable.load(golden_data('minitrio/mask.nt'))
    mu, sigma, data = dist([golden_data('minitrio/trio-proband.fq.gz')], mask, memory=4e4)
    assert mu == pytest.approx(15.32558, abs=1e-5) and sigma == pytest.approx(3.280581, abs=1e-5)
    assert list(data['Count'][-5:]) == [11.0, 8.0, 9.0, 7.0, 3.0]
    assert js['mu'] > 0 and js['sigma'] > 0


def test_dist_empty(kv):
    """kevlar/tests/test_dist.py:78-85."""
    from kevlar_b200.dist import dist, KevlarZeroAbundanceDistError
    mask = kv.khmer.Nodetable(31, 1e4, 4)
    mask.consume('GATTACA' * 10)
    mask.consume('A' * 50)
    with pytest.raises(KevlarZeroAbundanceDistError):
        dist([golden_data('minitrio/trio-proband.fq.gz')], mask, memory=4e4)


@pytest.mark.parametrize('counts_cls,track_cls', [('Counttable', 'Nodetable'), ('SmallCounttable', 'Nodetable'),
                                                  ('Countgraph', 'Nodegraph'), ('Counttable', 'Counttable')])
def test_abundance_distribution_matches_oracle(kv, oracle, counts_cls, track_cls):
    """Random ragged reads with repeats, heavy table collisions (tiny tracking tables), dirty bases
    and several batches sharing one tracking sketch: histogram and tracking tables bit-exact."""
    k = 21
    genome = LETTERS[np.random.default_rng(3).integers(0, 4, size=4000)]
    reads = random_reads(11, 900, lo=10, hi=150, genome=genome) + random_reads(12, 60, alphabet=b'ACGTNacgt')
    from oracle.khmer_oracle import reads_to_batch
    g_counts, c_counts = getattr(kv.khmer, counts_cls)(k, 3000, 4), getattr(oracle, counts_cls)(k, 3000, 4)
    bases, offs = reads_to_batch(reads)
    g_counts.consume_batch(bases, offs)
    c_counts.consume_batch(bases, offs)
    g_track = getattr(kv.khmer, track_cls)(k, 1, 1, primes=[1009, 997, 991])
    c_track = getattr(oracle, track_cls)(k, 1, 1, primes=[1009, 997, 991])
    for part in (reads[:300], reads[300:301], [], reads[301:]):
        bases, offs = reads_to_batch(part)
        want = np.zeros(256, dtype=np.uint64)
        if part:
            c_counts.abundance_distribution_batch(bases, offs, c_track, want)
        got = g_counts.abundance_distribution_batch(bases, offs, g_track)
        assert got.tolist() == want.tolist()
        assert_same_sketch(g_track, c_track)
    assert_same_sketch(g_counts, c_counts)   # the counts sketch is only read
    with pytest.raises(ValueError):
        g_counts.abundance_distribution_batch(bases, offs, getattr(kv.khmer, track_cls)(k + 2, 1000, 2))


# ------------------------------------------------------------------ simlike queries, mask generation

def test_simlike_spanning_abundances_golden(kv):
    """kevlar/simlike.py:22-96 on the minitrio sketches of kevlar/tests/test_simlike.py:21-31: equal
    to the reference's own function over the oracle (gen/simlike_spanning.json, whose first entry
    is the literal expectation of test_simlike.py:88-100), one window at a time and batched."""
    import json
    from kevlar_b200.simlike import spanning_kmer_abundances, spanning_kmer_abundances_many
    kid, mom, dad = (kv.khmer.Counttable(31, 1e6, 4) for _ in range(3))
    ref = kv.khmer.SmallCounttable(31, 125000, 4)
    for sk, fn in ((kid, 'trio-proband.fq.gz'), (mom, 'trio-mother.fq.gz'), (dad, 'trio-father.fq.gz'), (ref, 'refr.fa.gz')):
        sk.consume_seqfile(golden_data('minitrio/' + fn))
    cases = json.load(open(golden_gen('simlike_spanning.json')))
    assert len(cases) == 52
    for drop in (False, True):
        subset = [c for c in cases if c['dropoutliers'] == drop]
        many = spanning_kmer_abundances_many([(c['alt'], c['refr']) for c in subset], kid, (mom, dad), ref, dropoutliers=drop)
        for c, (abunds, refr_abunds, ndropped) in zip(subset, many):
            assert (abunds, refr_abunds, ndropped) == (c['abundances'], c['refr_abunds'], c['ndropped'])
    for c in cases[:6]:
        got = spanning_kmer_abundances(c['alt'], c['refr'], kid, (mom, dad), ref, dropoutliers=c['dropoutliers'])
        assert got == (c['abundances'], c['refr_abunds'], c['ndropped'])


def test_mask_from_windows(kv, oracle, tmp_path):
    """kevlar/call.py:136-172 / kevlar/alac.py:49-65: the --gen-mask Nodetable, byte-identical to
    consuming the windows one by one on the CPU."""
    from kevlar_b200.sketch import mask_from_windows
    genome = LETTERS[np.random.default_rng(8).integers(0, 4, size=5000)]
    windows = [w.decode() for w in random_reads(21, 300, lo=25, hi=90, genome=genome)] + [None, 'ACGT', '']
    out = str(tmp_path / 'mask.nt')
    log = io.StringIO()
    saved, kv.logstream = kv.logstream, log
    try:
        mask = mask_from_windows(windows, 31, 2000, maskfile=out, maxfpr=0.001)
    finally:
        kv.logstream = saved
    want = oracle.Nodetable(31, 2000 * 8 / 4, 4)
    for w in windows:
        if w is not None and len(w) >= 31:
            want.consume(w)
    assert_same_sketch(mask, want)
    ref_out = str(tmp_path / 'want.nt')
    want.save(ref_out)
    assert filecmp.cmp(out, ref_out, shallow=False)
    assert 'generating mask of variant-spanning k-mers' in log.getvalue()
    assert 'WARNING: mask FPR is' in log.getvalue() and 'exceeds user-specified limit of 0.0010' in log.getvalue()
    with pytest.raises(ValueError):
        mask_from_windows(['ACGTN' * 10], 31, 2000)


def test_simlike_reference_fixture_windows(kv):
    """The sketch + VCF fixtures of kevlar/tests/test_simlike.py (k = 49 and 31, 8-bit case/control
    and 4-bit reference sketches, one of them a one-bucket table): sketches loaded on the GPU, all
    windows of a fixture set in one batched query, equal to the reference's own
    spanning_kmer_abundances over the oracle (gen/simlike_fixture_windows.json)."""
    import json
    from kevlar_b200.simlike import spanning_kmer_abundances_many
    cases = json.load(open(golden_gen('simlike_fixture_windows.json')))
    assert len(cases) == 105
    groups = {}
    for c in cases:
        groups.setdefault((c['set'], tuple(c['sketches']), c['dropoutliers']), []).append(c)
    for (folder, files, drop), members in groups.items():
        sk = [kv.sketch.load(golden_data(folder + '/' + f)) for f in files]
        got = spanning_kmer_abundances_many([(c['alt'], c['refr']) for c in members], sk[0], sk[1:-1], sk[-1],
                                            dropoutliers=drop)
        for c, g in zip(members, got):
            assert g == (c['abundances'], c['refr_abunds'], c['ndropped']), (folder, c['alt'])


def test_single_bucket_table(kv, oracle, tmp_path):
    """khmer.Nodetable(31, 1, 1) (kevlar/tests/test_simlike.py:69) and the shipped one-bucket
    SmallCounttable term-high-abund/reference.sct."""
    assert kv._lib.primes_below(1, 1) == [1]
    shipped = kv.khmer.SmallCounttable.load(golden_data('term-high-abund/reference.sct'))
    assert shipped.hashsizes() == [1] and shipped.n_occupied() == 0
    out = str(tmp_path / 'one.sct')
    shipped.save(out)
    assert filecmp.cmp(out, golden_data('term-high-abund/reference.sct'), shallow=False)
    for name in ('Nodetable', 'SmallCounttable', 'Counttable'):
        g, c = getattr(kv.khmer, name)(31, 1, 1), getattr(oracle, name)(31, 1, 1)
        assert g.hashsizes() == [1]
        for seq in ('ACGT' * 10, 'GATTACA' * 9):
            g.consume(seq)
            c.consume(seq)
        assert_same_sketch(g, c)
        assert g.get('TTTT' * 7 + 'TTT') == c.get('TTTT' * 7 + 'TTT') > 0


def test_synth_reads_fixture(kv):
    """kv_synth_reads (measurement fixture): reads are a pure function of (seed, read index) -- slices
    drawn by different 'ranks' concatenate to the single-rank read set -- upper-case ACGT only, each read a
    haplotype window or its reverse complement up to the substitution errors."""
    torch = pytest.importorskip('torch')
    from kevlar_b200 import simtrio
    whole = simtrio.device_trio(200000, 5000, 0, 1)
    parts = [simtrio.device_trio(200000, 5000, r, 3) for r in range(3)]
    for s in range(3):
        b = whole[s][0].cpu().numpy()
        o = whole[s][1].cpu().numpy()
        assert (o == np.arange(5001) * 100).all()
        assert set(np.unique(b).tolist()) <= set(b'ACGT')
        joined = np.concatenate([p[s][0].cpu().numpy() for p in parts])
        assert (joined == b).all()
    haps = simtrio.trio_haplotypes(200000)[0]
    text = [h.tobytes() for h in haps]
    comp = bytes.maketrans(b'ACGT', b'TGCA')
    b = whole[0][0].cpu().numpy()
    exact = 0
    for r in range(200):
        read = b[r * 100:(r + 1) * 100].tobytes()
        if any(read in t or read.translate(comp)[::-1] in t for t in text):
            exact += 1
    assert 80 <= exact < 200   # (1 - 0.005)^100 = 61 % of the reads are error-free in expectation


@pytest.mark.parametrize('name', ['Counttable', 'SmallCounttable', 'Nodetable'])
def test_spanning_sketch_single_rank(kv, oracle, tmp_path, name):
    """A spanning sketch (CUDA virtual-memory tables + the collective tiled update) on ONE rank: the
    degenerate case of tests/_mgpu_worker.py, so that the allocation, the multi-source apply kernel
    and save / get / novel on such a sketch are covered on a single-GPU box too."""
    from kevlar_b200 import multigpu
    reads = random_reads(21, 3000, 40, 150) + [b'ACGTTGCAAGGCTTAACCGGTTAAACCCGGGTTTACGT'] * 500
    bases, offs = oracle.reads_to_batch(reads)
    sk = multigpu.SpanningSketch(getattr(kv.khmer, name), 21, 3000017, 4, chunk_positions=65536)
    c = getattr(oracle, name)(21, 3000017, 4)
    assert sk.consume_batch(bases, offs) == c.consume_batch(bases, offs)
    for t in range(4):
        assert sk.sketch.table_bytes(t) == c.table_bytes(t)
    assert sk.n_occupied() == c.n_occupied()
    kmer = 'ACGTTGCAAGGCTTAACCGGT'
    assert sk.sketch.get(kmer) == c.get(kmer) > 0
    sk.sketch.add(kmer)
    c.add(kmer)
    assert sk.sketch.get(kmer) == c.get(kmer)
    path = str(tmp_path / 'span.sketch')
    sk.save(path)
    c.save(path + '.o')
    assert open(path, 'rb').read() == open(path + '.o', 'rb').read()
    sk.clear()
    assert sk.n_occupied() == 0
    sk.close()


def _bands_namespace(kv, prefix, **over):
    from kevlar_b200 import bands
    na = [golden_data('microtrios/trio-na-{}.fq.gz'.format(w)) for w in ('proband', 'mother', 'father')]
    argv = ['--case', na[0], '--control', na[1], '--control', na[2], '-k', '31', '--memory', '500K', '--case-min', '5',
            '--ctrl-max', '1', '--num-bands', '8', '-n', '1', '--filter-memory', '1M', '--out-prefix', prefix]
    ns = bands.parser().parse_args(argv)
    for key, val in over.items():
        setattr(ns, key, val)
    return ns


def test_banded_chain_matches_reference(kv, tmp_path):
    """BASELINE config 5's chain on one GPU: `novel --num-bands 8 --band b` for b = 1..8, `unband`, `filter`
    recount -- every intermediate file equal to what the reference's own modules produce over the oracle
    (tests/golden/make_golden_bands.py).  The reference's band quirk makes the union of the bands differ
    from an unbanded run, so the pin is the reference chain, band by band."""
    from kevlar_b200 import bands
    prefix = str(tmp_path / 'chain')
    kv.logstream = io.StringIO()
    try:
        result = bands.run(_bands_namespace(kv, prefix), rank=0, world=1)
    finally:
        kv.logstream = None
    for b in range(1, 9):
        assert open('{}.band{}.augfastq'.format(prefix, b)).read() == open(golden_gen('bands8_band{}.out'.format(b))).read(), b
    assert open(result['unband']).read() == open(golden_gen('bands8_unband.out')).read()
    assert open(result['filter']).read() == open(golden_gen('bands8_filter.out')).read()
    assert open(result['filter']).read().count('@') >= 10


@pytest.mark.parametrize('name', ['Counttable', 'SmallCounttable', 'Nodetable', 'Countgraph'])
def test_unique_across_shards_on_one_gpu(kv, oracle, name, monkeypatch):
    """kv_sketch_occupancy + kv_unique_batch, the building blocks of n_unique_kmers across ranks
    (multigpu.unique_across_ranks), played through on ONE GPU: the reads are cut into three shards in file
    order, each shard is counted into its own zeroed sketch, shard r's first-touch passes run with the OR of the
    occupancy of shards < r as the occupied set, and the shares add up to the single-stream n_unique_kmers."""
    torch = pytest.importorskip('torch')
    from ctypes import byref, c_uint64, c_void_p
    from kevlar_b200 import multigpu
    from kevlar_b200._lib import check, lib
    reads = random_reads(61, 2500, 30, 140) + [b'ACGTTGCAAGGCTTAACCGGTTAAACCCGGGTTTACGT'] * 300 + random_reads(62, 500, 30, 140)
    bases, offs = oracle.reads_to_batch(reads)
    c = getattr(oracle, name)(21, 20000, 4)
    c.consume_batch(bases, offs)
    parts, shards = [], []
    for r in range(3):
        mb, mo = multigpu.shard_batch(bases, offs, r, 3)
        g = getattr(kv.khmer, name)(21, 20000, 4)
        g.set_unique_tracking(False)
        g.consume_batch(mb, mo)
        parts.append(g)
        shards.append((mb, mo))
    occs = []
    for g in parts:
        view, starts = multigpu._occupancy_view(g)
        kv._lib.sync(g.device)
        occs.append(view.clone())
    total = 0
    for r in range(3):
        lower = torch.zeros_like(occs[0])
        for q in range(r):
            lower |= occs[q]
        torch.cuda.synchronize()
        occ = (c_void_p * 4)(*[lower.data_ptr() + 4 * int(starts[t]) for t in range(4)])
        n = c_uint64()
        mb, mo = shards[r]
        check(lib().kv_unique_batch(parts[r]._h, occ, mb.ctypes.data, mo.ctypes.data, len(mo) - 1, kv.khmer.MEM_HOST, 0, 0, None, 0,
                                    0, byref(n), None))
        total += n.value
    assert total == c.n_unique_kmers()


@pytest.mark.parametrize('mode', [False, 'deferred'])
@pytest.mark.parametrize('name', ['Counttable', 'SmallCounttable', 'Nodetable', 'Countgraph'])
def test_unique_from_the_hashes_left_on_the_device(kv, oracle, name, mode):
    """kv_unique_last_batch: the same shares as kv_unique_batch, taken from the hashes the consume call left in
    the device scratch (no second copy of the reads, no second hash); KV_ESTATE -- and nothing done -- once
    another batch has been hashed.  'deferred': the consume call has already run table 0's first-touch pass."""
    torch = pytest.importorskip('torch')
    from ctypes import byref, c_uint64, c_void_p
    from kevlar_b200 import multigpu
    from kevlar_b200._lib import KV_ESTATE, check, lib
    reads = random_reads(71, 2500, 30, 140) + [b'ACGTTGCAAGGCTTAACCGGTTAAACCCGGGTTTACGT'] * 300 + random_reads(72, 500, 30, 140)
    reads[17] = b'ACGTNNACGT' * 5          # cleaned like everything else
    reads[18] = b'ACGT'                     # shorter than k
    bases, offs = oracle.reads_to_batch(reads)
    c = getattr(oracle, name)(21, 20000, 4)
    c.consume_batch(bases, offs)
    parts, occs, total = [], [], 0
    for r in range(3):
        mb, mo = multigpu.shard_batch(bases, offs, r, 3)
        g = getattr(kv.khmer, name)(21, 20000, 4)
        g.set_unique_tracking(mode)
        g.consume_batch(mb, mo)
        view, starts = multigpu._occupancy_view(g)
        lower = torch.zeros_like(view)
        for q in range(r):
            lower |= occs[q]
        torch.cuda.synchronize()
        occ = (c_void_p * 4)(*[lower.data_ptr() + 4 * int(starts[t]) for t in range(4)])
        n, n_again = c_uint64(), c_uint64()
        check(lib().kv_unique_last_batch(g._h, occ, byref(n), None))
        check(lib().kv_unique_batch(g._h, occ, mb.ctypes.data, mo.ctypes.data, len(mo) - 1, kv.khmer.MEM_HOST, 0, 0, None, 0, 0,
                                    byref(n_again), None))
        assert n.value == n_again.value, r
        if parts:   # the scratch now holds shard r: an earlier sketch cannot use the shortcut
            stale = c_uint64(123)
            assert lib().kv_unique_last_batch(parts[0]._h, occ, byref(stale), None) == KV_ESTATE
        kv._lib.sync(g.device)
        occs.append(view.clone())
        parts.append(g)
        total += n.value
    assert total == c.n_unique_kmers()
