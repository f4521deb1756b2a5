#!/usr/bin/env python
"""Regenerate tests/golden/ from the read-only reference checkout.

Run ONLY in the build container (needs /root/reference); the outputs are committed so
that neither the GPU box nor the CPU test-suite ever reads /root/reference.

Two kinds of vectors:

data/   verbatim copies of fixtures the reference's own tests assert on
        (kevlar/tests/data/...).  These are data files, not source.  Large FASTQ
        inputs are stored gzipped (decompressed content unchanged).
gen/    outputs of the reference's OWN, UNMODIFIED Python modules
        (kevlar/count.py, novel.py, filter.py, sketch.py, sequence.pyx ...) executed
        from a scratch copy under /tmp with the CPU oracle (oracle/khmer_oracle.py)
        standing in for the un-vendored ``khmer`` dependency.  They pin the host-side
        behaviour (read skipping, band quirk, abundance screen, annotation text,
        log lines) end to end.
reference_tests_over_oracle.log
        the reference's own pytest files for this path run on top of the oracle --
        the evidence that the oracle is a faithful khmer stand-in.
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
REFDATA = os.path.join(REF, 'kevlar', 'tests', 'data')

COPY = [
    'simple-genome-case.ct', 'simple-genome-ctrl1.ct', 'simple-genome-ctrl2.ct',
    'simple-genome-case-band-2-1.ct', 'simple-genome-case-band-16-7.ct',
    'simple-genome-case-reads.fa.gz', 'simple-genome-ctrl1-reads.fa.gz',
    'simple-genome-ctrl2-reads.fa.gz',
    'test.counttable', 'test.countgraph', 'test.smallcounttable', 'test.smallcountgraph',
    'test.nodetable', 'test.nodegraph', 'test.notasketchtype',
    'microtrios/trio-na-proband.fq.gz', 'microtrios/trio-na-mother.fq.gz',
    'microtrios/trio-na-father.fq.gz', 'microtrios/novel-na.augfastq.gz',
    'bogus-genome/refr.fa', 'bogus-genome/mask.nt', 'bogus-genome/mask-chr1.fa',
    'bogus-genome/mask-chr2.fa',
    'collect.alpha.txt', 'worm.augfasta', 'trio1/novel_3_1,2.txt',
    'screen-case.fa', 'screen-ctrl.fa', 'ambig.fasta',
    'example1.augfastq', 'example2.augfastq', 'example2.augfastq.gz',
    'minitrio/mask.nt', 'minitrio/trio-proband-mask-counts.ct', 'minitrio/trio-proband.fq.gz',
    'minitrio/trio-proband-dist.tsv', 'minitrio/trio-mother.fq.gz', 'minitrio/trio-father.fq.gz',
    # small sketches of the simlike tests: more OXLI files (k=49 Murmur, 4-bit tables, a size-1 table)
    'case-low-abund/dad.ct', 'case-low-abund/kid.ct', 'case-low-abund/mom.ct', 'case-low-abund/refr.sct',
    'ctrl-high-abund/cc57120.dad.sct', 'ctrl-high-abund/cc57120.kid.sct', 'ctrl-high-abund/cc57120.mom.sct',
    'ctrl-high-abund/cc57120.refr.sct',
    'term-high-abund/father.ct', 'term-high-abund/mother.ct', 'term-high-abund/proband.ct', 'term-high-abund/reference.sct',
    'partscore/partscore-father.ct', 'partscore/partscore-mother.ct', 'partscore/partscore-proband.ct',
    'partscore/partscore-refr.sct',
    'simlike-fast-mode/cc27.dad.ct', 'simlike-fast-mode/cc27.kid.ct', 'simlike-fast-mode/cc27.mom.ct',
    'simlike-fast-mode/cc27.refr.sct',
    'homopolymer/12175-dad.sct', 'homopolymer/12175-kid.sct', 'homopolymer/12175-mom.sct', 'homopolymer/12175-refr.sct',
]
COPY_GZ = ['trio1/case1.fq', 'trio1/ctrl1.fq', 'trio1/ctrl2.fq', 'minitrio/refr.fa']


def sha(path):
    h = hashlib.sha256()
    with open(path, 'rb') as fh:
        h.update(fh.read())
    return h.hexdigest()


def copy_fixtures():
    manifest = {}
    for rel in COPY:
        src = os.path.join(REFDATA, rel)
        if not os.path.exists(src):
            print('skip (absent):', rel)
            continue
        dst = os.path.join(HERE, 'data', rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = {'source': 'kevlar/tests/data/' + rel, 'sha256': sha(dst)}
    for rel in COPY_GZ:
        src = os.path.join(REFDATA, rel)
        dst = os.path.join(HERE, 'data', rel + '.gz')
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(src, 'rb') as fi, open(dst, 'wb') as fo:
            with gzip.GzipFile(fileobj=fo, mode='wb', mtime=0, compresslevel=9) as gz:
                gz.write(fi.read())
        manifest[rel + '.gz'] = {'source': 'kevlar/tests/data/' + rel + ' (gzipped here)',
                                 'sha256_uncompressed': sha(src)}
    return manifest


def scratch_reference():
    """Copy the reference package to /tmp, compile its Cython sequence module, and stub the
    imports that are irrelevant to this path (pysam, intervaltree, screed, the two other
    C extensions).  ``khmer`` resolves to the oracle."""
    root = tempfile.mkdtemp(prefix='kevlar_refrun_')
    shutil.copytree(os.path.join(REF, 'kevlar'), os.path.join(root, 'kevlar'),
                    ignore=shutil.ignore_patterns('data'))
    subprocess.check_call(['chmod', '-R', 'u+w', root])
    os.symlink(REFDATA, os.path.join(root, 'kevlar', 'tests', 'data'))
    subprocess.check_call(['cythonize', '-i', '-3', 'kevlar/sequence.pyx'], cwd=root,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with open(os.path.join(root, 'kevlar', 'alignment.py'), 'w') as fh:
        fh.write('def contig_align(*a, **k): raise NotImplementedError\n'
                 'def align_both_strands(*a, **k): raise NotImplementedError\n')
    with open(os.path.join(root, 'kevlar', 'assembly.py'), 'w') as fh:
        fh.write('def fml_asm(*a, **k): raise NotImplementedError\n')
    stubs = os.path.join(root, 'stubs')
    for mod, body in [
        ('khmer', 'import sys\nsys.path.insert(0, %r)\n'
                  'from oracle.khmer_oracle import (Counttable, SmallCounttable, Nodetable, Countgraph,\n'
                  '    SmallCountgraph, Nodegraph, ReadParser, _buckets_per_byte, khmer_args)\n'
                  'sys.modules["khmer.khmer_args"] = khmer_args\n'
                  'def calc_expected_collisions(*a, **k): return 0.0\n' % REPO),
        ('pysam', ''), ('screed', ''), ('intervaltree', 'class IntervalTree: pass\n'),
    ]:
        os.makedirs(os.path.join(stubs, mod))
        with open(os.path.join(stubs, mod, '__init__.py'), 'w') as fh:
            fh.write(body)
    env = dict(os.environ, PYTHONPATH=stubs + os.pathsep + root)
    return root, env


DRIVER = r'''
import io, json, sys, contextlib
import kevlar
D = sys.argv[1] + '/'
OUT = sys.argv[2] + '/'

def run(cmd, name):
    """Run one reference CLI invocation; save stdout/outfile text and the log."""
    args = kevlar.cli.parser().parse_args(cmd)
    log = io.StringIO()
    kevlar.logstream = log
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        kevlar.cli.mains[args.cmd](args)
    with open(OUT + name + '.out', 'w') as fh:
        fh.write(out.getvalue())
    with open(OUT + name + '.log', 'w') as fh:
        import re
        fh.write(re.sub(r'\d+\.\d\d sec(onds)?', 'T sec', log.getvalue()).replace(D, 'DATA/').replace(OUT, 'GEN/'))

na = [D + 'microtrios/trio-na-%s.fq.gz' % w for w in ('proband', 'mother', 'father')]
run(['novel', '-k', '31', '--case-min', '5', '--ctrl-max', '1', '--memory', '500K',
     '--case', na[0], '--control', na[1], '--control', na[2]], 'novel_microtrio_na')
run(['novel', '--case', na[0], '--ksize', '25', '--case-min', '7', '--control', na[2], '--control', na[1],
     '--num-bands', '2', '--band', '2', '--ctrl-max', '0', '--memory', '500K'], 'novel_microtrio_na_band2of2')
run(['novel', '--case', na[0], '--ksize', '25', '--case-min', '7', '--control', na[2], '--control', na[1],
     '--num-bands', '4', '--band', '3', '--ctrl-max', '0', '--memory', '500K'], 'novel_microtrio_na_band3of4')
run(['novel', '--ctrl-max', '0', '--case-min', '6', '--case', D + 'trio1/case1.fq',
     '--control', D + 'trio1/ctrl1.fq', '--control', D + 'trio1/ctrl2.fq',
     '--skip-until', 'bogus-genome-chr1_115_449_0:0:0_0:0:0_1f4/1'], 'novel_trio1_skipuntil')
run(['novel', '--ctrl-max', '0', '--case-min', '6', '--case', D + 'trio1/case1.fq',
     '--control', D + 'trio1/ctrl1.fq', '--control', D + 'trio1/ctrl2.fq'], 'novel_trio1')
run(['novel', '--ksize', '25', '--ctrl-max', '1', '--case-min', '8', '--case', D + 'screen-case.fa',
     '--control', D + 'screen-ctrl.fa', '--abund-screen', '3'], 'novel_abund_screen')
run(['novel', '--ksize', '25', '--ctrl-max', '1', '--case-min', '8', '--case', D + 'screen-case.fa',
     '--control', D + 'screen-ctrl.fa'], 'novel_no_abund_screen')
run(['novel', '-k', '25', '--case', D + 'simple-genome-case-reads.fa.gz', D + 'ambig.fasta',
     '--case-counts', D + 'simple-genome-case.ct',
     '--control-counts', D + 'simple-genome-ctrl1.ct', D + 'simple-genome-ctrl2.ct'], 'novel_load_counts')
run(['filter', '--mask', D + 'bogus-genome/mask.nt', '--memory', '10M', '--max-fpr', '0.001',
     '--case-min', '6', D + 'trio1/novel_3_1,2.txt'], 'filter_trio1_mask')
run(['count', '--ksize', '21', '--memory', '1M', '--mask', D + 'bogus-genome/mask.nt',
     OUT + 'count_refr_masked.ct', D + 'bogus-genome/refr.fa'], 'count_refr_masked')
run(['count', '--ksize', '21', '--memory', '200K', '-c', '4', OUT + 'count_refr_small.sct',
     D + 'bogus-genome/refr.fa'], 'count_refr_small')
run(['count', '--ksize', '21', '--memory', '100K', '-c', '1', '--num-bands', '3', '--band', '2',
     OUT + 'count_refr_node_band.nt', D + 'bogus-genome/refr.fa'], 'count_refr_node_band')
run(['count', '--ksize', '27', '--memory', '500K', OUT + 'count_na_proband.ct', na[0]], 'count_na_proband')

# Python-API filter calls used by kevlar/tests/test_filter.py
def run_filter(name, readfile, **kw):
    log = io.StringIO()
    kevlar.logstream = log
    out = io.StringIO()
    for rec in kevlar.filter.filter(readfile, **kw):
        kevlar.print_augmented_fastx(rec, out)
    with open(OUT + name + '.out', 'w') as fh:
        fh.write(out.getvalue())

run_filter('filter_alpha', D + 'collect.alpha.txt', memory=500)
run_filter('filter_worm', D + 'worm.augfasta', memory=1000, casemin=5, ctrlmax=0)
run_filter('filter_trio1_nomask', D + 'trio1/novel_3_1,2.txt', memory=1e7)

# simlike sketch queries: the reference's own spanning_kmer_abundances (kevlar/simlike.py:51-96)
# on the minitrio sketches of kevlar/tests/test_simlike.py:21-31
import khmer, random
from kevlar.simlike import spanning_kmer_abundances
kid, mom, dad = (khmer.Counttable(31, 1e6, 4) for _ in range(3))
ref = khmer.SmallCounttable(31, 125000, 4)
for sk, fn in ((kid, 'trio-proband.fq.gz'), (mom, 'trio-mother.fq.gz'), (dad, 'trio-father.fq.gz'), (ref, 'refr.fa')):
    sk.consume_seqfile(D + 'minitrio/' + fn)
ALT = 'TGTCTCCCTCCCCTCCACCCCCAGAAATGGGTTTTTGATAGTCTTCCAAAGTTAGGGTAGT'
windows = [(ALT, 'TGTCTCCCTCCCCTCCACCCCCAGAAATGGCTTTTTGATAGTCTTCCAAAGTTAGGGTAGT'),
           (ALT, 'TGTCTCCCTCCCCTCCACCCCCAGAAATGGGAAATTTTTGATAGTCTTCCAAAGTTAGGGTAGT')]
rng = random.Random(2018)
genome = ''.join(line.strip() for line in open(D + 'minitrio/refr.fa') if not line.startswith('>')).upper()
for _ in range(12):   # SNV-like and deletion-like windows cut from the genome
    at = rng.randrange(100, len(genome) - 200)
    refw = genome[at:at + 61]
    alt = refw[:30] + rng.choice([c for c in 'ACGT' if c != refw[30]]) + refw[31:]
    windows.append((alt, refw))
    windows.append((refw[:28] + refw[33:], refw))
cases = []
for alt, refw in windows:
    for drop in (False, True):
        abunds, refrabunds, ndropped = spanning_kmer_abundances(alt, refw, kid, (mom, dad), ref, dropoutliers=drop)
        cases.append({'alt': alt, 'refr': refw, 'dropoutliers': drop, 'abundances': abunds,
                      'refr_abunds': refrabunds, 'ndropped': ndropped})
assert cases[0]['ndropped'] == 3 and cases[0]['abundances'][0][:4] == [7, 6, 6, 6]   # test_simlike.py:88-91
with open(OUT + 'simlike_spanning.json', 'w') as fh:
    json.dump(cases, fh, indent=0)

# the same function on the sketch + VCF fixtures of kevlar/tests/test_simlike.py:46-80,250-330 (k = 49 and 31,
# 8-bit case/control sketches, 4-bit reference sketches): windows come from the calls' ALTWINDOW / REFRWINDOW
fixture_sets = [
    ('simlike-fast-mode', 'cc27.kid.ct', ('cc27.mom.ct', 'cc27.dad.ct'), 'cc27.refr.sct', 'cc27.calls.vcf'),
    ('ctrl-high-abund', 'cc57120.kid.sct', ('cc57120.mom.sct', 'cc57120.dad.sct'), 'cc57120.refr.sct', 'cc57120.calls.vcf'),
    ('case-low-abund', 'kid.ct', ('mom.ct', 'dad.ct'), 'refr.sct', 'calls.vcf.gz'),
    ('term-high-abund', 'proband.ct', ('mother.ct', 'father.ct'), 'reference.sct', 'calls.vcf'),
]
fixture_cases = []
for folder, kidf, ctrlf, refrf, vcff in fixture_sets:
    base = D + folder + '/'
    kidsk = kevlar.sketch.load(base + kidf)
    ctrlsk = [kevlar.sketch.load(base + f) for f in ctrlf]
    refsk = kevlar.sketch.load(base + refrf)
    for call in kevlar.vcf.VCFReader(kevlar.open(base + vcff, 'r')):
        alt, refw = call.window, call.refrwindow
        if alt is None or refw is None or len(alt) < kidsk.ksize() or len(refw) < kidsk.ksize():
            continue
        for drop in (False, True):
            try:
                abunds, refrabunds, ndropped = spanning_kmer_abundances(alt, refw, kidsk, ctrlsk, refsk, dropoutliers=drop)
            except ZeroDivisionError:   # every k-mer is in the reference: the outlier filter divides by zero
                continue
            fixture_cases.append({'set': folder, 'sketches': [kidf] + list(ctrlf) + [refrf], 'alt': alt, 'refr': refw,
                                  'dropoutliers': drop, 'abundances': abunds, 'refr_abunds': refrabunds,
                                  'ndropped': ndropped})
assert len(fixture_cases) > 20 and any(max(c['abundances'][0] + [0]) > 5 for c in fixture_cases)
with open(OUT + 'simlike_fixture_windows.json', 'w') as fh:
    json.dump(fixture_cases, fh, indent=0)
'''


def generate(env, root):
    gen = os.path.join(HERE, 'gen')
    shutil.rmtree(gen, ignore_errors=True)
    os.makedirs(gen)
    drv = os.path.join(root, 'driver.py')
    with open(drv, 'w') as fh:
        fh.write(DRIVER)
    subprocess.check_call([sys.executable, drv, REFDATA, gen], env=env, cwd=root)
    # test_simlike.py pins get_kmer_counts (k = 31 and 49, 8- and 4-bit tables) through the reference's
    # own likelihood scores and filter calls
    tests = ['test_count.py', 'test_sketch.py', 'test_novel.py', 'test_filter.py', 'test_seqio.py',
             'test_unband.py', 'test_simlike.py']
    # test_dist.py: the tests that go through compute_dist() need DataFrame.append, which the pandas
    # in this image (3.x) no longer has, and test_calc_mu_sigma asserts on a bare pytest.approx(),
    # which pytest 9 rejects -- the reference itself cannot run those here
    dist_tests = ['test_dist.py::' + t for t in ('test_count_first_pass', 'test_count_second_pass',
                                                 'test_musigma_empty_dist')]
    res = subprocess.run([sys.executable, '-m', 'pytest', '-q', '-p', 'no:cacheprovider', '-W',
                          'ignore'] + ['kevlar/tests/' + t for t in tests + dist_tests],
                         env=env, cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(os.path.join(HERE, 'reference_tests_over_oracle.log'), 'w') as fh:
        fh.write('# reference pytest files run UNMODIFIED with oracle/khmer_oracle.py as `khmer`\n')
        fh.write('# files: ' + ' '.join(tests + dist_tests) + '\n')
        keep = [ln for ln in res.stdout.splitlines() if 'passed' in ln or 'failed' in ln or 'error' in ln.lower()]
        fh.write('\n'.join(keep[-5:]) + '\n')
    print(res.stdout.splitlines()[-1])


def main():
    if not os.path.isdir(REF):
        sys.exit('needs /root/reference (build container only)')
    sys.path.insert(0, REPO)
    from oracle import khmer_oracle  # noqa: F401  (builds the oracle library)
    manifest = copy_fixtures()
    root, env = scratch_reference()
    try:
        generate(env, root)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    gen = os.path.join(HERE, 'gen')
    for fn in sorted(os.listdir(gen)):
        path = os.path.join(gen, fn)
        manifest['gen/' + fn] = {'source': 'reference Python over oracle (make_golden.py DRIVER)',
                                 'sha256': sha(path), 'bytes': os.path.getsize(path)}
        if os.path.getsize(path) > 300000:      # keep the repo small: hash only
            manifest['gen/' + fn]['stored'] = False
            os.remove(path)
    with open(os.path.join(HERE, 'MANIFEST.json'), 'w') as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print('wrote', len(manifest), 'entries')


if __name__ == '__main__':
    main()
