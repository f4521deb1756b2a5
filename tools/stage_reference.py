#!/usr/bin/env python
"""Stage an UNMODIFIED copy of the reference package under baseline/_ref/ (build container only).

baseline/_ref/ is git-ignored (no reference source enters the history) but travels to the GPU
box with the snapshot, where tests/test_reference_suite.py runs the reference's OWN pytest files
for the count / novel / filter path with `khmer` resolving to kevlar_b200.khmer -- i.e. the
reference's host code and tests on top of the CUDA library.

Layout:  baseline/_ref/kevlar/        the package as it is in /root/reference (tests + data included),
                                      sequence.pyx compiled in place with cythonize
         baseline/_ref/stubs/         `khmer` -> kevlar_b200.khmer; empty pysam / screed / intervaltree
                                      (imported by kevlar/__init__.py, unused on this path)
The two C extensions unrelated to the path (alignment, assembly) become stubs that raise.
"""
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
DEST = os.path.join(REPO, 'baseline', '_ref')


def main():
    if not os.path.isdir(REF):
        sys.exit('needs /root/reference (build container only)')
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST)
    shutil.copytree(os.path.join(REF, 'kevlar'), os.path.join(DEST, 'kevlar'))
    subprocess.check_call(['chmod', '-R', 'u+w', DEST])
    subprocess.check_call(['cythonize', '-i', '-3', 'kevlar/sequence.pyx'], cwd=DEST, stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    shutil.rmtree(os.path.join(DEST, 'build'), ignore_errors=True)
    with open(os.path.join(DEST, 'kevlar', 'alignment.py'), 'w') as fh:
        fh.write('def contig_align(*a, **k): raise NotImplementedError\n'
                 'def align_both_strands(*a, **k): raise NotImplementedError\n')
    with open(os.path.join(DEST, 'kevlar', 'assembly.py'), 'w') as fh:
        fh.write('def fml_asm(*a, **k): raise NotImplementedError\n')
    stubs = os.path.join(DEST, 'stubs')
    for mod, body in [
        ('khmer', 'from kevlar_b200.khmer import *  # noqa: F401,F403\n'
                  'from kevlar_b200.khmer import _buckets_per_byte, khmer_args, calc_expected_collisions  # noqa: F401\n'
                  'import sys\nsys.modules["khmer.khmer_args"] = khmer_args\n'),
        ('pysam', ''), ('screed', ''), ('intervaltree', 'class IntervalTree: pass\n'),
    ]:
        os.makedirs(os.path.join(stubs, mod))
        with open(os.path.join(stubs, mod, '__init__.py'), 'w') as fh:
            fh.write(body)
    with open(os.path.join(DEST, 'README'), 'w') as fh:
        fh.write('Unmodified copy of /root/reference/kevlar staged by tools/stage_reference.py (git-ignored).\n')
    size = subprocess.check_output(['du', '-sh', DEST], text=True).split()[0]
    print('staged', DEST, size)


if __name__ == '__main__':
    main()
