#!/usr/bin/env python
"""Golden outputs for BASELINE config 5's chain -- banded `kevlar novel` (--num-bands 8, every band),
`kevlar unband`, `kevlar filter` -- produced by the reference's OWN, UNMODIFIED modules over the CPU
oracle (same staging as make_golden.py; build container only, needs /root/reference).

Input: the microtrio `na` reads (3 x 6000 x 100 bp), k=31, --memory 500K, case-min 5, ctrl-max 1.
Outputs (tests/golden/gen/): bands8_band{1..8}.out, bands8_unband.out (reference unband with one batch, so
the order is by read name), bands8_filter.out (reference filter, --memory 1M).
Note the reference's band quirk (kevlar/novel.py:144-147): band 1 reports nothing and band
b >= 2 keeps hashes whose low bits equal b-2, inside the hash RANGE of band b -- the union over the bands is
NOT the unbanded result, which is why the chain is pinned against the reference and not against an
unbanded run."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402

DRIVER = r'''
import io, sys, contextlib
import kevlar
D = sys.argv[1] + '/'
OUT = sys.argv[2] + '/'
na = [D + 'microtrios/trio-na-%s.fq.gz' % w for w in ('proband', 'mother', 'father')]

def run(cmd):
    args = kevlar.cli.parser().parse_args(cmd)
    kevlar.logstream = io.StringIO()
    with contextlib.redirect_stdout(io.StringIO()):
        kevlar.cli.mains[args.cmd](args)

bands = []
for b in range(1, 9):
    out = OUT + 'bands8_band%d.out' % b
    run(['novel', '-k', '31', '--case-min', '5', '--ctrl-max', '1', '--memory', '500K', '--num-bands', '8', '--band', str(b),
         '--case', na[0], '--control', na[1], '--control', na[2], '--out', out])
    bands.append(out)
import os
# band 1 reports nothing (the quirk above) and the reference's reader yields None for an empty file, which
# its unband then trips over (kevlar/unband.py:36): only the non-empty band files are passed on
run(['unband', '-n', '1', '--out', OUT + 'bands8_unband.out'] + [b for b in bands if os.path.getsize(b)])
# `kevlar filter` needs --mask on the command line (kevlar/filter.py:100 loads it unconditionally): the Python API
# is what the reference's own tests call without one (kevlar/tests/test_filter.py:22-24)
kevlar.logstream = io.StringIO()
with open(OUT + 'bands8_filter.out', 'w') as fh:
    for rec in kevlar.filter.filter(OUT + 'bands8_unband.out', memory=1e6, casemin=5, ctrlmax=1):
        kevlar.print_augmented_fastx(rec, fh)
'''


def main():
    if not os.path.isdir(make_golden.REF):
        sys.exit('needs /root/reference (build container only)')
    sys.path.insert(0, make_golden.REPO)
    from oracle import khmer_oracle  # noqa: F401
    root, env = make_golden.scratch_reference()
    try:
        drv = os.path.join(root, 'driver_bands.py')
        with open(drv, 'w') as fh:
            fh.write(DRIVER)
        gen = os.path.join(HERE, 'gen')
        subprocess.check_call([sys.executable, drv, make_golden.REFDATA, gen], env=env, cwd=root)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    for fn in sorted(os.listdir(os.path.join(HERE, 'gen'))):
        if fn.startswith('bands8_'):
            print(fn, os.path.getsize(os.path.join(HERE, 'gen', fn)))


if __name__ == '__main__':
    main()
