"""Worker for tests/test_multigpu.py: the banded chain with one band per rank (torchrun), outputs compared
with the reference chain's golden files by rank 0."""
import io
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))


def main():
    import kevlar_b200 as kv
    from kevlar_b200 import bands, multigpu
    from conftest import golden_data, golden_gen
    rank, world = multigpu.init_from_env()
    prefix = os.path.join(sys.argv[1], 'chain')
    na = [golden_data('microtrios/trio-na-{}.fq.gz'.format(w)) for w in ('proband', 'mother', 'father')]
    ns = bands.parser().parse_args(['--case', na[0], '--control', na[1], '--control', na[2], '-k', '31', '--memory', '500K',
                                    '--case-min', '5', '--ctrl-max', '1', '--num-bands', '8', '-n', '1', '--filter-memory', '1M',
                                    '--out-prefix', prefix])
    kv.logstream = io.StringIO()
    result = bands.run(ns, rank, world)
    mine = bands.band_of_rank(8, rank, world)
    assert [os.path.basename(p) for p in result['bands']] == ['chain.band{}.augfastq'.format(b) for b in mine]
    ok = True
    if rank == 0:
        for b in range(1, 9):
            ok = ok and open('{}.band{}.augfastq'.format(prefix, b)).read() == open(golden_gen('bands8_band{}.out'.format(b))).read()
        ok = ok and open(result['unband']).read() == open(golden_gen('bands8_unband.out')).read()
        ok = ok and open(result['filter']).read() == open(golden_gen('bands8_filter.out')).read()
        print('banded chain OK on {} ranks'.format(world) if ok else 'banded chain MISMATCH')
    multigpu.dist().destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
