"""Banded runs across GPUs (BASELINE config 5): `kevlar novel --num-bands N --band b` for every band,
`kevlar unband`, `kevlar filter`.

The reference runs k-mer banding as N sequential passes for an N-fold memory cut (docs/banding.rst:4-8),
one `kevlar novel --band b` process per band (workflows/mark-I/Snakefile), then `kevlar unband`
(kevlar/unband.py:26-78) and the `kevlar filter` recount (kevlar/filter.py:15-96).  Bands are independent,
so here band b runs on rank (b-1) mod world -- one process per GPU under torchrun, no data-path collective --
and rank 0 merges.  Every band uses the reference's own predicates unchanged: khmer's hash-RANGE banding while
counting (kevlar/cli/count.py:64-69) and the bit-mask test of the novel loop (kevlar/novel.py:144-147, with
its quirk: band 1 reports nothing, band b keeps low bits b-2 -- SURVEY App. B.1).

    python -m torch.distributed.run --nproc-per-node 8 -m kevlar_b200.bands --num-bands 8 \\
        --case proband.fq --control mother.fq --control father.fq -k 31 --memory 500K --out-prefix run1
"""
import argparse
import sys

import kevlar_b200
from kevlar_b200 import multigpu


def band_of_rank(num_bands, rank, world):
    """The (1-based) bands rank `rank` runs: b with (b-1) mod world == rank."""
    return [b for b in range(1, num_bands + 1) if (b - 1) % world == rank]


def novel_band_args(ns, band, out):
    """Command line of the reference-shaped `novel` call for one band."""
    argv = ['novel', '--ksize', str(ns.ksize), '--memory', str(ns.memory), '--max-fpr', str(ns.max_fpr),
            '--case-min', str(ns.case_min), '--ctrl-max', str(ns.ctrl_max), '--num-bands', str(ns.num_bands),
            '--band', str(band), '--out', out]
    if ns.abund_screen:
        argv += ['--abund-screen', str(ns.abund_screen)]
    for files in ns.case:
        argv += ['--case'] + list(files)
    for files in ns.control:
        argv += ['--control'] + list(files)
    return argv


def run(ns, rank=None, world=None, barrier=None):
    """Run this rank's bands, then (rank 0) unband + filter.  Returns the paths written:
    {'bands': [...this rank's...], 'unband': path or None, 'filter': path or None}."""
    import kevlar_b200.cli
    if rank is None:
        rank, world = multigpu.init_from_env()
    written = []
    for band in band_of_rank(ns.num_bands, rank, world):
        out = '{}.band{}.augfastq'.format(ns.out_prefix, band)
        args = kevlar_b200.cli.parser().parse_args(novel_band_args(ns, band, out))
        kevlar_b200.plog('[kevlar::bands] rank {} of {}: band {} of {}'.format(rank, world, band, ns.num_bands))
        kevlar_b200.cli.mains['novel'](args)
        written.append(out)
    if barrier is not None:
        barrier()
    elif world > 1:
        multigpu.dist().barrier()
    result = {'bands': written, 'unband': None, 'filter': None}
    if rank == 0:
        band_files = ['{}.band{}.augfastq'.format(ns.out_prefix, b) for b in range(1, ns.num_bands + 1)]
        merged = ns.out_prefix + '.unband.augfastq'
        with kevlar_b200.open(merged, 'w') as fh:
            for read in kevlar_b200.unband.unband(kevlar_b200.unband.afxstream(band_files), ns.n_batches):
                kevlar_b200.print_augmented_fastx(read, fh)
        result['unband'] = merged
        if not ns.no_filter:
            mask = kevlar_b200.sketch.load(ns.mask) if ns.mask else None
            filtered = ns.out_prefix + '.filtered.augfastq'
            with kevlar_b200.open(filtered, 'w') as fh:
                for read in kevlar_b200.filter.filter(merged, mask=mask, memory=ns.filter_memory, maxfpr=ns.filter_max_fpr,
                                                      casemin=ns.case_min, ctrlmax=ns.ctrl_max):
                    kevlar_b200.print_augmented_fastx(read, fh)
            result['filter'] = filtered
    if world > 1:
        multigpu.dist().barrier()
    return result


def parser():
    from kevlar_b200.khmer import khmer_args
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--case', nargs='+', action='append', required=True, metavar='F')
    ap.add_argument('--control', nargs='+', action='append', default=[], metavar='F')
    ap.add_argument('-k', '--ksize', type=int, default=31)
    ap.add_argument('-M', '--memory', type=khmer_args.memory_setting, default=1e6, help='sketch memory per sample and band')
    ap.add_argument('--max-fpr', type=float, default=0.2)
    ap.add_argument('-y', '--case-min', type=int, default=6)
    ap.add_argument('-x', '--ctrl-max', type=int, default=1)
    ap.add_argument('--abund-screen', type=int, default=None)
    ap.add_argument('--num-bands', type=int, required=True)
    ap.add_argument('-n', '--n-batches', type=int, default=16, help='unband batches (kevlar/cli/unband.py)')
    ap.add_argument('--no-filter', action='store_true')
    ap.add_argument('--mask', default=None, help='mask sketch for the filter recount')
    ap.add_argument('--filter-memory', type=khmer_args.memory_setting, default=1e6)
    ap.add_argument('--filter-max-fpr', type=float, default=0.01)
    ap.add_argument('--out-prefix', required=True)
    return ap


def main(argv=None):
    ns = parser().parse_args(argv)
    rank, world = multigpu.init_from_env()
    try:
        result = run(ns, rank, world)
        if rank == 0:
            print(result['filter'] or result['unband'])
    finally:
        if world > 1:
            multigpu.dist().destroy_process_group()


if __name__ == '__main__':
    main(sys.argv[1:])
