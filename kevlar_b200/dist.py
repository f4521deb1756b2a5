"""`kevlar dist`: abundance distribution of the k-mers a mask selects (kevlar/dist.py:25-170).

Two passes over the reads, both on the GPU:
  1. count ONLY the k-mers present in the mask (`consume_seqfile_with_mask(threshold=1,
     consume_masked=True)`, kevlar/dist.py:25-46);
  2. histogram the counts of the DISTINCT k-mers, "distinct" being decided by a tracking
     Nodetable with the counts' table sizes (`abundance_distribution`, kevlar/dist.py:49-79) --
     one `kv_abund_dist_batch` call per batch of reads.
The reference runs pass 2 on several threads sharing one tracking table, which makes its result
depend on thread timing; here a file is one ordered stream of batches, i.e. the reference's
single-thread answer, whatever `threads` says.
"""
import json
import math

import numpy

import kevlar_b200
from kevlar_b200 import khmer

TSV_COLUMNS = ('Abundance', 'Count', 'CumulativeCount', 'CumulativeFraction')


class KevlarZeroAbundanceDistError(ValueError):
    pass


def count_first_pass(infiles, counts, mask, nthreads=1):
    kevlar_b200.plog('[kevlar::dist]', 'Processing input with {:d} threads'.format(nthreads))
    for filename in infiles:
        kevlar_b200.plog('    -', filename)
        counts.consume_seqfile_with_mask(khmer.ReadParser(filename), mask, threshold=1, consume_masked=True)
    kevlar_b200.plog('[kevlar::dist] Done processing input!')


def count_second_pass(infiles, counts, nthreads=1):
    kevlar_b200.plog('[kevlar::dist] Second pass over the data')
    tracking = khmer.Nodetable(counts.ksize(), 1, 1, primes=counts.hashsizes())
    total = numpy.zeros(65536, dtype=numpy.uint64)
    for filename in infiles:
        kevlar_b200.plog('    -', filename)
        total += numpy.asarray(counts.abundance_distribution(khmer.ReadParser(filename), tracking), dtype=numpy.uint64)
    kevlar_b200.plog('[kevlar::dist] Done second pass over input!')
    # abundance 0 (k-mers outside the mask) is not part of the distribution
    return {int(a): int(total[a]) for a in numpy.nonzero(total)[0] if a > 0}


def weighted_mean_std_dev(values, weights):
    values = numpy.asarray(values, dtype=float)
    mu = numpy.average(values, weights=weights)
    sigma = math.sqrt(numpy.average((values - mu) ** 2, weights=weights))
    return mu, sigma


def calc_mu_sigma(abundance):
    if sum(abundance.values()) == 0:
        raise KevlarZeroAbundanceDistError('all k-mer abundances are 0, please check input files')
    return weighted_mean_std_dev(list(abundance.keys()), list(abundance.values()))


def compute_dist(abundance):
    """Table of (Abundance, Count, CumulativeCount, CumulativeFraction), all float columns as in
    the reference's TSV (kevlar/tests/data/minitrio/trio-proband-dist.tsv)."""
    import pandas   # deferred: only this table needs it
    abunds = sorted(abundance)
    counts = numpy.array([abundance[a] for a in abunds], dtype=float)
    assert (counts > 0).all(), abundance
    cumulative = numpy.cumsum(counts)
    table = {
        'Abundance': numpy.array(abunds, dtype=float),
        'Count': counts,
        'CumulativeCount': cumulative,
        'CumulativeFraction': cumulative / counts.sum() if len(counts) else cumulative,
    }
    return pandas.DataFrame(table, columns=list(TSV_COLUMNS))


def dist(infiles, mask, ksize=31, memory=1e6, threads=1):
    counts = khmer.Counttable(ksize, memory / 4, 4)
    count_first_pass(infiles, counts, mask, nthreads=threads)
    abundance = count_second_pass(infiles, counts, nthreads=threads)
    mu, sigma = calc_mu_sigma(abundance)
    return mu, sigma, compute_dist(abundance)


def plot_dist(data, mu, sigma, filename, xlim=(0, 100)):
    try:
        import matplotlib
        matplotlib.use('Agg')
        from matplotlib import pyplot
    except ImportError:
        raise RuntimeError('--plot needs matplotlib, which is not installed here; use --tsv instead')
    figure, axes = pyplot.subplots(figsize=(12, 6))
    axes.plot(data['Abundance'], data['Count'], color='blue')
    axes.axvline(x=mu, color='blue', linestyle='--')
    for edge in (mu - sigma, mu + sigma):
        axes.axvline(x=edge, color='red', linestyle=':')
    axes.set_xlim(xlim)
    axes.set_xlabel('K-mer abundance', fontsize=16)
    axes.set_ylabel('Frequency', fontsize=16)
    figure.savefig(filename, dpi=300)


def main(args):
    mask = khmer.Nodetable.load(args.mask)
    mu, sigma, data = dist(args.infiles, mask, ksize=args.ksize, memory=args.memory, threads=args.threads)
    outstream = kevlar_b200.open(args.out, 'w') if getattr(args, 'out', None) else None
    print(json.dumps({'mu': mu, 'sigma': sigma}), file=outstream)
    if args.tsv:
        data.to_csv(args.tsv, sep='\t', index=False)
    if args.plot:
        plot_dist(data, mu, sigma, args.plot, xlim=args.plot_xlim)
