"""`kevlar count`: build a k-mer abundance sketch from reads.

Same public surface, thresholds and log lines as kevlar/count.py (load_sample_seqfile :18-99,
print_config :102-116, main :119-140); the read loop the reference delegates to khmer's C++
runs here as CUDA kernels, one kv_consume_batch per batch of reads."""
import threading

import kevlar_b200
from kevlar_b200 import khmer
from kevlar_b200.sketch import allocate, get_extension

NUM_TABLES = 4
_SKETCH_KIND = {(True, True): 'smallcountgraph', (True, False): 'countgraph',
                (False, True): 'nodegraph', (False, False): 'nodegraph'}
_COUNTER_DESCRIPTION = {
    1: 'Storing k-mers in a node table (Bloom filter) for k-mer presence/absence queries',
    4: 'Storing k-mers in a small count table, a CountMin sketch with a counter size of 4 bits, '
       'for k-mer abundance queries (max abundance 15)',
    8: 'Storing k-mers in a count table, a CountMin sketch with a counter size of 8 bits, '
       'for k-mer abundance queries (max abundance 255)',
}


def _consume_call(sketch, parser, mask, maskmaxabund, consume_masked, numbands, band):
    """Which khmer-style consume method a worker runs, with its arguments (kevlar/count.py:43-71)."""
    if mask:
        options = dict(consume_masked=consume_masked, threshold=1 if consume_masked else maskmaxabund)
        if numbands:
            return sketch.consume_seqfile_banding_with_mask, (parser, numbands, band, mask), options
        return sketch.consume_seqfile_with_mask, (parser, mask), options
    if numbands:
        return sketch.consume_seqfile_banding, (parser, numbands, band), {}
    return sketch.consume_seqfile, (parser,), {}


def _consume_file(sketch, seqfile, numthreads, **how):
    """All workers drain one shared parser into the sketch; returns the number of reads."""
    kevlar_b200.plog('[kevlar::count]', '- processing "{}"'.format(seqfile))
    parser = khmer.ReadParser(seqfile)
    pool = []
    for _ in range(numthreads):
        method, args, kwargs = _consume_call(sketch, parser, **how)
        pool.append(threading.Thread(target=method, args=args, kwargs=kwargs))
    for worker in pool:
        worker.start()
    for worker in pool:
        worker.join()
    return parser.num_reads


def load_sample_seqfile(seqfiles, ksize, memory, maxfpr=0.2, count=True, smallcount=False, mask=None,
                        maskmaxabund=0, consume_masked=False, numbands=None, band=None, outfile=None,
                        numthreads=1):
    """Sketch of all k-mers in one sample's FASTA/FASTQ files: four tables sized from `memory`;
    with `mask`, k-mers present in it are left out (or, with `consume_masked`, only those are
    counted).  Raises KevlarUnsuitableFPRError when the sketch is too full."""
    kind = _SKETCH_KIND[(bool(count), bool(smallcount))]
    tablesize = memory / NUM_TABLES * khmer._buckets_per_byte[kind]
    sketch = allocate(ksize, tablesize, num_tables=NUM_TABLES, count=count, smallcount=smallcount)
    how = dict(mask=mask, maskmaxabund=maskmaxabund, consume_masked=consume_masked, numbands=numbands, band=band)
    numreads = sum(_consume_file(sketch, seqfile, numthreads, **how) for seqfile in seqfiles)

    fpr = kevlar_b200.sketch.estimate_fpr(sketch)
    report = ['Done loading k-mers' + (' (band {:d}/{:d})'.format(band + 1, numbands) if numbands else ''),
              '    {:d} reads processed, {:d} distinct k-mers stored'.format(numreads, sketch.n_unique_kmers()),
              '    estimated false positive rate is {:1.3f}'.format(fpr)]
    if fpr > maxfpr:
        report[-1] += ' (FPR too high, bailing out!!!)'
        raise kevlar_b200.sketch.KevlarUnsuitableFPRError('[kevlar::count] ' + ';\n'.join(report))
    if outfile:
        extensions = get_extension(count=count, smallcount=smallcount)
        if not outfile.endswith(extensions):
            outfile += extensions[1]
        sketch.save(outfile)
        report.append('    saved to "{:s}"'.format(outfile))
    kevlar_b200.plog('[kevlar::count]', ';\n'.join(report))
    return sketch


def print_config(args):
    kevlar_b200.plog('[kevlar::count]', _COUNTER_DESCRIPTION[args.counter_size])


def main(args):
    if (args.num_bands is None) is not (args.band is None):
        raise ValueError('Must specify --num-bands and --band together')
    band = args.band - 1 if args.band else None   # 1-based on the command line
    if args.mask:
        args.mask = kevlar_b200.sketch.load(args.mask)
    print_config(args)
    clock = kevlar_b200.Timer()
    clock.start()
    load_sample_seqfile(args.seqfile, args.ksize, args.memory, args.max_fpr, count=args.counter_size > 1,
                        smallcount=args.counter_size == 4, mask=args.mask, consume_masked=args.count_masked,
                        numbands=args.num_bands, band=band, numthreads=args.threads, outfile=args.counttable)
    kevlar_b200.plog('[kevlar::count] Total time: {:.2f} seconds'.format(clock.stop()))
