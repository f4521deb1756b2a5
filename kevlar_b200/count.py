"""`kevlar count`: build a k-mer abundance sketch from reads -- same interface and log lines
as kevlar/count.py, with the read loop running as CUDA kernels (kv_consume_batch)."""
import threading

import kevlar_b200
from kevlar_b200 import khmer
from kevlar_b200.sketch import allocate, get_extension


def _consumer(sketch, parser, mask, maskmaxabund, consume_masked, numbands, band):
    """Pick the khmer-style consume call for one worker (kevlar/count.py:43-71)."""
    if mask:
        kwargs = {'consume_masked': consume_masked, 'threshold': 1 if consume_masked else maskmaxabund}
        if numbands:
            return sketch.consume_seqfile_banding_with_mask, (parser, numbands, band, mask), kwargs
        return sketch.consume_seqfile_with_mask, (parser, mask), kwargs
    if numbands:
        return sketch.consume_seqfile_banding, (parser, numbands, band), {}
    return sketch.consume_seqfile, (parser,), {}


def load_sample_seqfile(seqfiles, ksize, memory, maxfpr=0.2, count=True, smallcount=False, mask=None,
                        maskmaxabund=0, consume_masked=False, numbands=None, band=None, outfile=None,
                        numthreads=1):
    """Count the k-mers of one sample's FASTA/FASTQ files into a new 4-table sketch sized from
    `memory` (kevlar/count.py:18-99).  With a `mask`, only k-mers absent from it are counted
    (or only those present, with `consume_masked`)."""
    numtables = 4
    sketchtype = 'nodegraph'
    if count:
        sketchtype = 'smallcountgraph' if smallcount else 'countgraph'
    tablesize = memory / numtables * khmer._buckets_per_byte[sketchtype]
    sketch = allocate(ksize, tablesize, num_tables=numtables, count=count, smallcount=smallcount)
    numreads = 0
    for seqfile in seqfiles:
        kevlar_b200.plog('[kevlar::count]', '- processing "{}"'.format(seqfile))
        parser = khmer.ReadParser(seqfile)
        workers = []
        for _ in range(numthreads):
            target, args, kwargs = _consumer(sketch, parser, mask, maskmaxabund, consume_masked, numbands, band)
            worker = threading.Thread(target=target, args=args, kwargs=kwargs)
            workers.append(worker)
            worker.start()
        for worker in workers:
            worker.join()
        numreads += parser.num_reads

    message = 'Done loading k-mers'
    if numbands:
        message += ' (band {:d}/{:d})'.format(band + 1, numbands)
    fpr = kevlar_b200.sketch.estimate_fpr(sketch)
    message += ';\n    {:d} reads processed'.format(numreads)
    message += ', {:d} distinct k-mers stored'.format(sketch.n_unique_kmers())
    message += ';\n    estimated false positive rate is {:1.3f}'.format(fpr)
    if fpr > maxfpr:
        message += ' (FPR too high, bailing out!!!)'
        raise kevlar_b200.sketch.KevlarUnsuitableFPRError('[kevlar::count] ' + message)

    if outfile:
        extensions = get_extension(count=count, smallcount=smallcount)
        if not outfile.endswith(extensions):
            outfile += extensions[1]
        sketch.save(outfile)
        message += ';\n    saved to "{:s}"'.format(outfile)
    kevlar_b200.plog('[kevlar::count]', message)
    return sketch


def print_config(args):
    kind = {1: 'node', 4: 'small count', 8: 'count'}[args.counter_size]
    message = 'Storing k-mers in a {} table'.format(kind)
    if args.counter_size == 1:
        message += ' (Bloom filter) for k-mer presence/absence queries'
    else:
        message += ', a CountMin sketch with a counter size of {} bits'.format(args.counter_size)
        message += ', for k-mer abundance queries (max abundance {})'.format({4: 15, 8: 255}[args.counter_size])
    kevlar_b200.plog('[kevlar::count]', message)


def main(args):
    if (args.num_bands is None) is not (args.band is None):
        raise ValueError('Must specify --num-bands and --band together')
    myband = args.band - 1 if args.band else None
    if args.mask:
        args.mask = kevlar_b200.sketch.load(args.mask)
    print_config(args)

    timer = kevlar_b200.Timer()
    timer.start()
    load_sample_seqfile(
        args.seqfile, args.ksize, args.memory, args.max_fpr, count=args.counter_size > 1,
        smallcount=args.counter_size == 4, mask=args.mask, consume_masked=args.count_masked,
        numbands=args.num_bands, band=myband, numthreads=args.threads, outfile=args.counttable,
    )
    kevlar_b200.plog('[kevlar::count] Total time: {:.2f} seconds'.format(timer.stop()))
