"""`kevlar novel`: keep the case reads that carry k-mers abundant in every case sample and
(nearly) absent from every control -- same interface, thresholds, quirks and log lines as
kevlar/novel.py, but the per-read / per-k-mer Python loop of the reference is ONE fused CUDA
kernel per batch of reads (kv_novel_batch); Python only formats the few reads that hit."""
import os

import numpy as np

import kevlar_b200
from kevlar_b200 import _lib, khmer
from kevlar_b200.fastx import batch_from_sequences
from kevlar_b200.sequence import Record

NOVEL_BATCH_BASES = 64 << 20


class KevlarCaseSampleMismatchError(ValueError):
    pass


def kmer_is_interesting(kmer, casecounts, controlcounts, case_min=5, ctrl_max=1, screen_thresh=None):
    """Single k-mer form of the novelty test (kevlar/novel.py:21-53).  Returns
    (interesting, discard_read, case_abundances, control_abundances); evaluation stops at the
    first sample that disqualifies the k-mer."""
    caseabunds = []
    for sketch in casecounts:
        abund = sketch.get(kmer)
        if abund < case_min:
            return False, bool(screen_thresh and abund < screen_thresh), [], []
        caseabunds.append(abund)
    ctrlabunds = []
    for sketch in controlcounts:
        abund = sketch.get(kmer)
        if abund > ctrl_max:
            return False, False, [], []
        ctrlabunds.append(abund)
    return True, False, caseabunds, ctrlabunds


def load_samples(counttables=None, filelists=None, ksize=31, memory=1e6, maxfpr=0.2, numbands=None, band=None,
                 numthreads=1, outfilelist=None):
    """One sketch per sample: read from `counttables` when given (any reads are then ignored),
    otherwise counted from the files in `filelists` -- always into 8-bit Counttables
    (kevlar/novel.py:56-77) -- and optionally saved to `outfilelist`."""
    assert counttables or filelists
    if counttables:
        kevlar_b200.plog('[kevlar::novel]    INFO:',
                         'counttables for {:d} sample(s) provided, any corresponding FASTA/FASTQ input will be '
                         'ignored for computing k-mer abundances'.format(len(counttables)))
        return kevlar_b200.sketch.load_sketchfiles(counttables, maxfpr)
    counting = dict(maxfpr=maxfpr, numbands=numbands, band=band, numthreads=numthreads)
    samples = [kevlar_b200.count.load_sample_seqfile(files, ksize, memory, **counting) for files in filelists]
    if outfilelist:
        save_counts(outfilelist, samples)
    return samples


def save_counts(filelist, tablelist):
    """Write each sample's sketch; refuses (with a warning) when the counts do not line up."""
    if len(filelist) != len(tablelist):
        kevlar_b200.plog('[kevlar::novel] WARNING:',
                         'number of filenames provided ({:d})does not match the number of samples provided ({:d}); '
                         'stubbornly refusing to save k-mer counts'.format(len(filelist), len(tablelist)))
        return
    for outfile, counttable in zip(filelist, tablelist):
        if not outfile.endswith(('.ct', '.counttable')):
            outfile += '.counttable'
        kevlar_b200.plog('    saved to "{}"'.format(os.path.abspath(outfile)))
        counttable.save(outfile)


class ReadBatches(object):
    """Case reads of several files, iterable record by record like
    kevlar.multi_file_iter_khmer, but also able to hand `novel` whole batches."""

    def __init__(self, filenames):
        self.filenames = list(filenames)

    def __iter__(self):
        return kevlar_b200.multi_file_iter_khmer(self.filenames)

    def batches(self, max_bases=NOVEL_BATCH_BASES):
        for filename in self.filenames:
            for batch in khmer.ReadParser(filename).batches(max_bases, keep_text=True):
                yield batch


def _record_batches(stream, max_bases):
    """Group an arbitrary stream of record objects (.name .sequence .quality) into SeqBatches."""
    names, seqs, quals, total = [], [], [], 0
    for record in stream:
        names.append(record.name.encode('ascii'))
        seqs.append(record.sequence)
        quality = getattr(record, 'quality', None)
        quals.append(quality.encode('ascii') if quality is not None else None)
        total += len(record.sequence)
        if total >= max_bases:
            yield batch_from_sequences(seqs, names, quals)
            names, seqs, quals, total = [], [], [], 0
    if seqs:
        yield batch_from_sequences(seqs, names, quals)


def novel(casestream, casecounts, controlcounts, ksize=31, abundscreen=None, casemin=5, ctrlmax=0, numbands=None,
          band=None, skipuntil=None):
    """Generator of annotated Records for the reads with at least one novel k-mer
    (kevlar/novel.py:95-176).  `band` is the 0-based band; the band test is the reference's
    `(hash & (numbands - 1)) == band - 1`."""
    numbands_unset = not numbands
    band_unset = not band and band != 0
    if numbands_unset is not band_unset:
        raise ValueError('Must specify `numbands` and `band` together')
    if band is not None and band < 0:
        message = '`band` must be a value between 0 and {:d}'.format(numbands - 1)
        raise ValueError(message + ' (`numbands` - 1), inclusive')

    timer = kevlar_b200.Timer()
    timer.start()
    nkmers, nreads = 0, 0
    update_message = '[kevlar::novel]     processed {counter} reads'
    first_message = update_message
    if skipuntil:
        first_message += '; skipping reads in search of {read}'.format(read=skipuntil)
    progress = kevlar_b200.ProgressIndicator(first_message, interval=1e6, breaks=[1e7, 1e8, 1e9], usetimer=True)
    unique_kmers = set()
    nsamples = len(casecounts) + len(controlcounts)
    sketch_k = casecounts[0].ksize() if casecounts else ksize
    seen = 0

    if hasattr(casestream, 'batches'):
        batches = casestream.batches(NOVEL_BATCH_BASES)
    else:
        batches = _record_batches(casestream, NOVEL_BATCH_BASES)
    for batch in batches:
        if skipuntil:  # fast-forward: the matching read itself is skipped too (novel.py:125-132)
            at = batch.find_name(skipuntil.encode('ascii'))
            if at < 0:
                seen += len(batch)
                progress.update(len(batch))
                continue
            progress.update(at + 1)
            seen += at + 1
            message = 'Found read {:s} (skipped {:d} reads)'.format(skipuntil, seen)
            kevlar_b200.plog('[kevlar::novel]', message)
            skipuntil = False
            progress.message = update_message
            batch = batch.tail(at + 1)
            if len(batch) == 0:
                continue
        progress.update(len(batch))
        seen += len(batch)

        hits, flags, _ = khmer.novel_batch(casecounts, controlcounts, batch.bases, batch.offsets, casemin, ctrlmax,
                                           screen=abundscreen, num_bands=numbands,
                                           band_minus_1=(band - 1) if numbands else 0)
        if len(hits) == 0:
            continue
        reads, starts = np.unique(hits['read'], return_index=True)
        ends = list(starts[1:]) + [len(hits)]
        for read, lo, hi in zip(reads, starts, ends):
            source = batch.record(int(read))
            if len(source.sequence) < ksize:
                continue
            annotated = Record(source.name, source.sequence, source.quality)
            for hit in hits[lo:hi]:
                offset = int(hit['offset'])
                kmer = source.sequence[offset:offset + sketch_k]
                annotated.annotate(kmer, offset, tuple(int(a) for a in hit['abund'][:nsamples]))
                unique_kmers.add(kevlar_b200.revcommin(kmer))
            if flags[read] & _lib.READ_DISCARDED:
                continue   # its earlier k-mers still count as seen (novel.py:152-162)
            nreads += 1
            nkmers += len(annotated.annotations)
            yield annotated

    elapsed = timer.stop()
    message = 'Found {:d} instances'.format(nkmers)
    message += ' of {:d} unique novel kmers'.format(len(unique_kmers))
    message += ' in {:d} reads'.format(nreads)
    message += ' in {:.2f} seconds'.format(elapsed)
    kevlar_b200.plog('[kevlar::novel]', message)


def main(args):
    clock = kevlar_b200.Timer()
    clock.start()
    if (not args.num_bands) is not (not args.band):
        raise ValueError('Must specify --num-bands and --band together')
    band = args.band - 1 if args.band else None   # 1-based on the command line
    shared = (args.ksize, args.memory, args.max_fpr, args.num_bands, band, args.threads)

    def load(what, key, sketchfiles, readfiles, savefiles, done):
        kevlar_b200.plog('[kevlar::novel] Loading {} samples'.format(what))
        clock.start(key)
        sketches = load_samples(sketchfiles, readfiles, *shared, savefiles)
        kevlar_b200.plog(done.format(clock.stop(key)))
        return sketches

    clock.start('loadall')
    controls = load('control', 'loadctrl', args.control_counts, args.control, args.save_ctrl_counts,
                    '[kevlar::novel] Control samples loaded in {:.2f} sec')
    cases = load('case', 'loadcases', args.case_counts, args.case, args.save_case_counts,
                 '[kevlar::novel] Case samples loaded in {:.2f} sec')
    kevlar_b200.plog('[kevlar::novel] All samples loaded in {:.2f} sec'.format(clock.stop('loadall')))

    clock.start('iter')
    kevlar_b200.plog('[kevlar::novel]', 'Iterating over reads from {:d} case sample(s)'.format(len(args.case)))
    outstream = kevlar_b200.open(args.out, 'w')
    casereads = ReadBatches(f for filelist in args.case for f in filelist)
    for annotated in novel(casereads, cases, controls, ksize=args.ksize, abundscreen=args.abund_screen,
                           casemin=args.case_min, ctrlmax=args.ctrl_max, numbands=args.num_bands, band=band,
                           skipuntil=args.skip_until):
        kevlar_b200.print_augmented_fastx(annotated, outstream)
    kevlar_b200.plog('[kevlar::novel]', 'Iterated over all case reads in {:.2f} seconds'.format(clock.stop('iter')))
    kevlar_b200.plog('[kevlar::novel]', 'Total time: {:.2f} seconds'.format(clock.stop()))
