"""`kevlar filter`: recount the annotated k-mers of a `novel` output (minus a mask) and drop
what no longer passes the thresholds -- same interface and log lines as kevlar/filter.py.
The per-k-mer `add`/`get` calls of the reference are batched into a handful of GPU calls."""
import kevlar_b200
from kevlar_b200 import khmer
from kevlar_b200.sequence import KmerOfInterest

CHUNK_KMERS = 1 << 20


def _recount(counts, mask, kmers):
    if mask:
        present = mask.get_many(kmers)
        kmers = [km for km, c in zip(kmers, present) if not c > 0]
    counts.add_many(kmers)


def first_pass(reads, mask, memory, timer):
    """Count every annotated k-mer that is not in the mask into a fresh Counttable whose k comes
    from the first annotation seen (kevlar/filter.py:15-39)."""
    kevlar_b200.plog('[kevlar::filter] First pass: re-counting k-mers')
    timer.start('firstpass')
    counts = None
    progress = kevlar_b200.ProgressIndicator('[kevlar::filter]     processed {counter} reads', interval=1e5,
                                             breaks=[1e6, 1e7])
    pending = []
    for n, read in enumerate(reads, 1):
        progress.update()
        if len(read.annotations) == 0:
            continue
        if counts is None:
            counts = khmer.Counttable(read.annotations[0].ksize, memory / 4, 4)
        pending.extend(read.ikmerseq(ikmer) for ikmer in read.annotations)
        if len(pending) >= CHUNK_KMERS:
            _recount(counts, mask, pending)
            pending = []
    if pending:
        _recount(counts, mask, pending)
    elapsed = timer.stop('firstpass')
    message = 'First pass complete! Processed {:d} reads in {:.2f} seconds!'.format(n, elapsed)
    kevlar_b200.plog('[kevlar::filter]', message)
    return counts


def check_fpr(counts, maxfpr):
    fpr = kevlar_b200.sketch.estimate_fpr(counts)
    message = 'FPR for re-computed k-mer counts: {:1.3f}'.format(fpr)
    kevlar_b200.plog('[kevlar::filter]', message)
    if fpr > maxfpr:
        raise kevlar_b200.sketch.KevlarUnsuitableFPRError(message + 'FPR too high, bailing out!!!')


def _validate_chunk(chunk, counts, casemin, ctrlmax):
    """Second-pass logic for a list of reads: one batched lookup, then the per-k-mer rules
    (kevlar/filter.py:59-78)."""
    candidates = []   # (read index, ikmer) whose control abundances are fine
    for r, read in enumerate(chunk):
        for ikmer in read.annotations:
            if any(a > ctrlmax for a in ikmer.abund[1:]):
                continue
            candidates.append((r, ikmer))
    newcounts = counts.get_many([chunk[r].ikmerseq(ik) for r, ik in candidates]) if candidates else []
    validated = [[] for _ in chunk]
    for (r, ikmer), newcount in zip(candidates, newcounts):
        if newcount < casemin:
            continue
        abund = tuple([int(newcount)] + list(ikmer.abund[1:]))
        validated[r].append(KmerOfInterest(ikmer.ksize, ikmer.offset, abund))
    for read, kmers in zip(chunk, validated):
        if kmers:
            read.annotations = kmers
            yield read


def second_pass(reads, counts, casemin, ctrlmax, timer):
    """Replace each k-mer's case abundance by its recount and drop k-mers/reads that fail
    `casemin`/`ctrlmax` (kevlar/filter.py:51-82)."""
    kevlar_b200.plog('[kevlar::filter] Second pass: discarding k-mers/reads')
    timer.start('secondpass')
    kept = 0
    progress = kevlar_b200.ProgressIndicator('[kevlar::filter]     processed {counter} reads', interval=1e5,
                                             breaks=[1e6, 1e7])
    chunk, nk = [], 0
    for read in reads:
        progress.update()
        chunk.append(read)
        nk += len(read.annotations)
        if nk >= CHUNK_KMERS:
            for valid in _validate_chunk(chunk, counts, casemin, ctrlmax):
                yield valid
                kept += 1
            chunk, nk = [], 0
    for valid in _validate_chunk(chunk, counts, casemin, ctrlmax):
        yield valid
        kept += 1
    elapsed = timer.stop('secondpass')
    message = 'Second pass complete! Validated {:d} reads in {:.2f} seconds!'.format(kept, elapsed)
    kevlar_b200.plog('[kevlar::filter]', message)


def _records(readfile):
    return kevlar_b200.parse_augmented_fastx(kevlar_b200.open(readfile, 'r'))


def filter(readfile, mask=None, memory=1e6, maxfpr=0.01, casemin=6, ctrlmax=1):
    """Generator over the reads of `readfile` that keep at least one k-mer after the recount."""
    clock = kevlar_b200.Timer()
    clock.start()
    counts = first_pass(_records(readfile), mask, memory, clock)
    check_fpr(counts, maxfpr)
    yield from second_pass(_records(readfile), counts, casemin, ctrlmax, clock)
    kevlar_b200.plog('[kevlar::filter]', 'Total time: {:.2f} seconds'.format(clock.stop()))


def main(args):
    mask = kevlar_b200.sketch.load(args.mask)
    outstream = kevlar_b200.open(args.out, 'w')
    for record in filter(args.augfastq, mask=mask, memory=args.memory, maxfpr=args.max_fpr,
                         casemin=args.case_min, ctrlmax=args.ctrl_max):
        kevlar_b200.print_augmented_fastx(record, outstream)
