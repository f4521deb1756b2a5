"""`kevlar unband`: merge the per-band outputs of banded `kevlar novel` runs into one record
per read carrying all its annotations (kevlar/unband.py:26-78).

The reference spills records into N temporary files keyed by Python's salted hash() of the read
name -- so which batch a read lands in differs from run to run -- and emits each batch sorted
by name.  Here the batch key is a stable CRC32 of the name: outputs are reproducible and, like
the reference's, sorted by read name within each batch.
"""
from tempfile import TemporaryDirectory
import zlib

import kevlar_b200


def batch_of(name, numbatches):
    """Deterministic stand-in for the reference's `hash(record.name) % numbatches`
    (kevlar/unband.py:36), which Python salts per process."""
    return zlib.crc32(name.encode('utf-8')) % numbatches


def create_batch_files(numbatches, tempdir):
    """One gzipped spill file per batch (kevlar/unband.py:15-23)."""
    return [kevlar_b200.open('{:s}/kevlar-unband-batch{:d}.augfastq.gz'.format(tempdir, i), 'w')
            for i in range(numbatches)]


def write_records_to_batches(recordstream, batchfiles):
    """Spread the records over the spill files by read name, so that all copies of a read --
    one per band it was reported in -- meet in the same file (kevlar/unband.py:26-38)."""
    kevlar_b200.plog('[kevlar::unband]', 'writing records to {:d} temp batch files'.format(len(batchfiles)))
    progress = kevlar_b200.ProgressIndicator('[kevlar::unband]     processed {counter} reads', interval=1e5,
                                             breaks=[1e6, 1e7])
    for record in recordstream:
        progress.update()
        kevlar_b200.print_augmented_fastx(record, batchfiles[batch_of(record.name, len(batchfiles))])


def resolve_batch(batchfile):
    """Re-read one spill file, fold the copies of each read into the first one seen, emit the reads
    in name order with their annotations in offset order (kevlar/unband.py:41-58)."""
    filename = batchfile.name
    batchfile.close()
    merged = {}
    with kevlar_b200.open(filename, 'r') as fh:
        for read in kevlar_b200.parse_augmented_fastx(fh):
            if read is None:
                continue
            if read.name in merged:
                merged[read.name].annotations.extend(read.annotations)
            else:
                merged[read.name] = read
    for name in sorted(merged):
        merged[name].annotations.sort(key=lambda ikmer: ikmer.offset)
        yield merged[name]


def resolve_batches(batchfiles):
    kevlar_b200.plog('[kevlar::unband]', 'resolving duplicate reads in {:d} batches'.format(len(batchfiles)))
    for n, batchfile in enumerate(batchfiles):
        for read in resolve_batch(batchfile):
            yield read
        kevlar_b200.plog('[kevlar::unband]     batch {:d} complete'.format(n))
    kevlar_b200.plog('[kevlar::unband] Done!')


def unband(recordstream, numbatches=16):
    """kevlar/unband.py:72-77 (the temporary directory goes away with the generator)."""
    with TemporaryDirectory() as tempdir:
        batchfiles = create_batch_files(numbatches, tempdir)
        write_records_to_batches(recordstream, batchfiles)
        for read in resolve_batches(batchfiles):
            yield read


def afxstream(filenames):
    """All records of several augmented FASTA/FASTQ files, one stream (the record source of kevlar/unband.py:80-90)."""
    for filename in filenames:
        with kevlar_b200.open(filename, 'r') as fh:
            for record in kevlar_b200.parse_augmented_fastx(fh):
                if record is not None:
                    yield record


def main(args):
    outstream = kevlar_b200.open(args.out, 'w')
    for read in unband(afxstream(args.infile), args.n_batches):
        kevlar_b200.print_augmented_fastx(read, outstream)
