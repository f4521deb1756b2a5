"""Sequence records and the augmented FASTA/FASTQ text format.

Behavioural mirror of the reference's Cython module kevlar/sequence.pyx (Record :34-83,
copy_record :86-90, print_augmented_fastx :93-126, parse_augmented_fastx :133-178) and of the
format description in docs/formats.rst.  `format_hit_records` is the batch formatter the GPU
novel scan uses: it turns kv_hit rows straight into output text.
"""
from collections import namedtuple
import re

KmerOfInterest = namedtuple('KmerOfInterest', 'ksize offset abund')

_PAIRS = {'A': 'T', 'T': 'A', 'U': 'A', 'G': 'C', 'C': 'G', 'Y': 'R', 'R': 'Y', 'S': 'S', 'W': 'W',
          'K': 'M', 'M': 'K', 'B': 'V', 'D': 'H', 'H': 'D', 'V': 'B', 'N': 'N'}
_COMPLEMENT = str.maketrans(''.join(_PAIRS) + ''.join(_PAIRS).lower(), ''.join(_PAIRS.values()) * 2)

MARGIN = ' ' * 10


def revcom(sequence):
    """Reverse complement (IUPAC aware, result upper-case like the reference's table)."""
    return sequence.translate(_COMPLEMENT)[::-1]


class Record(object):
    """A read plus its annotated "interesting" k-mers and mate sequences."""
    __slots__ = ('name', 'sequence', 'quality', 'annotations', 'mates', 'ikmers')

    def __init__(self, name, sequence, quality=None, annotations=None, mates=None, ikmers=None):
        self.name = name
        self.sequence = sequence
        self.quality = quality
        self.mates = [] if mates is None else mates
        self.ikmers = {}
        if annotations is None:
            self.annotations = []
        else:
            self.annotations = annotations
            if ikmers is not None:
                self.ikmers = ikmers
            else:
                for ikmer in annotations:
                    seq = self.ikmerseq(ikmer)
                    self.ikmers[seq] = ikmer
                    self.ikmers[revcom(seq)] = ikmer

    def __len__(self):
        return len(self.sequence)

    @property
    def id(self):
        return self.name.split()[0]

    def add_mate(self, mateseq):
        self.mates.append(mateseq)

    def ikmerseq(self, ikmer):
        return self.sequence[ikmer.offset:ikmer.offset + ikmer.ksize]

    def annotate(self, sequence, offset, abundances):
        inplace = self.sequence[offset:offset + len(sequence)]
        assert inplace == sequence, (inplace, sequence)
        ikmer = KmerOfInterest(len(sequence), offset, abundances)
        self.annotations.append(ikmer)
        self.ikmers[sequence] = ikmer
        self.ikmers[revcom(sequence)] = ikmer


def copy_record(record):
    quality = getattr(record, 'quality', None)
    return Record(record.name, record.sequence, quality)


def _head_text(name, sequence, quality):
    if quality is not None:
        return '@' + name + '\n' + sequence + '\n+\n' + quality + '\n'
    return '>' + name + '\n' + sequence + '\n'


def augmented_fastx_text(record):
    parts = [_head_text(record.name, record.sequence, record.quality)]
    seq = record.sequence
    for ikmer in sorted(record.annotations, key=lambda ik: ik.offset):
        parts.append(' ' * ikmer.offset + seq[ikmer.offset:ikmer.offset + ikmer.ksize] + MARGIN +
                     ' '.join(str(a) for a in ikmer.abund) + '#\n')
    for mate in record.mates:
        parts.append('#mateseq=' + mate + '#\n')
    return ''.join(parts)


def print_augmented_fastx(record, outstream):
    text = augmented_fastx_text(record)
    try:
        outstream.write(bytes(text, 'ascii'))
    except TypeError:
        outstream.write(text)


def write_record(record, outstream):
    print_augmented_fastx(record, outstream)


_MATE = re.compile(r'^#mateseq=(\S+)#\n$')


def parse_augmented_fastx(instream):
    """Generator over the records of an augmented FASTA/FASTQ stream."""
    record = None
    for line in instream:
        if line.strip() == '':
            continue
        if line[0] in '@>':
            if record is not None:
                yield record
            name = line[1:].strip()
            sequence = next(instream).strip()
            quality = None
            if line[0] == '@':
                next(instream)
                quality = next(instream).strip()
            record = Record(name=name, sequence=sequence, quality=quality)
        elif line.endswith('#\n'):
            if line.startswith('#mateseq='):
                record.add_mate(_MATE.search(line).group(1))
                continue
            offset = len(line) - len(line.lstrip())
            fields = re.split(r'\s+', line.strip()[:-1])
            record.annotate(fields[0], offset, tuple(int(a) for a in fields[1:]))
        else:
            raise Exception(line)
    yield record
