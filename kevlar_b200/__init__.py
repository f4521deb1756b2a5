"""kevlar_b200 -- the `kevlar count` -> `kevlar novel` -> `kevlar filter` k-mer path on B200.

A from-scratch, GPU-native implementation of the one data-parallel hot path of
kevlar-dev/kevlar behind the reference's own Python API:

    kevlar_b200.khmer      drop-in for the slice of `khmer` the path uses (the boundary)
    kevlar_b200.sketch     allocate / load / autoload / estimate_fpr   (kevlar/sketch.py)
    kevlar_b200.count      load_sample_seqfile, main                   (kevlar/count.py)
    kevlar_b200.novel      novel, kmer_is_interesting, load_samples    (kevlar/novel.py)
    kevlar_b200.filter     filter, first_pass, second_pass             (kevlar/filter.py)
    kevlar_b200.sequence   Record, augmented FASTA/FASTQ I/O           (kevlar/sequence.pyx)
    kevlar_b200.cli        argparse front end, same flags              (kevlar/cli/*.py)
    kevlar_b200.multigpu   read sharding + sketch merge over NCCL/NVLink

All hashing and counting runs in libkvsketch.so (kevlar_b200/csrc, hand-written sm_100a
CUDA).  There is no CPU fallback.
"""
import builtins
from gzip import open as gzopen
import sys

from kevlar_b200 import _lib           # noqa: F401
from kevlar_b200 import khmer          # noqa: F401
from kevlar_b200.timer import Timer
from kevlar_b200.progress import ProgressIndicator
from kevlar_b200 import sequence
from kevlar_b200.sequence import parse_augmented_fastx, print_augmented_fastx, revcom
from kevlar_b200 import sketch
from kevlar_b200 import count
from kevlar_b200 import novel
from kevlar_b200 import filter
from kevlar_b200 import unband
from kevlar_b200 import dist
from kevlar_b200 import simlike
from kevlar_b200 import cli

__version__ = '0.1.0+b200'

logstream = None
teelog = False


def plog(*args, **kwargs):
    """Diagnostic output: to the log stream if one is set, else (or also, with --tee) to stderr
    (kevlar/__init__.py:76-81)."""
    if logstream is not None:
        print(*args, **kwargs, file=logstream)
    if logstream is None or teelog:
        print(*args, **kwargs, file=sys.stderr)


def open(filename, mode):
    """Plain or gzip text file by suffix; '-'/None mean stdin/stdout (kevlar/__init__.py:84-94)."""
    if mode not in ('r', 'w'):
        raise ValueError('invalid mode "{}"'.format(mode))
    if filename in ['-', None]:
        return sys.stdin if mode == 'r' else sys.stdout
    if filename.endswith('.gz'):
        return gzopen(filename, mode + 't')
    return builtins.open(filename, mode)


def revcommin(seq):
    """The lexicographically smaller of a sequence and its reverse complement."""
    rc = revcom(seq)
    return seq if seq <= rc else rc


def same_seq(seq1, seq2, seq2revcom=None):
    if seq2revcom is None:
        seq2revcom = revcom(seq2)
    return seq1 == seq2 or seq1 == seq2revcom


def multi_file_iter_khmer(filenames):
    """Records of several FASTA/FASTQ files, one after another (kevlar/__init__.py:125-128)."""
    for filename in filenames:
        for record in khmer.ReadParser(filename):
            yield record
