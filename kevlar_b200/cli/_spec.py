"""Declarative description of the command-line surface.

The flags, metavars, defaults and types are the reference's (kevlar/cli/count.py:42-80,
kevlar/cli/novel.py:65-155, kevlar/cli/filter.py:23-52, kevlar/cli/unband.py,
kevlar/cli/dist.py:14-46) because scripts and
workflows depend on them; everything is data here and `build` turns a spec into an argparse
sub-parser."""
import argparse

from kevlar_b200.khmer import khmer_args


def opt(*flags, **kwargs):
    return flags, kwargs


KSIZE = opt('-k', '--ksize', type=int, default=31, metavar='K', help='k-mer size; default is 31')
THREADS = opt('-t', '--threads', type=int, default=1, metavar='T',
              help='host threads feeding read batches to the GPU; default is 1')
NUM_BANDS = opt('--num-bands', type=int, metavar='N', default=None,
                help='split the hashed k-mer space into N bands')
BAND = opt('--band', type=int, metavar='I', default=None, help='which band (1..N) this run processes')
CTRL_MAX = opt('-x', '--ctrl-max', metavar='X', type=int, default=1,
               help='a k-mer seen more than X times in any control is not novel; default 1')
CASE_MIN = opt('-y', '--case-min', metavar='Y', type=int, default=6,
               help='a k-mer seen fewer than Y times in any case is not novel; default 6')
OUT = opt('-o', '--out', metavar='FILE', help='where to write the result; default is stdout')

COUNT = {
    'name': 'count',
    'description': 'Build a k-mer abundance sketch of one sample on the GPU (supports k-mer banding and masks).',
    'groups': [(None, [
        KSIZE,
        opt('-c', '--counter-size', type=int, choices=(1, 4, 8), metavar='C', default=8,
            help='counter width in bits: 1 (Bloom filter), 4 (max 15) or 8 (max 255, default)'),
        opt('-M', '--memory', type=khmer_args.memory_setting, default=1e6, metavar='MEM',
            help='bytes of sketch to allocate, e.g. 500M or 12G'),
        opt('--max-fpr', type=float, default=0.2, metavar='FPR',
            help='give up if the estimated false positive rate exceeds FPR; default 0.2'),
        opt('--mask', metavar='MSK', help='sketch of k-mers to leave out of the count'),
        opt('--count-masked', action='store_true', help='count ONLY the k-mers present in the mask'),
        NUM_BANDS, BAND, THREADS,
        opt('counttable', type=str,
            help='output sketch file; the extension matching the sketch type is appended when missing'),
        opt('seqfile', type=str, nargs='+', help='FASTA/FASTQ input, plain or gzipped'),
    ])],
}

NOVEL = {
    'name': 'novel',
    'description': 'Report the case reads holding k-mers that are abundant in every case sample and (nearly) '
                   'absent from every control sample.',
    'add_help': False,
    'groups': [
        ('Case/control config', [
            opt('--case', metavar='F', nargs='+', required=True, action='append',
                help='reads of one case sample; give the flag once per case sample'),
            opt('--case-counts', metavar='F', nargs='+', help='precomputed sketch per case sample'),
            opt('--control', metavar='F', nargs='+', action='append',
                help='reads of one control sample; give the flag once per control'),
            opt('--control-counts', metavar='F', nargs='+', help='precomputed sketch per control sample'),
            CTRL_MAX, CASE_MIN,
            opt('-M', '--memory', default='1e6', type=khmer_args.memory_setting, metavar='MEM',
                help='bytes of sketch per sample when counting from reads; default 1M'),
            opt('--max-fpr', type=float, default=0.2, metavar='FPR',
                help='give up if any sample\'s estimated false positive rate exceeds FPR; default 0.2'),
        ]),
        ('K-mer banding', [NUM_BANDS, BAND]),
        ('Output settings', [
            OUT,
            opt('--save-case-counts', metavar='CT', nargs='+', help='also save the case sketches here'),
            opt('--save-ctrl-counts', metavar='CT', nargs='+', help='also save the control sketches here'),
        ]),
        ('Miscellaneous settings', [
            opt('-h', '--help', action='help', help='show this help message and exit'),
            KSIZE,
            opt('--abund-screen', type=int, default=None, metavar='INT',
                help='drop a read outright if one of its k-mers is rarer than INT in a case sample'),
            THREADS,
            opt('--skip-until', type=str, metavar='ID', help='ignore case reads up to and including read ID'),
        ]),
    ],
}

FILTER = {
    'name': 'filter',
    'description': 'Recount the annotated k-mers of a `novel` output (minus a mask) and drop k-mers and reads '
                   'that no longer pass the thresholds.',
    'groups': [(None, [
        opt('-M', '--memory', type=khmer_args.memory_setting, default=1e6, metavar='MEM',
            help='bytes of sketch for the recount'),
        opt('--max-fpr', type=float, default=0.01, metavar='FPR',
            help='give up if the recount\'s estimated false positive rate exceeds FPR; default 0.01'),
        opt('--mask', metavar='MSK', help='sketch of k-mers to leave out of the recount'),
        CTRL_MAX, CASE_MIN, OUT,
        opt('augfastq', help='augmented FASTQ written by `novel`'),
    ])],
}

UNBAND = {
    'name': 'unband',
    'description': 'Merge the outputs of banded `novel` runs into one record per read with all its annotations.',
    'groups': [(None, [
        opt('-n', '--n-batches', metavar='N', type=int, default=16,
            help='number of temporary batches the records are spread over by read name; default 16'),
        OUT,
        opt('infile', nargs='+', help='augmented FASTA/FASTQ files'),
    ])],
}

DIST = {
    'name': 'dist',
    'description': 'Abundance distribution of the k-mers a mask selects (e.g. single-copy exonic k-mers): '
                   'mean, standard deviation and optionally the whole table.',
    'groups': [(None, [
        OUT, KSIZE,
        opt('-M', '--memory', type=khmer_args.memory_setting, default=1e6, metavar='MEM',
            help='bytes of sketch for the masked count'),
        THREADS,
        opt('-p', '--plot', metavar='PNG', help='draw the distribution into PNG (needs matplotlib)'),
        opt('--tsv', metavar='TSV', help='write the distribution as a tab-separated table'),
        opt('--plot-xlim', metavar=('MIN', 'MAX'), type=int, nargs=2, default=(0, 100),
            help='abundance range of the plot; default `0 100`'),
        opt('mask', help='nodetable of the k-mers to count'),
        opt('infiles', nargs='+', help='FASTA/FASTQ input, plain or gzipped'),
    ])],
}


def build(subparsers, spec):
    parser = subparsers.add_parser(spec['name'], description=spec['description'], add_help=spec.get('add_help', True),
                                   formatter_class=argparse.RawDescriptionHelpFormatter)
    for title, options in spec['groups']:
        target = parser.add_argument_group(title) if title else parser
        for flags, kwargs in options:
            target.add_argument(*flags, **kwargs)
    return parser
