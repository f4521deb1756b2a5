"""`kevlar unband` arguments (kevlar/cli/unband.py)."""


def subparser(subparsers):
    desc = ('Consolidate the outputs of banded `kevlar novel` runs: reads that appear in several band outputs '
            'become one record carrying all their novel k-mer annotations.')
    sub = subparsers.add_parser('unband', description=desc)
    sub.add_argument('-n', '--n-batches', metavar='N', type=int, default=16,
                     help='number of temporary batches the records are split into by read name; default is 16')
    sub.add_argument('-o', '--out', metavar='FILE', help='output file; default is terminal (stdout)')
    sub.add_argument('infile', nargs='+', help='input files in augmented Fasta/Fastq format')
