"""`kevlar filter` arguments (kevlar/cli/filter.py:23-52)."""
from kevlar_b200.khmer import khmer_args


def subparser(subparsers):
    desc = ('Discard k-mers and reads that are contaminant in origin or whose abundances were inflated during '
            'the preliminary k-mer counting stage.')
    sub = subparsers.add_parser('filter', description=desc)
    sub.add_argument('-M', '--memory', type=khmer_args.memory_setting, default=1e6, metavar='MEM',
                     help='memory to allocate for the k-mer re-counting')
    sub.add_argument('--max-fpr', type=float, default=0.01, metavar='FPR',
                     help='terminate if the FPR of the recomputed abundances exceeds FPR; default is 0.01')
    sub.add_argument('--mask', metavar='MSK', help='counttable or nodetable of k-mers to ignore when re-counting')
    sub.add_argument('-x', '--ctrl-max', metavar='X', type=int, default=1,
                     help='k-mers with abund > X in any control sample are uninteresting; default is X=1')
    sub.add_argument('-y', '--case-min', metavar='Y', type=int, default=6,
                     help='k-mers with abund < Y in any case sample are uninteresting; default is Y=6')
    sub.add_argument('-o', '--out', metavar='FILE', help='output file; default is terminal (stdout)')
    sub.add_argument('augfastq', help='putatively novel reads in augmented Fastq format')
