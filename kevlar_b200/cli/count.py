"""`kevlar count` arguments (kevlar/cli/count.py:42-80)."""
import argparse

from kevlar_b200.khmer import khmer_args


def subparser(subparsers):
    desc = 'Compute k-mer abundances for the provided sample on the GPU. Supports k-mer banding.'
    sub = subparsers.add_parser('count', description=desc, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub.add_argument('-k', '--ksize', type=int, default=31, metavar='K', help='k-mer size; default is 31')
    sub.add_argument('-c', '--counter-size', type=int, choices=(1, 4, 8), metavar='C', default=8,
                     help='bits per counter: 1 (max count 1), 4 (max 15) or 8 (max 255); default is 8')
    sub.add_argument('-M', '--memory', type=khmer_args.memory_setting, default=1e6, metavar='MEM',
                     help='memory to allocate for the count table')
    sub.add_argument('--max-fpr', type=float, default=0.2, metavar='FPR',
                     help='terminate if the estimated false positive rate is higher than "FPR"; default is 0.2')
    sub.add_argument('--mask', metavar='MSK', help='counttable or nodetable of k-mers to ignore when counting')
    sub.add_argument('--count-masked', action='store_true',
                     help='invert the mask: count only k-mers that ARE in the mask')
    sub.add_argument('--num-bands', type=int, metavar='N', default=None,
                     help='number of bands into which to divide the hashed k-mer space')
    sub.add_argument('--band', type=int, metavar='I', default=None,
                     help='a number between 1 and N (inclusive) indicating the band to be processed')
    sub.add_argument('-t', '--threads', type=int, default=1, metavar='T',
                     help='number of host threads feeding the GPU; default is 1')
    sub.add_argument('counttable', type=str,
                     help='output file; ".counttable" (or the matching type extension) is appended when missing')
    sub.add_argument('seqfile', type=str, nargs='+', help='input files in Fastq/Fasta format')
