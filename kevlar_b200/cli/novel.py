"""`kevlar novel` arguments (kevlar/cli/novel.py:65-155)."""
import argparse

from kevlar_b200.khmer import khmer_args


def subparser(subparsers):
    desc = ('Identify "interesting" (potentially novel) k-mers -- abundant in every case sample, effectively '
            'absent from every control sample -- and output the reads that contain them.')
    sub = subparsers.add_parser('novel', description=desc, add_help=False,
                                formatter_class=argparse.RawDescriptionHelpFormatter)

    samp = sub.add_argument_group('Case/control config')
    samp.add_argument('--case', metavar='F', nargs='+', required=True, action='append',
                      help='FASTA/FASTQ file(s) of one case sample; repeat the flag for more case samples')
    samp.add_argument('--case-counts', metavar='F', nargs='+',
                      help='pre-computed counttable file(s), one per case sample')
    samp.add_argument('--control', metavar='F', nargs='+', action='append',
                      help='FASTA/FASTQ file(s) of one control sample; repeat the flag for more controls')
    samp.add_argument('--control-counts', metavar='F', nargs='+',
                      help='pre-computed counttable file(s), one per control sample')
    samp.add_argument('-x', '--ctrl-max', metavar='X', type=int, default=1,
                      help='k-mers with abund > X in any control sample are uninteresting; default is X=1')
    samp.add_argument('-y', '--case-min', metavar='Y', type=int, default=6,
                      help='k-mers with abund < Y in any case sample are uninteresting; default is Y=6')
    samp.add_argument('-M', '--memory', default='1e6', type=khmer_args.memory_setting, metavar='MEM',
                      help='memory for the k-mer abundances of each sample; default is 1M')
    samp.add_argument('--max-fpr', type=float, default=0.2, metavar='FPR',
                      help='terminate if the expected false positive rate of any sample exceeds FPR; default 0.2')

    band = sub.add_argument_group('K-mer banding')
    band.add_argument('--num-bands', type=int, metavar='N', default=None,
                      help='number of bands into which to divide the hashed k-mer space')
    band.add_argument('--band', type=int, metavar='I', default=None,
                      help='a number between 1 and N (inclusive) indicating the band to be processed')

    out = sub.add_argument_group('Output settings')
    out.add_argument('-o', '--out', metavar='FILE', help='output file; default is terminal (stdout)')
    out.add_argument('--save-case-counts', metavar='CT', nargs='+',
                     help='save the computed k-mer counts of each case sample to these files')
    out.add_argument('--save-ctrl-counts', metavar='CT', nargs='+',
                     help='save the computed k-mer counts of each control sample to these files')

    misc = sub.add_argument_group('Miscellaneous settings')
    misc.add_argument('-h', '--help', action='help', help='show this help message and exit')
    misc.add_argument('-k', '--ksize', type=int, default=31, metavar='K', help='k-mer size; default is 31')
    misc.add_argument('--abund-screen', type=int, default=None, metavar='INT',
                      help='discard reads with any k-mers whose abundance is < INT')
    misc.add_argument('-t', '--threads', type=int, default=1, metavar='T',
                      help='number of host threads feeding the GPU while counting; default is 1')
    misc.add_argument('--skip-until', type=str, metavar='ID',
                      help='skip all case reads until the read named ID has been seen')
