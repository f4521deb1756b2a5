"""Command-line front end: the `count`, `novel`, `filter`, `unband` and `dist` subcommands with the
flags, defaults and dispatch of kevlar/cli/__init__.py:31-108."""
import argparse
import sys

import kevlar_b200
from . import count, novel, filter, unband, dist

mains = {
    'count': kevlar_b200.count.main,
    'novel': kevlar_b200.novel.main,
    'filter': kevlar_b200.filter.main,
    'unband': kevlar_b200.unband.main,
    'dist': kevlar_b200.dist.main,
}

subparser_funcs = {
    'count': count.subparser,
    'novel': novel.subparser,
    'filter': filter.subparser,
    'unband': unband.subparser,
    'dist': dist.subparser,
}


def parser():
    banner = 'kevlar k-mer hot path on B200: reference-free variant discovery, GPU sketch construction'
    commands = '", "'.join(sorted(mains))
    top = argparse.ArgumentParser(description=banner, formatter_class=argparse.RawDescriptionHelpFormatter)
    top._positionals.title = 'Subcommands'
    top._optionals.title = 'Global arguments'
    top.add_argument('-v', '--version', action='version', version='kevlar v{}'.format(kevlar_b200.__version__))
    top.add_argument('-l', '--logfile', metavar='F', help='log file for diagnostic messages, warnings, and errors')
    top.add_argument('--tee', action='store_true', help='write diagnostic output to logfile AND terminal (stderr)')
    subparsers = top.add_subparsers(dest='cmd', metavar='cmd', help='"' + commands + '"')
    for func in subparser_funcs.values():
        func(subparsers)
    return top


def parse_args(arglist=None):
    args = parser().parse_args(arglist)
    kevlar_b200.logstream = sys.stderr
    if args.logfile and args.logfile != '-':
        kevlar_b200.logstream = kevlar_b200.open(args.logfile, 'w')
    kevlar_b200.teelog = args.tee
    return args
