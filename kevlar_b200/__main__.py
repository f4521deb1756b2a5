"""`python -m kevlar_b200 <cmd> ...` (kevlar/__main__.py:14-30)."""
import kevlar_b200


def main(arglist=None):
    args = kevlar_b200.cli.parse_args(arglist)
    if args.cmd is None:
        kevlar_b200.cli.parser().parse_args(['-h'])
    assert args.cmd in kevlar_b200.cli.mains
    kevlar_b200.plog('[kevlar] running version {}'.format(kevlar_b200.__version__))
    kevlar_b200.cli.mains[args.cmd](args)


if __name__ == '__main__':
    main()
