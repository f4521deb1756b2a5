"""`kevlar dist` arguments: see the DIST table in cli/_spec.py."""
from kevlar_b200.cli import _spec


def subparser(subparsers):
    return _spec.build(subparsers, _spec.DIST)
