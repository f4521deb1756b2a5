"""The one piece of ``khmer.khmer_args`` the kevlar CLI uses (kevlar/cli/count.py:49,
kevlar/cli/novel.py:100, kevlar/cli/filter.py:24)."""

_MULTIPLIER = {'K': 1e3, 'M': 1e6, 'G': 1e9, 'T': 1e12}


def memory_setting(label):
    """Parse ``10K`` / ``1M`` / ``8G`` / ``1e7`` / ``97`` into a number of bytes (float)."""
    try:
        return float(label)
    except ValueError:
        pass
    number, suffix = label[:-1], label[-1:].upper()
    if suffix not in _MULTIPLIER:
        raise ValueError('cannot parse memory setting "{}"'.format(label))
    return float(number) * _MULTIPLIER[suffix]
