"""Synthetic trio reads for measurement.

Not product code: a vectorised stand-in for the reference's data recipe -- `kevlar gentrio`
(kevlar/gentrio.py:116-257: a random genome, inherited and de novo SNVs/indels, genotypes per
sample) followed by wgsim-style 100 bp reads with 0.5 % substitution errors
(kevlar/tests/data/minitrio/README).  Output is directly in the C-ABI batch layout
(concatenated bases + offsets), one batch per sample: (proband, mother, father).
"""
import numpy as np

LETTERS = np.frombuffer(b'ACGT', dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[LETTERS] = np.frombuffer(b'TGCA', dtype=np.uint8)

READ_SEEDS = (678678, 12893475, 5647348)   # proband, mother, father (minitrio/README)


def random_genome(length, seed=42):
    rng = np.random.Generator(np.random.PCG64(seed))
    return LETTERS[rng.integers(0, 4, size=length)]


def _apply_sequential(hap, variants):
    """Apply (pos, kind, payload) edits to a haplotype, right to left so positions stay valid."""
    for pos, kind, payload in sorted(variants, key=lambda v: -v[0]):
        if kind == 'snv':
            hap[pos] = payload
        elif kind == 'ins':
            hap = np.concatenate([hap[:pos], payload, hap[pos:]])
        else:
            hap = np.concatenate([hap[:pos], hap[pos + payload:]])
    return hap


def _apply(hap, variants):
    """`_apply_sequential` assembled from pieces in one pass (a 100 Mbp haplotype with hundreds of
    indels would otherwise be copied once per indel).  Identical to it whenever no variant lies
    inside the footprint of a deletion to its left (always the case for the 1 Mbp benchmark trio,
    tests/test_host.py checks it); such a variant is dropped here."""
    pieces, cursor = [], 0
    for pos, kind, payload in sorted(variants, key=lambda v: v[0]):
        if pos < cursor:
            continue
        pieces.append(hap[cursor:pos])
        if kind == 'snv':
            pieces.append(np.array([payload], dtype=np.uint8))
            cursor = pos + 1
        elif kind == 'ins':
            pieces.append(np.asarray(payload, dtype=np.uint8))
            cursor = pos
        else:
            cursor = pos + payload
    pieces.append(hap[cursor:])
    return np.concatenate(pieces)


def _draw_variants(rng, genome, n):
    out = []
    for _ in range(n):
        pos = int(rng.integers(1000, len(genome) - 1000))
        kind = rng.choice(['snv', 'ins', 'del'], p=[0.8, 0.1, 0.1])
        if kind == 'snv':
            alts = LETTERS[LETTERS != genome[pos]]
            out.append((pos, 'snv', alts[int(rng.integers(0, 3))]))
        elif kind == 'ins':
            out.append((pos, 'ins', LETTERS[rng.integers(0, 4, size=int(rng.integers(5, 351)))]))
        else:
            out.append((pos, 'del', int(rng.integers(5, 351))))
    return out


def trio_haplotypes(genome_len, n_inherited=None, n_denovo=None, seed=2018):
    """Two haplotypes each for (proband, mother, father).  Variant counts default to the
    gentrio defaults (20 inherited, 10 de novo) per Mbp."""
    scale = max(1, genome_len // 1000000)
    n_inherited = 20 * scale if n_inherited is None else n_inherited
    n_denovo = 10 * scale if n_denovo is None else n_denovo
    genome = random_genome(genome_len)
    rng = np.random.Generator(np.random.PCG64(seed))
    inherited = _draw_variants(rng, genome, n_inherited)
    denovo = _draw_variants(rng, genome, n_denovo)
    parent_sets = {('m', 0): [], ('m', 1): [], ('f', 0): [], ('f', 1): []}
    child_sets = [[], []]   # maternal, paternal haplotype of the proband
    for var in inherited:
        parent = 'm' if rng.random() < 0.5 else 'f'
        hap = int(rng.integers(0, 2))
        parent_sets[(parent, hap)].append(var)
        if rng.random() < 0.5:   # transmitted
            child_sets[0 if parent == 'm' else 1].append(var)
    for var in denovo:
        child_sets[int(rng.integers(0, 2))].append(var)
    mother = [_apply(genome.copy(), parent_sets[('m', h)]) for h in (0, 1)]
    father = [_apply(genome.copy(), parent_sets[('f', h)]) for h in (0, 1)]
    proband = [_apply(genome.copy(), child_sets[h]) for h in (0, 1)]
    return proband, mother, father


def sample_reads(haplotypes, n_reads, read_len=100, error_rate=0.005, seed=0):
    """n_reads reads of read_len from random haplotype/position/strand with iid substitution
    errors.  Returns (bases uint8[n*len], offsets uint64[n+1])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    which = rng.integers(0, len(haplotypes), size=n_reads)
    bases = np.empty((n_reads, read_len), dtype=np.uint8)
    span = np.arange(read_len)
    for h, hap in enumerate(haplotypes):
        rows = np.nonzero(which == h)[0]
        starts = rng.integers(0, len(hap) - read_len, size=len(rows))
        bases[rows] = hap[starts[:, None] + span[None, :]]
    minus = rng.random(n_reads) < 0.5
    bases[minus] = _COMP[bases[minus][:, ::-1]]
    errors = rng.random(bases.shape) < error_rate
    n_err = int(errors.sum())
    # substitute with a DIFFERENT base: rotate by 1..3 in ACGT order
    code = np.searchsorted(LETTERS, bases[errors])   # ACGT is sorted
    bases[errors] = LETTERS[(code + rng.integers(1, 4, size=n_err)) % 4]
    offsets = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len))
    return bases.reshape(-1), offsets


def simulate_trio(genome_len=1000000, coverage=30, read_len=100, error_rate=0.005, seed_offset=0,
                  reads_per_sample=None):
    """Reads for (proband, mother, father): list of (bases, offsets)."""
    haps = trio_haplotypes(genome_len)
    n_reads = reads_per_sample or int(coverage * genome_len / read_len)
    return [sample_reads(h, n_reads, read_len, error_rate, seed + seed_offset) for h, seed in zip(haps, READ_SEEDS)]


def device_trio(genome_len, reads_per_sample, rank=0, world=1, read_len=100, error_rate=0.005, device=None):
    """Reads for (proband, mother, father) drawn ON THE DEVICE (kv_synth_reads): list of
    (bases, offsets) torch tensors holding this rank's contiguous slice of every sample's
    `reads_per_sample` reads.  Read r of a sample is a pure function of (sample seed, r), so the
    union over ranks is the same read set for every world size.  For inputs the size of BASELINE
    configs 3-4, which a host generator plus PCIe would turn into the bottleneck."""
    import ctypes
    import torch
    from kevlar_b200 import _lib, multigpu
    dev_index = _lib.current_device() if device is None else int(device)
    dev = torch.device('cuda', dev_index)
    lo, hi = multigpu.shard_bounds(reads_per_sample, rank, world)
    out = []
    for haps, seed in zip(trio_haplotypes(genome_len), READ_SEEDS):
        dhaps = [torch.from_numpy(np.ascontiguousarray(h)).to(dev) for h in haps]
        ptrs = (ctypes.c_void_p * len(dhaps))(*[t.data_ptr() for t in dhaps])
        lens = (ctypes.c_uint64 * len(dhaps))(*[t.numel() for t in dhaps])
        n = hi - lo
        bases = torch.empty(n * read_len + 16, dtype=torch.uint8, device=dev)[:n * read_len]
        offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        _lib.check(_lib.lib().kv_synth_reads(dev_index, ptrs, lens, len(dhaps), n, lo, read_len, float(error_rate),
                                             int(seed), bases.data_ptr(), offsets.data_ptr()))
        out.append((bases, offsets))
        del dhaps
    return out
