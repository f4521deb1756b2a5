"""Sketch queries of `kevlar simlike` (kevlar/simlike.py:22-96).

Only the part of simlike that touches the sketches lives here: the abundances of the k-mers
spanning a call's alternate allele in the case, control and reference sketches.  The likelihood
arithmetic on those few numbers (kevlar/simlike.py:99-200) and the VCF plumbing stay with the
reference; they take the lists returned here unchanged.

The reference asks one sketch for one window at a time (`get_kmer_counts`).  Here all windows of
a run go to each sketch in ONE `kv_kmer_counts_batch` call (`spanning_kmer_abundances_many`);
the single-window function is the same code on a list of one.
"""
import numpy as np


def _valid_positions(refr_counts_of_alt):
    return refr_counts_of_alt == 0


def discard_nonunique_kmers(altseq, case, controls, refr):
    """kevlar/simlike.py:22-35: drop alt-allele k-mers that also occur in the reference genome."""
    alt_counts_refr = refr.get_kmer_counts_many([altseq])[0]
    keep = _valid_positions(alt_counts_refr)
    case_counts = case.get_kmer_counts_many([altseq])[0][keep]
    ctrl_counts = [control.get_kmer_counts_many([altseq])[0][keep] for control in controls]
    return case_counts.tolist(), [c.tolist() for c in ctrl_counts], alt_counts_refr.tolist()


def _drop_outliers(counts):
    """keep abundances within 20 of the list's mean (kevlar/simlike.py:38-48)"""
    counts = np.asarray(counts, dtype=np.int64)
    if len(counts) == 0:
        raise ZeroDivisionError('division by zero')   # what the reference's mean of an empty list raises
    mean = counts.sum() / len(counts)
    return counts[np.abs(counts - mean) < 20]


def discard_outlier_abunds(case_counts, ctrl_counts):
    return _drop_outliers(case_counts).tolist(), [_drop_outliers(c).tolist() for c in ctrl_counts]


def spanning_kmer_abundances_many(windows, case, controls, refr, dropoutliers=False):
    """`spanning_kmer_abundances` for a list of (altseq, refrseq) windows with one query batch
    per sketch.  Returns a list of (abundances, refr_abunds, ndropped) in window order."""
    k = case.ksize()
    alts = [alt for alt, _ in windows]
    alt_in_refr = refr.get_kmer_counts_many(alts)
    case_counts = case.get_kmer_counts_many(alts)
    ctrl_counts = [control.get_kmer_counts_many(alts) for control in controls]
    # SNV/MNV windows (same length): the reference allele's k-mers pair up with the alt allele's
    paired = [i for i, (alt, ref) in enumerate(windows) if len(alt) == len(ref)]
    refr_allele = dict(zip(paired, refr.get_kmer_counts_many([windows[i][1] for i in paired]))) if paired else {}
    results = []
    for i, (altseq, refrseq) in enumerate(windows):
        keep = _valid_positions(alt_in_refr[i])
        kid = case_counts[i][keep].tolist()
        ctrls = [c[i][keep].tolist() for c in ctrl_counts]
        if dropoutliers:
            kid, ctrls = discard_outlier_abunds(kid, ctrls)
        ndropped = (len(altseq) - k + 1) - len(kid)
        if i in refr_allele:
            refr_abunds = refr_allele[i][keep].tolist()
        else:   # indel: no correspondence between alt and refr k-mers
            refr_abunds = [None] * len(kid)
        results.append(([kid] + ctrls, refr_abunds, ndropped))
    return results


def spanning_kmer_abundances(altseq, refrseq, case, controls, refr, dropoutliers=False):
    """kevlar/simlike.py:51-96.  abundances = [case list, control-1 list, ...] over the alt-allele
    k-mers absent from the reference genome; refr_abunds = genomic counts of the paired
    reference-allele k-mers (None per k-mer for indels); ndropped = k-mers left out."""
    return spanning_kmer_abundances_many([(altseq, refrseq)], case, controls, refr, dropoutliers=dropoutliers)[0]
