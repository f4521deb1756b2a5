"""Sketch allocation, loading and FPR estimation -- same interface as kevlar/sketch.py,
on top of the GPU-backed classes in kevlar_b200.khmer."""
import kevlar_b200
from kevlar_b200 import khmer
from kevlar_b200.fastx import batch_from_sequences

# (count, graph, smallcount) -> class; the reference nests three dicts (kevlar/sketch.py:29-51)
_CLASS_BY_TRAIT = {
    (True, True, True): khmer.SmallCountgraph, (True, True, False): khmer.Countgraph,
    (True, False, True): khmer.SmallCounttable, (True, False, False): khmer.Counttable,
    (False, True, True): khmer.Nodegraph, (False, True, False): khmer.Nodegraph,
    (False, False, True): khmer.Nodetable, (False, False, False): khmer.Nodetable,
}

_EXTENSIONS = {
    khmer.Nodetable: ('.nt', '.nodetable'), khmer.Nodegraph: ('.ng', '.nodegraph'),
    khmer.Counttable: ('.ct', '.counttable'), khmer.Countgraph: ('.cg', '.countgraph'),
    khmer.SmallCounttable: ('.sct', '.smallcounttable'), khmer.SmallCountgraph: ('.scg', '.smallcountgraph'),
}

sketch_loader_by_filename_extension = {
    ext: cls.load for cls, exts in _EXTENSIONS.items() for ext in exts
}


class KevlarSketchTypeError(ValueError):
    pass


class KevlarUnsuitableFPRError(SystemExit):
    pass


def estimate_fpr(sketch):
    """Occupancy of table 0 over the SMALLEST table, to the power of the table count
    (kevlar/sketch.py:62-74)."""
    sizes = sketch.hashsizes()
    return (float(sketch.n_occupied()) / min(sizes)) ** float(len(sizes))


def load(filename):
    """Load a sketch; its type comes from the file name extension (kevlar/sketch.py:77-92)."""
    if not filename.endswith(tuple(sketch_loader_by_filename_extension)):
        raise KevlarSketchTypeError('unable to determine sketch type from filename ' + filename)
    return sketch_loader_by_filename_extension['.' + filename.split('.')[-1]](filename)


def get_extension(count=False, graph=False, smallcount=False):
    return _EXTENSIONS[_CLASS_BY_TRAIT[(bool(count), bool(graph), bool(smallcount))]]


def allocate(ksize, target_tablesize, num_tables=4, count=False, graph=False, smallcount=False):
    """New empty sketch in GPU memory (kevlar/sketch.py:99-119)."""
    cls = _CLASS_BY_TRAIT[(bool(count), bool(graph), bool(smallcount))]
    return cls(ksize, target_tablesize, num_tables)


def autoload(infile, count=True, graph=False, ksize=31, table_size=1e4, num_tables=4, num_bands=None, band=None):
    """Load `infile` if its extension names a sketch type, otherwise count its reads into a
    fresh sketch (kevlar/sketch.py:122-153)."""
    try:
        return load(infile)
    except KevlarSketchTypeError:
        sketch = allocate(ksize, table_size, num_tables, count=count, graph=graph, smallcount=False)
        if num_bands:
            assert band >= 0 and band < num_bands
            sketch.consume_seqfile_banding(infile, num_bands, band)
        else:
            sketch.consume_seqfile(infile)
        return sketch


def load_sketchfiles(sketchfiles, maxfpr=0.2):
    """Load pre-computed abundances, refusing sketches whose FPR is too high
    (kevlar/sketch.py:156-170)."""
    sketches = []
    for sketchfile in sketchfiles:
        kevlar_b200.plog('[kevlar::sketch]    ', 'loading sketchfile "{}"...'.format(sketchfile), end='')
        sketch = autoload(sketchfile)
        fpr = estimate_fpr(sketch)
        message = 'done! estimated false positive rate is {:1.3f}'.format(fpr)
        if fpr > maxfpr:
            raise KevlarUnsuitableFPRError(message + ' (FPR too high, bailing out!!!)')
        kevlar_b200.plog(message)
        sketches.append(sketch)
    return sketches


def mask_from_windows(windows, ksize, maskmem, maskfile=None, maxfpr=0.01, logprefix='[kevlar::call]'):
    """Nodetable of the k-mers spanning called variants -- the `--gen-mask` step of `kevlar call`
    and `kevlar alac` (kevlar/call.py:136-172, kevlar/alac.py:49-65), which the reference feeds
    one `mask.consume(window)` at a time: here all windows go to the GPU as one batch.  Windows
    that are None or shorter than k are ignored, like there; the FPR warning text is unchanged."""
    kevlar_b200.plog(logprefix, 'generating mask of variant-spanning k-mers')
    numtables = 4
    buckets = maskmem * khmer._buckets_per_byte['nodegraph'] / numtables
    mask = khmer.Nodetable(ksize, buckets, numtables)
    usable = [w for w in windows if w is not None and len(w) >= ksize]
    bad = [w for w in usable if set(w) - set('ACGT')]
    if bad:   # khmer's consume(str) raises on these
        raise ValueError('invalid DNA character in sequence')
    if usable:
        batch = batch_from_sequences(usable)
        mask.consume_batch(batch.bases, batch.offsets)
    fpr = khmer.calc_expected_collisions(mask, max_false_pos=1.0)
    if fpr > maxfpr:
        kevlar_b200.plog(logprefix, 'WARNING: mask FPR is {:.4f}; exceeds user-specified limit of {:.4f}'.format(fpr, maxfpr))
    if maskfile:
        mask.save(maskfile)
    return mask
