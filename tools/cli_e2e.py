#!/usr/bin/env python
"""Wall-clock of the file-based CLI on config-2-sized FASTQ files (host parsing included)."""
import os
import sys
import tempfile
import time


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def write_fastq(path, bases, offs, tag):
    n = len(offs) - 1
    L = int(offs[1] - offs[0])
    seqs = bases.reshape(n, L)
    with open(path, 'wb') as fh:
        chunk = []
        q = b'I' * L
        for i in range(n):
            chunk.append(b'@%s_%d\n%s\n+\n%s\n' % (tag.encode(), i, seqs[i].tobytes(), q))
            if len(chunk) == 50000:
                fh.write(b''.join(chunk))
                chunk = []
        fh.write(b''.join(chunk))


def main():
    import kevlar_b200 as kv
    from kevlar_b200 import simtrio
    trio = simtrio.simulate_trio(1000000)
    d = tempfile.mkdtemp()
    names = ['proband', 'mother', 'father']
    for (b, o), nm in zip(trio, names):
        write_fastq(os.path.join(d, nm + '.fq'), b, o, nm)
    kv.khmer.Counttable(31, 1e4, 4)   # context + library warm-up
    t0 = time.time()
    per_count = []
    for nm in names:
        t = time.time()
        args = kv.cli.parser().parse_args(['count', '--memory', '64M', os.path.join(d, nm + '.ct'), os.path.join(d, nm + '.fq')])
        kv.count.main(args)
        per_count.append(round(time.time() - t, 3))
    t1 = time.time()
    print('per count call:', per_count)
    args = kv.cli.parser().parse_args(['novel', '--case', os.path.join(d, 'proband.fq'), '--case-counts', os.path.join(d, 'proband.ct'),
                                       '--control-counts', os.path.join(d, 'mother.ct'), os.path.join(d, 'father.ct'),
                                       '-o', os.path.join(d, 'novel.augfastq')])
    kv.novel.main(args)
    t2 = time.time()
    t3 = time.time()
    n = 0
    for batch in kv.khmer.ReadParser(os.path.join(d, 'proband.fq')).batches(64 << 20):
        n += len(batch)
    t4 = time.time()
    print('count x3 (parse + GPU + save): %.2f s; novel (load 3 sketches + parse + GPU + write): %.2f s; '
          'parse only 300k reads: %.2f s (%d reads)' % (t1 - t0, t2 - t1, t4 - t3, n))


if __name__ == '__main__':
    main()
