#!/usr/bin/env python
"""Kernel throughput on HBM-resident sketches (BASELINE config 3 shape, scaled to one GPU-minute).

Not the headline benchmark: a probe of the large-sketch regime where every counter update is a
random DRAM sector read-modify-write.  10 Mbp genome at 30x -> 3 M reads x 100 bp per sample,
k=31, 8-bit Counttable with 4 tables of `--memory` bytes in total (default 4 GB)."""
import argparse
import json
import os
import sys


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--memory', type=float, default=4e9)
    ap.add_argument('--genome', type=int, default=10000000)
    ap.add_argument('--reads', type=int, default=3000000)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--track', action='store_true')
    ap.add_argument('--bits', type=int, default=8, choices=[8, 4, 1])
    ap.add_argument('--label', default='')
    args = ap.parse_args()
    import torch
    from kevlar_b200 import _lib, khmer, simtrio
    dev = torch.device('cuda', 0)
    dtrio = simtrio.device_trio(args.genome, args.reads)      # reads drawn on the device (kv_synth_reads)
    cls = {8: khmer.Counttable, 4: khmer.SmallCounttable, 1: khmer.Nodetable}[args.bits]
    sketches = [cls(31, args.memory / 4 * (8 // args.bits), 4) for _ in range(3)]
    for sk in sketches:
        sk.set_unique_tracking(args.track)
    stream = torch.cuda.ExternalStream(_lib.stream_ptr(0), device=dev)
    kmers = args.reads * 70

    def count_all():
        for sk, (b, o) in zip(sketches, dtrio):
            sk.clear()
            sk.consume_batch(b.data_ptr(), (o.data_ptr(), o.numel() - 1), where=khmer.MEM_DEVICE, wait=False)

    def novel():
        b, o = dtrio[0]
        return khmer.novel_batch(sketches[:1], sketches[1:], b.data_ptr(), (o.data_ptr(), o.numel() - 1, b.numel()), 6, 1,
                                 where=khmer.MEM_DEVICE)[0]
    count_all()
    hits = novel()
    _lib.sync(0)
    res = {}
    for name, fn, nk in (('count_x3', count_all, 3 * kmers), ('novel', novel, kmers)):
        _lib.profile(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            fn()
        e1.record(stream)
        _lib.sync(0)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        prof = {k: round(v[0] / args.steps, 3) for k, v in _lib.profile(0).items() if v[1]}
        res[name] = {'ms': round(ms, 3), 'G_kmers_per_s': round(nk / ms / 1e6, 3), 'kernel_ms': prof}
    res['config'] = {'sketch_bytes': args.memory, 'reads_per_sample': args.reads, 'tracking': args.track, 'bits': args.bits,
                     'novel_hits': int(len(hits)), 'label': args.label,
                     'env': {k: v for k, v in os.environ.items() if k.startswith('KV_')}}
    inc_ms = (res['count_x3']['kernel_ms'].get('increment', 0) + res['count_x3']['kernel_ms'].get('partition', 0)) / 3
    if inc_ms:
        res['increment'] = {'ms_per_sample': round(inc_ms, 3), 'G_updates_per_s': round(4 * kmers / inc_ms / 1e6, 2),
                            'sector_GBps_read_plus_write': round(4 * kmers * 64 / inc_ms / 1e6, 1)}
    print(json.dumps(res))


if __name__ == '__main__':
    main()
