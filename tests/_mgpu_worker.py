"""Worker for tests/test_multigpu.py: one rank per GPU under torchrun (NCCL).

Each rank counts its shard of the same synthetic reads into a partial sketch, the partial
sketches are merged with every strategy of kevlar_b200.multigpu, and the result must be
byte-identical to the CPU oracle counting all reads in one process.  The novel scan then runs
shard-local and the gathered hits must equal the oracle's."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    import kevlar_b200 as kv
    from kevlar_b200 import multigpu
    from oracle import khmer_oracle as ko
    rank, world = multigpu.init_from_env()
    rng = np.random.default_rng(77)
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)
    genome = letters[rng.integers(0, 4, size=20000)]
    child = genome.copy()
    child[7000] = letters[(np.searchsorted(letters, child[7000]) + 1) % 4]

    def reads(g, n, seed):
        r = np.random.default_rng(seed)
        out = []
        for _ in range(n):
            s = int(r.integers(0, len(g) - 120))
            out.append(g[s:s + int(r.integers(40, 120))].tobytes())
        return out
    samples = [reads(child, 6000, 1) + [b'ACGTTGCAAGGCTTAACCGGTTAAACCCGGGTTTACGT'] * 700, reads(genome, 6000, 2),
               reads(genome, 6000, 3)]
    failures = []
    only = os.environ.get('KV_MGPU_ONLY', '')
    for how in (() if only else ('allreduce', 'allgather', 'p2p', 'p2p_host')):
        for cls in ('Counttable', 'SmallCounttable', 'Nodetable', 'Countgraph'):
            gpu, cpu = [], []
            for seqs in samples:
                bases, offs = ko.reads_to_batch(seqs)
                mb, mo = multigpu.shard_batch(bases, offs, rank, world)
                g = getattr(kv.khmer, cls)(25, 30000, 4)
                g.consume_batch(mb, mo)
                multigpu.merge_sketch(g, how=how)
                c = getattr(ko, cls)(25, 30000, 4)
                c.consume_batch(bases, offs)
                for t in range(4):
                    if g.table_bytes(t) != c.table_bytes(t):
                        failures.append('{} {} table {} differs on rank {}'.format(how, cls, t, rank))
                if g.n_occupied() != c.n_occupied():
                    failures.append('{} {} n_occupied differs'.format(how, cls))
                gpu.append(g)
                cpu.append(c)
            if cls in ('Counttable', 'Countgraph'):
                bases, offs = ko.reads_to_batch(samples[0])
                mb, mo = multigpu.shard_batch(bases, offs, rank, world)
                lo, _ = multigpu.shard_bounds(len(samples[0]), rank, world)
                hits, flags, _ = kv.khmer.novel_batch(gpu[:1], gpu[1:], mb, mo, 6, 1)
                allhits = multigpu.gather_hits(hits, lo)
                ohits, _ = ko.novel_batch(cpu[:1], cpu[1:], bases, offs, 6, 1)
                order = np.lexsort((allhits['offset'], allhits['read']))
                allhits = allhits[order]
                same = len(allhits) == len(ohits) and (allhits['read'] == ohits['read']).all() and \
                    (allhits['offset'] == ohits['offset']).all() and \
                    (allhits['abund'][:, :3] == ohits['abund'][:, :3]).all()
                if not same or len(ohits) == 0:
                    failures.append('{} {} novel hits differ ({} vs {})'.format(how, cls, len(allhits), len(ohits)))
            multigpu.peer_sync_status()
            multigpu.release_p2p(gpu)   # collective: unmap everywhere, barrier, only then free
            del gpu
    # ---- n_unique_kmers across ranks: reads sharded in file order, partial sketches merged, the merged sketch
    # reports the number the single-threaded reference logs (order-dependent)
    for cls in (() if only else ('Counttable', 'SmallCounttable', 'Nodetable', 'Countgraph')):
        for si, seqs in enumerate(samples):
            bases, offs = ko.reads_to_batch(seqs)
            mb, mo = multigpu.shard_batch(bases, offs, rank, world)
            bands = (5, 3) if si == 1 else (None, None)
            g = getattr(kv.khmer, cls)(25, 30000, 4)
            c = getattr(ko, cls)(25, 30000, 4)
            multigpu.count_sharded(g, (mb, mo), num_bands=bands[0], band=bands[1])
            c.consume_batch(bases, offs, num_bands=bands[0] or 0, band=bands[1] or 0)
            if g.n_unique_kmers() != c.n_unique_kmers():
                failures.append('count_sharded {} sample {}: n_unique {} != {} on rank {}'.format(
                    cls, si, g.n_unique_kmers(), c.n_unique_kmers(), rank))
            for t in range(4):
                if g.table_bytes(t) != c.table_bytes(t):
                    failures.append('count_sharded {} table {} differs'.format(cls, t))
            multigpu.peer_sync_status()
            multigpu.release_p2p([g])
            del g
    # ---- the whole trio in one call: each sample's merge runs on the merge lane while the next sample is being
    # counted (kv_merge_fork / kv_merge_join); twice over the same sketches (clear in between), with and without
    # the exact n_unique_kmers, then the shard-local novel scan on the merged sketches
    for cls in (() if only else ('Counttable', 'SmallCounttable')):
        trio_g = [getattr(kv.khmer, cls)(25, 30000, 4) for _ in samples]
        trio_c = [getattr(ko, cls)(25, 30000, 4) for _ in samples]
        shards = []
        for seqs, c in zip(samples, trio_c):
            bases, offs = ko.reads_to_batch(seqs)
            shards.append(multigpu.shard_batch(bases, offs, rank, world))
            c.consume_batch(bases, offs)
        for exact in (True, False, True):
            for g in trio_g:
                g.clear()
            multigpu.count_sharded(trio_g, shards, exact_unique=exact)
            for si, (g, c) in enumerate(zip(trio_g, trio_c)):
                if exact and g.n_unique_kmers() != c.n_unique_kmers():
                    failures.append('pipelined count_sharded {} sample {}: n_unique {} != {} on rank {}'.format(
                        cls, si, g.n_unique_kmers(), c.n_unique_kmers(), rank))
                for t in range(4):
                    if g.table_bytes(t) != c.table_bytes(t):
                        failures.append('pipelined count_sharded {} sample {} table {} differs on rank {} (exact={})'.format(
                            cls, si, t, rank, exact))
        bases, offs = ko.reads_to_batch(samples[0])
        lo, _ = multigpu.shard_bounds(len(samples[0]), rank, world)
        hits, flags, _ = kv.khmer.novel_batch(trio_g[:1], trio_g[1:], shards[0][0], shards[0][1], 6, 1)
        allhits = multigpu.gather_hits(hits, lo)
        ohits, _ = ko.novel_batch(trio_c[:1], trio_c[1:], bases, offs, 6, 1)
        allhits = allhits[np.lexsort((allhits['offset'], allhits['read']))]
        if len(allhits) != len(ohits) or not (allhits['offset'] == ohits['offset']).all() or \
                not (allhits['abund'][:, :3] == ohits['abund'][:, :3]).all():
            failures.append('pipelined count_sharded {}: novel hits differ ({} vs {})'.format(cls, len(allhits), len(ohits)))
        multigpu.peer_sync_status()
        multigpu.release_p2p(trio_g)
        del trio_g
    import tempfile
    shared = os.environ.get('KV_TEST_SHARED_DIR') or tempfile.gettempdir()
    # ---- plan B, second design: spanning sketches (tables spread over the HBM of all ranks, updates exchanged
    # inside the tiled apply kernel, queries as plain loads over NVLink).  Tables of 20 M buckets so that several
    # ranks hold pieces; banded and masked counts; the saved file, n_occupied and the shard-local novel scan
    # must equal the single-process oracle.
    for cls in ('Counttable', 'SmallCounttable', 'Nodetable', 'Countgraph'):
        span, cpu = [], []
        for si, seqs in enumerate(samples):
            bases, offs = ko.reads_to_batch(seqs)
            mb, mo = multigpu.shard_batch(bases, offs, rank, world)
            sk = multigpu.SpanningSketch(getattr(kv.khmer, cls), 25, 20000003, 4, chunk_positions=131072)
            c = getattr(ko, cls)(25, 20000003, 4)
            bands = (3, 1) if si == 2 else (None, None)
            n_span = sk.consume_batch(mb, mo, num_bands=bands[0], band=bands[1])
            n_cpu = c.consume_batch(bases, offs, num_bands=bands[0] or 0, band=bands[1] or 0)
            if n_span != n_cpu:
                failures.append('spanning {} k-mer count {} != {}'.format(cls, n_span, n_cpu))
            if sk.n_occupied() != c.n_occupied():
                failures.append('spanning {} n_occupied {} != {}'.format(cls, sk.n_occupied(), c.n_occupied()))
            for t in range(4):   # every rank reads the whole table through its own mapping
                if sk.sketch.table_bytes(t) != c.table_bytes(t):
                    failures.append('spanning {} sample {} table {} differs on rank {}'.format(cls, si, t, rank))
            path = os.path.join(shared, 'kv_span_{}_{}.sketch'.format(cls, si))
            sk.save(path)
            if rank == 0:
                ref = path + '.oracle'
                c.save(ref)
                if open(path, 'rb').read() != open(ref, 'rb').read():
                    failures.append('spanning {} sample {}: saved file differs from the oracle file'.format(cls, si))
                os.remove(ref)
                os.remove(path)
            span.append(sk)
            cpu.append(c)
        if cls in ('Counttable', 'Countgraph'):
            bases, offs = ko.reads_to_batch(samples[0])
            mb, mo = multigpu.shard_batch(bases, offs, rank, world)
            lo, _ = multigpu.shard_bounds(len(samples[0]), rank, world)
            hits, flags, _ = kv.khmer.novel_batch([span[0].sketch], [s.sketch for s in span[1:]], mb, mo, 6, 1)
            allhits = multigpu.gather_hits(hits, lo)
            ohits, _ = ko.novel_batch(cpu[:1], cpu[1:], bases, offs, 6, 1)
            allhits = allhits[np.lexsort((allhits['offset'], allhits['read']))]
            same = len(allhits) == len(ohits) and (allhits['read'] == ohits['read']).all() and \
                (allhits['offset'] == ohits['offset']).all() and (allhits['abund'][:, :3] == ohits['abund'][:, :3]).all()
            if not same or len(ohits) == 0:
                failures.append('spanning {} novel hits differ ({} vs {})'.format(cls, len(allhits), len(ohits)))
            # a second batch on top (clear, then two halves == one batch)
            span[0].clear()
            half = len(mo) // 2
            span[0].consume_batch(mb[:int(mo[half])], mo[:half + 1])
            span[0].consume_batch(mb[int(mo[half]):], mo[half:] - mo[half])
            for t in range(4):
                if span[0].sketch.table_bytes(t) != cpu[0].table_bytes(t):
                    failures.append('spanning {} two-batch table {} differs'.format(cls, t))
        for sk in span:
            sk.close()
        del span
    if only == 'span':
        multigpu.release_peer_sync()
        torch.distributed.barrier()
        if failures:
            print('RANK', rank, 'FAILURES:', failures)
            sys.exit(1)
        if rank == 0:
            print('spanning sketches OK on', world, 'ranks')
        torch.distributed.destroy_process_group()
        return
    multigpu.release_peer_sync()
    # ---- plan B: bin-range-sharded sketches: count shard-local reads, exchange hashes, save one file
    for cls in ('Counttable', 'SmallCounttable', 'Nodetable', 'Countgraph'):
        sharded, cpu = [], []
        for si, seqs in enumerate(samples):
            bases, offs = ko.reads_to_batch(seqs)
            mb, mo = multigpu.shard_batch(bases, offs, rank, world)
            sk = multigpu.ShardedSketch(getattr(kv.khmer, cls), 25, 30011, 4)
            c = getattr(ko, cls)(25, 30011, 4)
            n_sharded = sk.consume_batch(mb, mo)
            n_cpu = c.consume_batch(bases, offs)
            if n_sharded != n_cpu:
                failures.append('sharded {} k-mer count {} != {}'.format(cls, n_sharded, n_cpu))
            if sk.n_occupied() != c.n_occupied():
                failures.append('sharded {} n_occupied {} != {}'.format(cls, sk.n_occupied(), c.n_occupied()))
            path = os.path.join(shared, 'kv_sharded_{}_{}.sketch'.format(cls, si))
            sk.save(path)
            if rank == 0:
                ref = path + '.oracle'
                c.save(ref)
                if open(path, 'rb').read() != open(ref, 'rb').read():
                    failures.append('sharded {} sample {}: saved file differs from the oracle file'.format(cls, si))
                os.remove(ref)
            torch.distributed.barrier()
            if rank == 0:
                os.remove(path)
            sharded.append(sk)
            cpu.append(c)
        if cls in ('Counttable', 'Countgraph'):
            bases, offs = ko.reads_to_batch(samples[0])
            mb, mo = multigpu.shard_batch(bases, offs, rank, world)
            lo, _ = multigpu.shard_bounds(len(samples[0]), rank, world)
            hits, flags, _ = multigpu.novel_batch_sharded(sharded[:1], sharded[1:], mb, mo, 6, 1)
            allhits = multigpu.gather_hits(hits, lo)
            ohits, _ = ko.novel_batch(cpu[:1], cpu[1:], bases, offs, 6, 1)
            allhits = allhits[np.lexsort((allhits['offset'], allhits['read']))]
            same = len(allhits) == len(ohits) and (allhits['read'] == ohits['read']).all() and \
                (allhits['offset'] == ohits['offset']).all() and (allhits['abund'][:, :3] == ohits['abund'][:, :3]).all()
            if not same or len(ohits) == 0:
                failures.append('sharded {} novel hits differ ({} vs {})'.format(cls, len(allhits), len(ohits)))
        del sharded
    torch.distributed.barrier()
    if failures:
        print('RANK', rank, 'FAILURES:', failures)
        sys.exit(1)
    if rank == 0:
        print('multi-GPU merge OK on', world, 'ranks: allreduce/allgather/p2p/p2p_host x 4 sketch types, novel hits identical; '
              'spanning and bin-range-sharded count / save / novel identical to the oracle')
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
