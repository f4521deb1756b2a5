"""Worker for tests/test_host.py::test_gloo_* -- one of WORLD_SIZE CPU processes (gloo).

Exercises the multi-GPU host logic without a GPU: read sharding + the widen -> all-reduce ->
clamp merge choreography of kevlar_b200.multigpu, with oracle sketches standing in for the GPU
ones (the adapter below is the numpy twin of GpuSketchAdapter)."""
import ctypes
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


class OracleAdapter(object):
    def __init__(self, sketch, ko):
        import torch
        self.torch, self.sketch, self.ko = torch, sketch, ko
        self.bits = sketch._bits

    def _tables(self):
        return [np.frombuffer(self.sketch.table_bytes(t), dtype=np.uint8) for t in range(len(self.sketch.hashsizes()))]

    def widen(self):
        flat = np.concatenate(self._tables())
        if self.bits == 4:
            flat = np.stack([flat >> 4, flat & 15], axis=1).reshape(-1)
        return self.torch.from_numpy(flat.astype(np.int32))

    def _store(self, flat):
        pos = 0
        for t in range(len(self.sketch.hashsizes())):
            n = self.ko._lib.ko_table_nbytes(self.sketch._h, t)
            ctypes.memmove(self.ko._lib.ko_table_ptr(self.sketch._h, t), flat[pos:pos + n].tobytes(), n)
            pos += n

    def narrow(self, wide):
        vals = wide.numpy()
        if self.bits == 8:
            flat = np.minimum(vals, 255).astype(np.uint8)
        else:
            v = np.minimum(vals, 15).astype(np.uint8).reshape(-1, 2)
            flat = (v[:, 0] << 4) | v[:, 1]
        self._store(flat)

    def flat_tensor(self):
        return self.torch.from_numpy(np.concatenate(self._tables()).copy())

    def merge_from(self, tensors):
        acc = np.concatenate(self._tables())
        for t in tensors:
            acc = acc | t.numpy()
        self._store(acc)


def main():
    rank, world, port, outdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=port)
    from kevlar_b200 import multigpu
    from oracle import khmer_oracle as ko
    r, w = multigpu.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    rng = np.random.default_rng(1234)              # same reads on every rank, then sharded
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)
    genome = letters[rng.integers(0, 4, size=3000)]
    seqs = []
    for _ in range(4000):
        s = int(rng.integers(0, 2900))
        seqs.append(genome[s:s + int(rng.integers(25, 100))].tobytes())
    bases, offs = ko.reads_to_batch(seqs)
    mine_b, mine_o = multigpu.shard_batch(bases, offs, rank, world)
    results = {}
    for name in ('Counttable', 'SmallCounttable', 'Nodetable'):
        sk = getattr(ko, name)(21, 900, 4)          # tiny tables: saturation on every rank
        sk.consume_batch(mine_b, mine_o)
        multigpu.merge_allreduce(OracleAdapter(sk, ko), None)
        results[name] = [sk.table_bytes(t) for t in range(4)]
    hits = np.zeros(3, dtype=[('read', '<u4'), ('offset', '<u4'), ('abund', 'u1', (16,))])
    hits['read'] = np.arange(3)
    lo, _ = multigpu.shard_bounds(len(seqs), rank, world)
    allhits = multigpu.gather_hits(hits, lo)
    results['hit_reads'] = allhits['read'].tolist()
    np.save(os.path.join(outdir, 'rank{}.npy'.format(rank)), np.array([results], dtype=object), allow_pickle=True)
    import torch.distributed as td
    td.barrier()
    td.destroy_process_group()


if __name__ == '__main__':
    main()
