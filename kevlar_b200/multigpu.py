"""Multi-GPU count + novel: one process per GPU (torchrun), reads sharded across ranks.

SURVEY.md 8(e), plan A.  Reads are independent units and saturating increments commute, so

    final[bucket] = min(MAX, sum over ranks of partial[bucket])

holds exactly (each partial is itself clamped at MAX, which does not change the clamped sum).
Every rank counts its shard of reads into a full-size partial sketch, then the partial
sketches are merged -- the one real exchange step of the path -- and every rank holds the
complete sketch.  The novel scan then runs shard-local against the replicated sketches with
no further communication; ranks only exchange their (tiny) hit lists at the end.

Merge strategies (`merge_sketch(..., how=)`):
  'allreduce'  widen counters on the GPU (kv_sketch_widen: u8->fp16 (exact to 2048), nibble->u8),
               NCCL all-reduce (SUM; MAX-free OR for bit tables via all-gather), clamp + repack
               (kv_sketch_narrow).  Transport by NCCL over NVLink/NVSwitch, arithmetic by our kernels.
  'allgather'  NCCL all-gather of the raw tables, then ONE saturating-merge kernel
               (kv_sketch_merge_peers, per-byte __vaddus4) over the gathered copies.
  'p2p'        no NCCL on the data path: ranks exchange CUDA IPC handles once and the merge kernel
               reads the peers' tables directly over NVLink (peer-mapped loads).

torch / torch.distributed are plumbing here (rendezvous, NCCL communicator, device buffers).
"""
import ctypes
import os
from ctypes import byref, c_int, c_uint64, c_void_p

import numpy as np

from kevlar_b200 import _lib
from kevlar_b200._lib import check, lib


def dist():
    import torch.distributed as td
    return td


def init_from_env(backend=None):
    """Join the torchrun rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  Returns
    (rank, world).  Single-process runs need no rendezvous."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1:
        import torch
        td = dist()
        if not td.is_initialized():
            if backend is None:
                backend = 'nccl' if torch.cuda.is_available() else 'gloo'
            if backend == 'nccl':
                torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
            td.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def shard_bounds(n_items, rank, world):
    """Contiguous, near-equal split of n_items over `world` ranks: [lo, hi) for `rank`."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(bases, offsets, rank, world):
    """The slice of a (bases, offsets) batch that `rank` owns: reads [lo, hi) with offsets
    rebased to start at 0.  Every read lands on exactly one rank."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    lo, hi = shard_bounds(len(offsets) - 1, rank, world)
    b0, b1 = int(offsets[lo]), int(offsets[hi])
    return np.asarray(bases)[b0:b1], np.ascontiguousarray(offsets[lo:hi + 1] - offsets[lo])


class GpuSketchAdapter(object):
    """What merge_sketch needs from a sketch, implemented with the CUDA library."""

    def __init__(self, sketch):
        import torch
        self.torch = torch
        self.sketch = sketch
        self.bits = sketch._bits
        self.device = torch.device('cuda', sketch.device)

    def widen(self):
        n, eb = c_uint64(), c_int()
        check(lib().kv_sketch_widen(self.sketch._h, None, byref(n), byref(eb)))
        dtype = self.torch.float16 if eb.value == 2 else self.torch.uint8
        buf = self.torch.empty(n.value, dtype=dtype, device=self.device)
        check(lib().kv_sketch_widen(self.sketch._h, buf.data_ptr(), byref(n), byref(eb)))
        return buf

    def narrow(self, buf):
        self.torch.cuda.synchronize(self.device)
        check(lib().kv_sketch_narrow(self.sketch._h, buf.data_ptr()))

    def flat_tensor(self):
        """The sketch's own table storage viewed as a torch uint8 tensor (no copy)."""
        ptr, nbytes = self.sketch.flat_device_buffer()

        class _Raw(object):
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 3}
        raw._keepalive = self.sketch
        return self.torch.as_tensor(raw, device=self.device)

    def merge_from(self, tensors):
        self.torch.cuda.synchronize(self.device)
        for i in range(0, len(tensors), 8):
            group = tensors[i:i + 8]
            ptrs = (c_void_p * len(group))(*[t.data_ptr() for t in group])
            check(lib().kv_sketch_merge_peers(self.sketch._h, ptrs, len(group), 0, 0))
        _lib.sync(self.sketch.device)


def merge_allreduce(adapter, group=None):
    """widen -> all-reduce(SUM) -> clamp.  Bit tables (merge = OR) go through all-gather since
    NCCL has no bitwise reduction."""
    td = dist()
    world = td.get_world_size(group)
    if adapter.bits == 1 or (adapter.bits == 8 and world > 8) or (adapter.bits == 4 and world > 17):
        # OR has no NCCL op; fp16 sums of 8-bit counters are exact up to 8 ranks (8 x 255 < 2048); nibbles
        # travel as u8 and 15 x 18 would wrap
        return merge_allgather(adapter, group)
    wide = adapter.widen()
    td.all_reduce(wide, op=td.ReduceOp.SUM, group=group)
    adapter.narrow(wide)


def merge_allgather(adapter, group=None):
    td = dist()
    mine = adapter.flat_tensor()
    world, rank = td.get_world_size(group), td.get_rank(group)
    gathered = [mine.new_empty(mine.shape) for _ in range(world)]
    td.all_gather(gathered, mine, group=group)
    adapter.merge_from([g for r, g in enumerate(gathered) if r != rank])


def slice_bounds(nbytes, rank, world, align=256):
    """Byte range of the table storage that `rank` reduces in the p2p merge (aligned)."""
    units = nbytes // align
    lo, hi = shard_bounds(units, rank, world)
    return lo * align, hi * align


def _p2p_peers(sketch, group=None):
    """Map the peers' table storage once per sketch (CUDA IPC open is slow: ~ms) and reuse it.  The
    mappings hang on the sketch OBJECT (not on its handle address, which a later sketch may
    reuse); `release_p2p` closes them."""
    peers = getattr(sketch, '_p2p_peers', None)
    if peers is not None:
        return peers
    td = dist()
    world, rank = td.get_world_size(group), td.get_rank(group)
    handle = (ctypes.c_uint8 * 64)()
    check(lib().kv_sketch_ipc_export(sketch._h, handle))
    handles = [None] * world
    td.all_gather_object(handles, bytes(handle), group=group)
    peers = {}
    for r, h in enumerate(handles):
        if r != rank:
            ptr = c_void_p()
            check(lib().kv_ipc_open(sketch.device, (ctypes.c_uint8 * 64)(*h), byref(ptr)))
            peers[r] = ptr
    sketch._p2p_peers = peers
    return peers


def close_p2p(sketch):
    """Unmap the peers' storage of ONE sketch on THIS rank (local; `_Sketch.__del__` calls it).  It does
    not make freeing the sketch safe by itself: the peers may still have this rank's tables mapped
    -- that needs the collective `release_p2p`."""
    peers = getattr(sketch, '_p2p_peers', None)
    if peers:
        _lib.sync(sketch.device)
        for ptr in peers.values():
            lib().kv_ipc_close(sketch.device, ptr)
    if peers is not None:
        sketch._p2p_peers = None


def release_p2p(sketches, group=None):
    """COLLECTIVE: every rank unmaps its peers' table storage of `sketches`, then all ranks meet at a
    barrier.  Only after it returns may any rank free these sketches (CUDA leaves freeing memory that
    a peer still has mapped undefined).  Call it on every rank, with the same sketches in the same
    order, before the sketches go out of scope or the process group is destroyed."""
    if not isinstance(sketches, (list, tuple)):
        sketches = [sketches]
    for sk in sketches:
        close_p2p(sk)
    td = dist()
    if td.is_initialized() and td.get_world_size(group) > 1:
        td.barrier(group=group)


_PEER_SYNC = {}   # (device, group id, lane) -> kv_peer_sync handle with every peer connected


def _peer_sync(device, group=None, lane=0):
    """This rank's device-side barrier object (created and connected once per device, group and lane; lane 1 =
    the merge lane of kv_merge_fork, whose barriers must not interleave with those of the compute stream)."""
    key = (device, id(group), lane)
    if key in _PEER_SYNC:
        return _PEER_SYNC[key]
    td = dist()
    world, rank = td.get_world_size(group), td.get_rank(group)
    ps, handle = c_void_p(), (ctypes.c_uint8 * 64)()
    check(lib().kv_peer_sync_create(device, rank, world, byref(ps), handle))
    handles = [None] * world
    td.all_gather_object(handles, bytes(handle), group=group)
    for r, h in enumerate(handles):
        if r != rank:
            check(lib().kv_peer_sync_connect(ps, r, (ctypes.c_uint8 * 64)(*h)))
    if lane:
        check(lib().kv_peer_sync_set_lane(ps, lane))
    _PEER_SYNC[key] = ps
    return ps


def peer_sync_status(device=None):
    """Drain the compute stream and raise if a device-side barrier gave up waiting for a peer."""
    for (dev, _, _), ps in _PEER_SYNC.items():
        if device is None or dev == device:
            check(lib().kv_peer_sync_status(ps))


def release_peer_sync():
    """Unmap the peers' barrier flags (COLLECTIVE: call on every rank, after the last merge)."""
    td = dist()
    for ps in _PEER_SYNC.values():
        check(lib().kv_peer_sync_status(ps))
    if _PEER_SYNC and td.is_initialized():
        td.barrier()                                   # nobody still spins on flags about to be unmapped
    for ps in _PEER_SYNC.values():
        check(lib().kv_peer_sync_destroy(ps))
    _PEER_SYNC.clear()


def shutdown(sketches=(), destroy_group=True):
    """COLLECTIVE teardown of everything the peer-to-peer merge set up: peer mappings of `sketches`,
    the device-side barrier objects, and (optionally) the process group.  Every rank must call it --
    typically from a `finally:` -- before its sketches are freed."""
    td = dist()
    if not td.is_initialized():
        return
    try:
        release_p2p(list(sketches))
        release_peer_sync()
    finally:
        if destroy_group:
            td.destroy_process_group()


def merge_p2p(sketches, group=None, host_barriers=False, lane=0):
    """Peer-to-peer merge with no NCCL on the data path.  Ranks exchange CUDA IPC handles of
    their table storage once; then, for all sketches together, rank r owns byte slice(r) of every
    table: it loads that slice from every peer over NVLink, folds it into its own table and

      default:  stores the finished slice straight into every peer's table from the same kernel
                (kv_sketch_allreduce_peers) -- nobody else reads or writes slice(r) anywhere, so
                the whole all-reduce is ONE kernel per sketch between two barriers, and the
                barriers are device-side (kv_peer_barrier: flags in peer-mapped HBM, enqueued on
                the compute stream), so the merge never blocks the host;
      host_barriers=True (how='p2p_host'):  two phases, reduce-scatter (kv_sketch_merge_peers)
                then all-gather (kv_sketch_copy_from_peer pulls slice(p) from peer p), separated
                by stream syncs + process-group barriers -- for ranks that may reach the merge
                more than KV_PEER_TIMEOUT_MS apart, or more than 16 ranks.

    `lane=1` (after kv_merge_fork, device-side barriers only): barriers and merge kernels run on the library's
    merge lane instead of its compute stream."""
    td = dist()
    if not isinstance(sketches, (list, tuple)):
        sketches = [sketches]
    world, rank = td.get_world_size(group), td.get_rank(group)
    if world > 16:
        host_barriers = True
    peers = [_p2p_peers(sk, group) for sk in sketches]
    device = sketches[0].device

    def slices():
        for sk, pr in zip(sketches, peers):
            _, nbytes = sk.flat_device_buffer()
            yield sk, pr, nbytes

    if not host_barriers:
        sync = _peer_sync(device, group, lane)
        check(lib().kv_peer_barrier(sync))            # every partial table is complete
        for sk, pr, nbytes in slices():
            lo, hi = slice_bounds(nbytes, rank, world)
            ptrs = (c_void_p * len(pr))(*[pr[r].value for r in sorted(pr)])
            check(lib().kv_sketch_allreduce_peers(sk._h, ptrs, len(pr), lo, hi))
        check(lib().kv_peer_barrier(sync))            # every slice is reduced and stored everywhere
        return

    def barrier():
        _lib.sync(device)
        td.barrier(group=group)

    barrier()                                         # every partial table is complete
    for sk, pr, nbytes in slices():
        lo, hi = slice_bounds(nbytes, rank, world)
        order = [pr[r] for r in sorted(pr)]
        for i in range(0, len(order), 8):
            grp = order[i:i + 8]
            ptrs = (c_void_p * len(grp))(*[p.value for p in grp])
            check(lib().kv_sketch_merge_peers(sk._h, ptrs, len(grp), lo, hi))
    barrier()                                         # every slice is reduced
    for sk, pr, nbytes in slices():
        for r in sorted(pr):
            plo, phi = slice_bounds(nbytes, r, world)
            check(lib().kv_sketch_copy_from_peer(sk._h, pr[r], plo, phi))
    barrier()                                         # nobody still reads my table


def merge_sketches(sketches, how='p2p', group=None):
    """Combine per-rank partial sketches in place; afterwards every rank holds the full sketches."""
    td = dist()
    if not td.is_initialized() or td.get_world_size(group) == 1:
        return sketches
    if how in ('p2p', 'p2p_host'):
        merge_p2p(sketches, group, host_barriers=(how == 'p2p_host'))
    elif how in ('allgather', 'allreduce'):
        for sketch in sketches:
            (merge_allgather if how == 'allgather' else merge_allreduce)(GpuSketchAdapter(sketch), group)
    else:
        raise ValueError('unknown merge strategy ' + how)
    return sketches


def merge_sketch(sketch, how='p2p', group=None):
    merge_sketches([sketch], how=how, group=group)
    return sketch


def _device_words(ptr, n_words, device, keepalive):
    """A torch int32 view of `n_words` device words at `ptr` (no copy)."""
    import torch

    class _Raw(object):
        pass
    raw = _Raw()
    raw.__cuda_array_interface__ = {'shape': (int(n_words),), 'typestr': '<i4', 'data': (int(ptr), False), 'version': 3}
    raw._keepalive = keepalive
    return torch.as_tensor(raw, device=torch.device('cuda', device))


def _occupancy_view(sketch):
    """(int32 tensor over the sketch's occupancy bitmaps of ALL tables as one range -- no copy --, word offset of
    every table inside it).  The per-table bitmaps sit back to back (with zero padding) in one allocation."""
    nt = len(sketch.hashsizes())
    ptrs, nwords = (c_void_p * nt)(), (c_uint64 * nt)()
    check(lib().kv_sketch_occupancy(sketch._h, ptrs, nwords))
    base = int(ptrs[0])
    starts = [(int(ptrs[t]) - base) // 4 for t in range(nt)]
    assert all(st >= 0 for st in starts) and (int(ptrs[nt - 1]) - base) % 4 == 0
    total = starts[-1] + int(nwords[nt - 1])
    return _device_words(base, total, sketch.device, sketch), starts


def _unique_share(sketch, batch, total_ptr, keep, kw, group, world, rank, from_scratch):
    """Enqueue THIS rank's share of one sketch's n_unique_kmers (a device int64 at `total_ptr`): all-gather of the
    occupancy bitmaps, OR of the lower ranks', first-touch passes over this rank's reads.  `from_scratch`: try the
    hashes the consume call left on the device first (kv_unique_last_batch)."""
    import torch
    td = dist()
    bases, offsets = batch
    mine, starts = _occupancy_view(sketch)
    lower = torch.zeros_like(mine)
    if world > 1:
        gathered = torch.empty((world, mine.numel()), dtype=mine.dtype, device=mine.device)
        td.all_gather_into_tensor(gathered, mine, group=group)
        for q in range(rank):
            lower |= gathered[q]
        keep.append(gathered)
    keep.append(lower)
    occ = (c_void_p * len(starts))(*[lower.data_ptr() + 4 * st for st in starts])
    if from_scratch:
        rc = lib().kv_unique_last_batch(sketch._h, occ, None, total_ptr)
        if rc != _lib.KV_ESTATE:
            check(rc)
            return
    where = kw['where']
    if where == _lib.MEM_HOST:
        bases, offsets = _lib.as_u8(bases), _lib.as_u64(offsets)
        bptr, optr, n_reads = bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1
    else:
        bptr, (optr, n_reads) = bases, offsets
    mask = kw['mask']
    check(lib().kv_unique_batch(sketch._h, occ, bptr, optr, n_reads, where, int(kw['num_bands'] or 0), int(kw['band'] or 0),
                                mask._h if mask is not None else None, int(kw['threshold']), int(bool(kw['consume_masked'])),
                                None, total_ptr))


def _retire(tensors, stream):
    """Scratch tensors must outlive the kernels that read them: hand them to the stream."""
    for tensor in tensors:
        tensor.record_stream(stream)


def unique_across_ranks(sketches, batches, where=_lib.MEM_HOST, num_bands=None, band=None, mask=None, threshold=0,
                        consume_masked=False, group=None, store=True):
    """khmer's n_unique_kmers for samples whose reads are sharded contiguously over the ranks (rank order = file
    order).  COLLECTIVE; call it after every rank has counted its shards into its zeroed partial `sketches`
    and BEFORE the merge.  Each rank ORs the occupancy bitmaps of all lower ranks' partial sketches and re-runs
    the first-touch passes over its own reads with that as the occupied set (kv_unique_batch); the sum over the
    ranks is the number the reference logs (kevlar/count.py:84).  `batches` = one (bases, offsets) per sketch, in
    the form `consume_batch` took for `where` -- for each sketch the batch counted into it LAST.

    Everything is enqueued on the library's stream -- the torch ops and the NCCL collectives run under it too --
    so the host never waits.  Returns a device tensor (int64, one entry per sketch) holding the totals; with
    `store` they are also copied into the sketches (valid once the stream has drained; a merge enqueued later on
    does not disturb them)."""
    import torch
    td = dist()
    world = td.get_world_size(group) if td.is_initialized() else 1
    rank = td.get_rank(group) if td.is_initialized() else 0
    device = sketches[0].device
    dev = torch.device('cuda', device)
    stream = torch.cuda.ExternalStream(_lib.stream_ptr(device), device=dev)
    kw = dict(num_bands=num_bands, band=band, mask=mask, threshold=threshold, consume_masked=consume_masked, where=where)
    keep = []
    with torch.cuda.stream(stream):
        totals = torch.zeros(len(sketches), dtype=torch.int64, device=dev)
        for i, (sketch, batch) in enumerate(zip(sketches, batches)):
            _unique_share(sketch, batch, totals.data_ptr() + 8 * i, keep, kw, group, world, rank, from_scratch=True)
        if world > 1:
            td.all_reduce(totals, op=td.ReduceOp.SUM, group=group)
        if store:
            for i, sketch in enumerate(sketches):
                check(lib().kv_sketch_set_unique_dev(sketch._h, totals.data_ptr() + 8 * i))
        _retire([totals] + keep, stream)
    return totals


def count_sharded(sketches, batches, how='p2p', where=_lib.MEM_HOST, num_bands=None, band=None, mask=None, threshold=0,
                  consume_masked=False, exact_unique=True, group=None, overlap=True):
    """`kevlar count` of one or more samples on all ranks (plan A): THIS rank's contiguous shard of each sample's
    reads goes into the matching sketch (which must be empty), the partial sketches are merged, and -- with
    `exact_unique` -- every merged sketch reports the reference's n_unique_kmers.  COLLECTIVE.  `sketches` /
    `batches` may be single objects or equally long lists.

    With `exact_unique` each sample's share of n_unique_kmers is taken right after its shard has been counted,
    while the shard's hashes are still on the device (kv_unique_last_batch: the reads are copied and hashed
    once); a shard too long for one chunk goes through kv_unique_batch instead.

    `overlap` (several samples, how='p2p'): the merge of sample i runs on the merge lane (kv_merge_fork) while
    sample i+1 is being counted -- each sample's sketch is independent, so only the last merge is exposed."""
    import torch
    single = not isinstance(sketches, (list, tuple))
    if single:
        sketches, batches = [sketches], [batches]
    td = dist()
    world = td.get_world_size(group) if td.is_initialized() else 1
    rank = td.get_rank(group) if td.is_initialized() else 0
    kw = dict(num_bands=num_bands, band=band, mask=mask, threshold=threshold, consume_masked=consume_masked, where=where)
    if world == 1:
        for sketch, (bases, offsets) in zip(sketches, batches):
            sketch.consume_batch(bases, offsets, wait=False, **kw)
        return
    device = sketches[0].device
    dev = torch.device('cuda', device)
    stream = torch.cuda.ExternalStream(_lib.stream_ptr(device), device=dev)
    totals, keep = None, []
    pipelined = overlap and how == 'p2p' and world <= 16 and len(sketches) > 1
    with torch.cuda.stream(stream):
        if exact_unique:
            totals = torch.zeros(len(sketches), dtype=torch.int64, device=dev)
        for i, (sketch, (bases, offsets)) in enumerate(zip(sketches, batches)):
            sketch.set_unique_tracking('deferred' if exact_unique else False)
            sketch.consume_batch(bases, offsets, wait=False, **kw)
            if exact_unique:
                _unique_share(sketch, (bases, offsets), totals.data_ptr() + 8 * i, keep, kw, group, world, rank, from_scratch=True)
            if pipelined:
                check(lib().kv_merge_fork(device))
                merge_p2p([sketch], group, lane=1)
        if exact_unique:
            td.all_reduce(totals, op=td.ReduceOp.SUM, group=group)
    if pipelined:
        check(lib().kv_merge_join(device))
    else:
        if how != 'p2p':
            _lib.sync(device)
        merge_sketches(sketches, how=how, group=group)
    if totals is not None:   # after the merge, which marks the sketches' own counter as stale
        if how != 'p2p':
            _lib.sync(device)
        for i, sketch in enumerate(sketches):
            check(lib().kv_sketch_set_unique_dev(sketch._h, totals.data_ptr() + 8 * i))
        _retire([totals] + keep, stream)


def gather_hits(hits, read_base, group=None):
    """Collect every rank's novel hits on rank 0 with batch-global read indices."""
    td = dist()
    hits = hits.copy()
    local = np.zeros(len(hits), dtype=[('read', '<u8'), ('offset', '<u4'), ('abund', 'u1', (_lib.MAX_SAMPLES,))])
    local['read'] = hits['read'].astype(np.uint64) + np.uint64(read_base)
    local['offset'] = hits['offset']
    local['abund'] = hits['abund']
    if not td.is_initialized() or td.get_world_size(group) == 1:
        return local
    parts = [None] * td.get_world_size(group)
    td.all_gather_object(parts, local, group=group)
    return np.concatenate(parts)


# ----------------------------------------------------------------------------------------------
# Plan B: bin-range-sharded sketches -- for sketches that do not fit one GPU.
# Every rank holds 1/world of every table; reads stay sharded across ranks; what crosses NVLink
# is the hash stream (8 bytes per k-mer position, one NCCL all-gather per batch) and, for the
# novel scan, one byte per position and sample (MIN all-reduce of the partial abundances).

class HashStream(object):
    """The canonical hashes of this rank's batch, gathered from every rank (device tensors)."""

    def __init__(self, sketch, bases, offsets, num_bands=None, band=None, group=None):
        import torch
        td = dist()
        bases, offsets = _lib.as_u8(bases), _lib.as_u64(offsets)
        self.world = td.get_world_size(group) if td.is_initialized() else 1
        self.rank = td.get_rank(group) if td.is_initialized() else 0
        dev = torch.device('cuda', sketch.device)
        total = int(offsets[-1]) if len(offsets) else 0
        cap = torch.tensor([(total + 1023) // 1024 * 1024], dtype=torch.int64, device=dev)
        if self.world > 1:
            td.all_reduce(cap, op=td.ReduceOp.MAX, group=group)
        self.cap = max(int(cap.item()), 1024)
        mine_h = torch.zeros(self.cap, dtype=torch.int64, device=dev)
        mine_v = torch.zeros(self.cap // 32 + 1, dtype=torch.int32, device=dev)
        npos, nk = c_uint64(), c_uint64()
        torch.cuda.synchronize(dev)
        check(lib().kv_hash_batch_dev(sketch._hasher, sketch.ksize(), bases.ctypes.data, offsets.ctypes.data,
                                      len(offsets) - 1, _lib.MEM_HOST, int(num_bands or 0), int(band or 0), sketch.device,
                                      mine_h.data_ptr(), mine_v.data_ptr(), self.cap, byref(npos), byref(nk)))
        self.n_kmers = nk.value
        if self.world > 1:
            self.hashes = [torch.empty_like(mine_h) for _ in range(self.world)]
            self.valid = [torch.empty_like(mine_v) for _ in range(self.world)]
            td.all_gather(self.hashes, mine_h, group=group)
            td.all_gather(self.valid, mine_v, group=group)
            torch.cuda.synchronize(dev)
        else:
            self.hashes, self.valid = [mine_h], [mine_v]
        self.device = dev


class ShardedSketch(object):
    """A khmer-style sketch whose tables are split by bin range over the ranks of `group`.
    `consume_batch` takes THIS rank's reads; afterwards every rank's shard has seen the k-mers
    of all ranks.  `save` writes one ordinary OXLI file, byte-identical to the unsharded one."""

    def __init__(self, cls, ksize, starting_size, n_tables, primes=None, group=None):
        td = dist()
        self.group = group
        self.world = td.get_world_size(group) if td.is_initialized() else 1
        self.rank = td.get_rank(group) if td.is_initialized() else 0
        sizes = [int(p) for p in primes] if primes else _lib.primes_below(int(starting_size), int(n_tables))
        arr = (c_uint64 * len(sizes))(*sizes)
        handle = c_void_p()
        check(lib().kv_sketch_create_shard(cls._hasher, cls._bits, int(ksize), len(sizes), arr, self.rank, self.world,
                                           _lib.current_device(), byref(handle)))
        self.local = cls(0, 0, 0, _handle=handle)

    def ksize(self):
        return self.local.ksize()

    def hashsizes(self):
        return self.local.hashsizes()

    def consume_batch(self, bases, offsets, num_bands=None, band=None):
        """Count this rank's reads into the sharded sketch; returns the k-mers counted by all ranks."""
        import torch
        td = dist()
        stream = HashStream(self.local, bases, offsets, num_bands, band, self.group)
        for h, v in zip(stream.hashes, stream.valid):
            check(lib().kv_add_hashes_dev(self.local._h, h.data_ptr(), v.data_ptr(), stream.cap))
        total = torch.tensor([stream.n_kmers], dtype=torch.int64, device=stream.device)
        if self.world > 1:
            td.all_reduce(total, op=td.ReduceOp.SUM, group=self.group)
        return int(total.item())

    def n_occupied(self):
        import torch
        td = dist()
        n = torch.tensor([self.local.n_occupied()], dtype=torch.int64, device=torch.device('cuda', self.local.device))
        if self.world > 1:
            td.all_reduce(n, op=td.ReduceOp.SUM, group=self.group)
        return int(n.item())

    def counts(self, stream):
        """Abundance of every position of `stream` (this rank's part): partial minima over the
        buckets held here, MIN-all-reduced over the shards."""
        import torch
        td = dist()
        parts = []
        for h, v in zip(stream.hashes, stream.valid):
            c = torch.empty(stream.cap, dtype=torch.uint8, device=stream.device)
            check(lib().kv_get_hashes_dev(self.local._h, h.data_ptr(), v.data_ptr(), stream.cap, c.data_ptr()))
            parts.append(c)
        allc = torch.stack(parts)
        if self.world > 1:
            td.all_reduce(allc, op=td.ReduceOp.MIN, group=self.group)
        mine = allc[stream.rank].contiguous()
        torch.cuda.synchronize(stream.device)   # the library's kernels run on its own stream, not torch's
        return mine

    def save(self, filename):
        """One OXLI v4 file (on a filesystem all ranks share), shards appended in rank order."""
        td = dist()
        occupied = self.n_occupied()
        path = str(filename).encode()

        def turn(piece, t, writer):
            if self.rank == writer:
                check(lib().kv_sketch_save_part(self.local._h, path, piece, t, occupied))
            if self.world > 1:
                td.barrier(group=self.group)
        turn(0, 0, 0)
        for t in range(len(self.hashsizes())):
            turn(1, t, 0)
            for r in range(self.world):
                turn(2, t, r)
        turn(3, 0, 0)


def novel_batch_sharded(cases, ctrls, bases, offsets, case_min, ctrl_max, screen=None, num_bands=None, band_minus_1=0,
                        max_hits=None):
    """kevlar_b200.khmer.novel_batch for ShardedSketch samples: this rank's reads against sketches
    spread over all ranks.  Same return value (hits, read_flags, discard_pos) for this rank's reads."""
    samples = list(cases) + list(ctrls)
    like = samples[0].local
    bases, offsets = _lib.as_u8(bases), _lib.as_u64(offsets)
    stream = HashStream(like, bases, offsets, group=samples[0].group)
    counts = [s.counts(stream) for s in samples]
    n_reads = len(offsets) - 1
    ptrs = (c_void_p * len(counts))(*[c.data_ptr() for c in counts])
    flags = np.zeros(max(1, n_reads), dtype=np.uint8)
    discard = np.full(max(1, n_reads), 0xffffffff, dtype=np.uint32)
    if max_hits is None:
        max_hits = max(4096, len(bases) // 64)
    while True:
        hits = np.empty(max_hits, dtype=_lib.HIT_DTYPE)
        n = c_uint64()
        rc = lib().kv_novel_from_counts(like._h, len(cases), len(ctrls), ptrs, bases.ctypes.data, offsets.ctypes.data,
                                        n_reads, _lib.MEM_HOST, int(case_min), int(ctrl_max), int(screen or 0),
                                        int(num_bands or 0), int(band_minus_1), hits.ctypes.data, max_hits, byref(n),
                                        flags.ctypes.data, discard.ctypes.data if screen else None)
        if rc == _lib.KV_EOVERFLOW:
            max_hits = int(n.value) + 1024
            continue
        check(rc)
        return hits[:n.value], flags[:n_reads], discard[:n_reads]


# ----------------------------------------------------------------------------------------------
# Plan B, second design: spanning sketches.  Every table is ONE virtual address range mapped on every rank,
# piece r in the HBM of rank r (CUDA VMM, include/kvsketch.h: kv_sketch_create_span).  Reads stay sharded;
# what crosses NVLink per update is a 2-byte offset (pulled by the rank that owns the region, inside the
# kernel that applies it) instead of an 8-byte hash broadcast to every rank; queries are plain loads
# that land in whichever GPU holds the page, so khmer.novel_batch, get, save, n_occupied work unchanged.

def _exchange_fds(my_fds, group=None):
    """All-to-all exchange of file descriptors between the ranks of one node (SCM_RIGHTS over Unix
    sockets).  Returns {rank: [fds]} for every peer."""
    import socket
    import struct
    import uuid
    td = dist()
    world, rank = td.get_world_size(group), td.get_rank(group)
    tag = [uuid.uuid4().hex if rank == 0 else None]
    td.broadcast_object_list(tag, src=0, group=group)
    path = lambda r: '/tmp/kvspan_{}_{}.sock'.format(tag[0], r)   # noqa: E731
    server = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    server.bind(path(rank))
    server.listen(world)
    td.barrier(group=group)
    try:
        for p in range(world):
            if p == rank:
                continue
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(path(p))
            socket.send_fds(c, [struct.pack('i', rank)], list(my_fds))
            c.close()
        got = {}
        for _ in range(world - 1):
            conn, _ = server.accept()
            msg, fds, _, _ = socket.recv_fds(conn, 16, len(my_fds))
            got[struct.unpack('i', msg[:4])[0]] = list(fds)
            conn.close()
    finally:
        server.close()
        td.barrier(group=group)
        os.unlink(path(rank))
    return got


class SpanningSketch(object):
    """A khmer-style sketch whose tables span the HBM of all ranks of `group`.  `.sketch` is an ordinary
    kevlar_b200.khmer sketch object for everything that reads (get, novel_batch, save, n_occupied, ...);
    `consume_batch`, `clear` and `close` are COLLECTIVE."""

    def __init__(self, cls, ksize, starting_size, n_tables, primes=None, group=None, chunk_positions=0):
        td = dist()
        self.group = group
        self.world = td.get_world_size(group) if td.is_initialized() else 1
        self.rank = td.get_rank(group) if td.is_initialized() else 0
        sizes = [int(p) for p in primes] if primes else _lib.primes_below(int(starting_size), int(n_tables))
        arr = (c_uint64 * len(sizes))(*sizes)
        handle, fds, xchg = c_void_p(), (c_int * len(sizes))(), (ctypes.c_uint8 * 128)()
        device = _lib.current_device()
        check(lib().kv_sketch_create_span(cls._hasher, cls._bits, int(ksize), len(sizes), arr, self.rank, self.world, device,
                                          int(chunk_positions), byref(handle), fds, xchg))
        self.sketch = cls(0, 0, 0, _handle=handle)
        if self.world > 1:
            peer_fds = _exchange_fds(list(fds), group)
            handles = [None] * self.world
            td.all_gather_object(handles, bytes(xchg), group=group)
            for r in range(self.world):
                if r != self.rank:
                    check(lib().kv_sketch_span_attach(handle, r, (c_int * len(sizes))(*peer_fds[r]),
                                                      (ctypes.c_uint8 * 128)(*handles[r])))
                    for f in peer_fds[r]:
                        os.close(f)
        for f in fds:
            os.close(f)
        check(lib().kv_sketch_span_ready(handle))
        chunk = c_uint64()
        check(lib().kv_sketch_span_info(handle, None, None, None, byref(chunk)))
        self.chunk_positions = chunk.value
        if self.world > 1:
            self._sync = _peer_sync(device, group)
            td.barrier(group=group)
        else:   # a one-rank "collective": same code path, the barriers have nobody to wait for
            ps, unused = c_void_p(), (ctypes.c_uint8 * 64)()
            check(lib().kv_peer_sync_create(device, 0, 1, byref(ps), unused))
            self._sync = ps

    def ksize(self):
        return self.sketch.ksize()

    def hashsizes(self):
        return self.sketch.hashsizes()

    def n_occupied(self):
        return self.sketch.n_occupied()

    def save(self, filename):
        """One ordinary OXLI file, written by rank 0 (it reads the peers' pieces over NVLink)."""
        td = dist()
        if self.rank == 0:
            self.sketch.save(filename)
        if self.world > 1:
            td.barrier(group=self.group)

    def clear(self):
        self.sketch.clear()
        if self.world > 1:
            _lib.sync(self.sketch.device)
            dist().barrier(group=self.group)

    def consume_batch(self, bases, offsets, num_bands=None, band=None, mask=None, threshold=0, consume_masked=False,
                      where=_lib.MEM_HOST, n_positions=None):
        """Count THIS rank's reads; afterwards the sketch has seen the reads of all ranks.  Returns the number
        of k-mers counted by all ranks."""
        import torch
        td = dist()
        if where == _lib.MEM_HOST:
            bases, offsets = _lib.as_u8(bases), _lib.as_u64(offsets)
            bptr, optr, n_reads, n_pos = bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, len(bases)
        else:
            bptr, (optr, n_reads) = bases, offsets
            n_pos = int(n_positions)
        mine = -(-((n_pos + 1023) // 1024 * 1024) // self.chunk_positions)
        dev = torch.device('cuda', self.sketch.device)
        chunks = torch.tensor([mine], dtype=torch.int64, device=dev)
        if self.world > 1:
            td.all_reduce(chunks, op=td.ReduceOp.MAX, group=self.group)
        n = c_uint64()
        check(lib().kv_consume_batch_span(self.sketch._h, self._sync, bptr, optr, n_reads, where, int(chunks.item()),
                                          int(num_bands or 0), int(band or 0), mask._h if mask is not None else None,
                                          int(threshold), int(bool(consume_masked)), byref(n)))
        if self.world == 1:
            return n.value
        total = torch.tensor([n.value], dtype=torch.int64, device=dev)
        td.all_reduce(total, op=td.ReduceOp.SUM, group=self.group)
        return int(total.item())

    def close(self):
        """COLLECTIVE: nobody touches the sketch any more; unmap and free everywhere."""
        if self.sketch is None:
            return
        _lib.sync(self.sketch.device)
        if self.world > 1:
            dist().barrier(group=self.group)
        self.sketch = None   # kv_sketch_destroy: unmaps the peers' pieces, frees mine
        if self.world > 1:
            dist().barrier(group=self.group)
        else:
            lib().kv_peer_sync_destroy(self._sync)
        self._sync = None
