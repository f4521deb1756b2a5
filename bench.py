#!/usr/bin/env python
"""Headline benchmark: k-mers/s for `kevlar count` x3 + `kevlar novel` on a synthetic trio.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one trio batch: zero three sketches, count the
proband / mother / father reads into them (hash + exact-n_unique bookkeeping + saturating
increments), [N>1: merge the per-GPU partial sketches], then scan the proband reads against
all three sketches (kevlar novel).  k-mers per step = sum over the 3 samples of counted
k-mer positions + scanned k-mer positions.

`value` / `e2e` workload = BASELINE.json configs[1] (C2): 1 Mbp genome, 30x => 300,000 reads x 100 bp
per sample, k=31, 64 MB / 4-table 8-bit sketches.  Under torchrun every rank gets its own 300k-read
shard per sample (weak scaling; reads sharded, sketches merged, novel shard-local).

`c3` = BASELINE.json configs[2] in the same line: 100 Mbp genome at 30x => 30 M reads per sample,
3 x 4 GB sketches (HBM-resident), reads generated on the device and sharded over the ranks (STRONG
scaling: the same 90 M reads at every N), partial sketches merged over NVLink, novel shard-local.

`value`  inputs already resident in HBM when the timed region starts.
`e2e`    the same step through the public host-buffer API (pinned host batches, H2D copies and
         the D2H of the results inside the timed region).
`--impl reference` times the reference's CPU algorithm (the oracle port -- khmer itself is not
         vendored) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

K = 31
MEMORY = 64e6
READS_PER_SAMPLE = 300000
READ_LEN = 100
CASE_MIN, CTRL_MAX = 6, 1
N_TABLES = 4

C3_GENOME = 100000000
C3_READS = 30000000
C3_MEMORY = 4e9


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--hasher', default='murmur', choices=['murmur', 'twobit'],
                    help='murmur = Counttable (what `kevlar count` builds); twobit = Countgraph')
    ap.add_argument('--counter-size', type=int, default=8, choices=[8, 4, 1],
                    help='bits per counter at the same 64 MB per sketch (4 = SmallCount*, 1 = Node*); the headline is 8')
    ap.add_argument('--merge', default='p2p', choices=['allreduce', 'allgather', 'p2p', 'p2p_host', 'sharded', 'span'],
                    help="N>1: how per-GPU work is combined; 'sharded' = plan B (hash all-gather), 'span' = plan B with "
                         "spanning sketches (tables spread over all GPUs, exchange fused into the apply kernel)")
    ap.add_argument('--reads-per-sample', type=int, default=READS_PER_SAMPLE)
    ap.add_argument('--no-unique', action='store_true', help='skip the exact n_unique_kmers bookkeeping')
    ap.add_argument('--no-overlap', action='store_true', help="N > 1: merge all sketches after the last sample instead of "
                    "merging each sample's sketch on the merge lane while the next sample is counted")
    ap.add_argument('--no-files', action='store_true', help='N = 1: skip the file-based path (FASTQ files -> sketch files -> augmented FASTQ)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-variants', action='store_true')
    ap.add_argument('--no-c3', action='store_true', help='skip the config-3 (100 Mbp, 3 x 4 GB sketches) part')
    ap.add_argument('--c3-reads', type=int, default=C3_READS, help='reads per sample of the config-3 part')
    ap.add_argument('--c3-genome', type=int, default=C3_GENOME)
    ap.add_argument('--c3-memory', type=float, default=C3_MEMORY)
    ap.add_argument('--c3-steps', type=int, default=2)
    ap.add_argument('--c3-no-parity', action='store_true')
    ap.add_argument('--c5', action='store_true', help='also time the banded chain (8 bands + unband + filter) over FASTQ files; default at N>1')
    ap.add_argument('--no-c5', action='store_true')
    ap.add_argument('--c4', action='store_true', help='also run the config-4-shaped part: 4-bit spanning sketches across all GPUs')
    ap.add_argument('--c4-gb-per-gpu', type=float, default=16.0, help='GB of each of the 3 sketches held per GPU')
    ap.add_argument('--c4-reads', type=int, default=C3_READS, help='reads per sample (all ranks together)')
    ap.add_argument('--c4-chunk', type=int, default=1 << 30, help='positions per rank and exchange round of the spanning update')
    return ap.parse_args()


SKETCH_CLASS = {('murmur', 8): 'Counttable', ('murmur', 4): 'SmallCounttable', ('murmur', 1): 'Nodetable',
                ('twobit', 8): 'Countgraph', ('twobit', 4): 'SmallCountgraph', ('twobit', 1): 'Nodegraph'}


def sketch_shape(args):
    """(khmer class name, buckets per table): kevlar/count.py:33 sizes tables as memory / 4 * buckets per byte."""
    return SKETCH_CLASS[(args.hasher, args.counter_size)], MEMORY / N_TABLES * (8 // args.counter_size)


def workload_name(args):
    kind = '(MurmurHash3 canonical)' if args.hasher == 'murmur' else '(2-bit canonical)'
    return ('C2: gentrio-style synthetic trio, 1 Mbp genome at 30x = {} reads x {} bp per sample, k={}, '
            '3 x {:.0f} MB {}-bit {} {} with {} tables; step = count x3 + novel scan of the proband reads '
            '(case_min {}, ctrl_max {})').format(args.reads_per_sample, READ_LEN, K, MEMORY / 1e6, args.counter_size,
                                                  sketch_shape(args)[0], kind, N_TABLES, CASE_MIN, CTRL_MAX)


def config_of(args, world):
    """The workload description; BOTH arms print exactly this dict."""
    return {
        'workload': workload_name(args),
        'reads_sharded': world > 1,
        'exact_n_unique_tracking': not args.no_unique,
        'l2': 'inputs larger than L2: per step 120 MB of reads + 192 MB of sketches + 240 MB of hash scratch '
              'stream through a 126 MB L2; no explicit flush',
    }


def kmers_per_step(trio):
    total = 0
    for bases, offs in trio:
        lens = np.diff(offs.astype(np.int64))
        total += int(np.maximum(lens - K + 1, 0).sum())
    lens = np.diff(trio[0][1].astype(np.int64))
    return total + int(np.maximum(lens - K + 1, 0).sum())


# ----------------------------------------------------------------------------- clocks

class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 2 ms from a
    thread (a 10-step timed region lasts ~80 ms, too short for `nvidia-smi -lms`), nvidia-smi as the
    fallback when the NVML binding is missing."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []
        self.stop_flag = False

    def _physical_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
        if vis:
            try:
                return int(vis.split(',')[self.index])
            except (ValueError, IndexError):
                pass
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, reasons))
            except Exception:
                pass
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            nv = self.nvml
            try:
                smax = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            except Exception:
                smax = None
            if not self.samples:
                return {'sm_mhz': None, 'sm_max_mhz': smax, 'reasons': ['no samples']}
            names = (('hw_slowdown', 'nvmlClocksThrottleReasonHwSlowdown'),
                     ('hw_thermal_slowdown', 'nvmlClocksThrottleReasonHwThermalSlowdown'),
                     ('sw_thermal_slowdown', 'nvmlClocksThrottleReasonSwThermalSlowdown'),
                     ('sw_power_cap', 'nvmlClocksThrottleReasonSwPowerCap'))
            reasons = set()
            for _, bits in self.samples:
                for label, const in names:
                    if bits & getattr(nv, const, 0):
                        reasons.add(label)
            return {'sm_mhz': float(np.median([s for s, _ in self.samples])), 'sm_max_mhz': smax,
                    'samples': len(self.samples), 'reasons': sorted(reasons), 'source': 'NVML, 2 ms poll during the timed region'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(smax)), 'samples': len(sm),
                'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arms

def cpu_step(ko, trio, threads, shape=('Counttable', MEMORY / N_TABLES)):
    """count x3 + novel with the oracle; returns (seconds, n_hits, per-sample k-mer counts)."""
    t0 = time.perf_counter()
    sks, counted = [], []
    for bases, offs in trio:
        sk = getattr(ko, shape[0])(K, shape[1], N_TABLES)
        counted.append(sk.consume_batch(bases, offs, threads=threads))
        sks.append(sk)
    hits, _ = ko.novel_batch(sks[:1], sks[1:], trio[0][0], trio[0][1], CASE_MIN, CTRL_MAX, threads=threads)
    return time.perf_counter() - t0, hits, counted, sks


def reference_python_novel(ko, trio, max_reads=1500, sks=None):
    """The reference's REAL novel loop shape (kevlar/novel.py:123-169: a Python loop over reads and over
    k-mers with one `get` per sample and k-mer, single-threaded) over the oracle's khmer-shaped
    objects, on the first `max_reads` proband reads.  Returns (k-mers/s, reads used)."""
    if sks is None:
        sks = []
        for bases, offs in trio:
            sk = ko.Counttable(K, MEMORY / N_TABLES, N_TABLES)
            sk.consume_batch(bases, offs, threads=os.cpu_count() or 1)
            sks.append(sk)
    bases, offs = trio[0]
    n = min(max_reads, len(offs) - 1)
    seqs = [bases[int(offs[i]):int(offs[i + 1])].tobytes().decode() for i in range(n)]
    case, ctrls = sks[0], sks[1:]
    t0 = time.perf_counter()
    nk = 0
    for seq in seqs:
        for kmer in case.get_kmers(seq):
            nk += 1
            if case.get(kmer) < CASE_MIN:
                continue
            for ct in ctrls:
                if ct.get(kmer) > CTRL_MAX:
                    break
    return nk / (time.perf_counter() - t0), n


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import khmer_oracle as ko
    from kevlar_b200 import simtrio
    cores = os.cpu_count() or 1
    n_reads = args.reads_per_sample if cores >= 8 else args.reads_per_sample // 3
    trio = simtrio.simulate_trio(1000000, reads_per_sample=n_reads)
    nk = kmers_per_step(trio)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_step(ko, trio, cores, sketch_shape(args))
    times = []
    for _ in range(args.steps):
        times.append(cpu_step(ko, trio, cores, sketch_shape(args))[0])
    total = sum(times)
    value = nk * args.steps / total
    sample = ('{} of {} reads/sample per step; count in {} pthreads pulling read chunks (khmer model), novel scan in C '
              'with {} threads (the reference runs it single-threaded in Python)').format(
                  n_reads, args.reads_per_sample, cores, cores)
    py_rate, py_reads = reference_python_novel(ko, trio)
    line = {
        'impl': 'reference', 'metric': 'kmers_per_sec_count_plus_novel', 'value': value, 'unit': 'k-mers/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': config_of(args, max(1, int(os.environ.get('WORLD_SIZE', '1')))),
        'arm': 'CPU oracle port of the khmer path (khmer is not vendored)',
        'cpu_baseline': {'value': value, 'unit': 'k-mers/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'reference_python_novel_loop': {
            'value': py_rate, 'unit': 'k-mers/s', 'cores': 1,
            'sample': 'kevlar/novel.py:123-169 as the reference runs it: a single-threaded Python loop with one get() per '
                      'sample and k-mer over the oracle sketches, first {} proband reads'.format(py_reads)},
        'e2e': {'value': value, 'unit': 'k-mers/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm

class GpuTrio(object):
    """Sketches + per-step driver for one rank."""

    def __init__(self, args, trio, world):
        import torch
        import kevlar_b200
        from kevlar_b200 import _lib, khmer, multigpu
        self.torch, self.kv, self.lib, self.khmer, self.multigpu = torch, kevlar_b200, _lib, khmer, multigpu
        self.args, self.world = args, world
        self.device = _lib.current_device()
        name, buckets = sketch_shape(args)
        cls = getattr(khmer, name)
        self.sharded = world > 1 and args.merge == 'sharded'
        self.span = world > 1 and args.merge == 'span'
        if self.sharded:   # plan B: every rank holds 1/world of every table; reads stay sharded
            self.sketches = [multigpu.ShardedSketch(cls, K, buckets, N_TABLES) for _ in range(3)]
        elif self.span:    # plan B, spanning sketches: one virtual table over the HBM of all ranks
            self.spans = [multigpu.SpanningSketch(cls, K, buckets, N_TABLES) for _ in range(3)]
            self.sketches = [sp.sketch for sp in self.spans]
        else:
            self.sketches = [cls(K, buckets, N_TABLES) for _ in range(3)]
        if args.no_unique and not self.sharded and not self.span:
            for sk in self.sketches:
                sk.set_unique_tracking(False)
        self.trio = trio
        dev = torch.device('cuda', self.device)
        # device-resident copies (value arm) and pinned host copies (e2e arm)
        self.dev = [(torch.from_numpy(b).to(dev), torch.from_numpy(o.view(np.int64)).to(dev)) for b, o in trio]
        self.pinned = [(torch.from_numpy(b).pin_memory(), torch.from_numpy(o.view(np.int64)).pin_memory())
                       for b, o in trio]
        self.stream = torch.cuda.ExternalStream(_lib.stream_ptr(self.device), device=dev)
        self.last_hits = None

    def plain_sketches(self):
        return [] if (self.sharded or self.span) else list(self.sketches)

    def step_span(self, resident):
        khmer = self.khmer
        for sp in self.spans:
            sp.clear()
        for i, sp in enumerate(self.spans):
            if resident:
                b, o = self.dev[i]
                sp.consume_batch(b.data_ptr(), (o.data_ptr(), o.numel() - 1), where=khmer.MEM_DEVICE, n_positions=b.numel())
            else:
                b, o = self.pinned[i]
                sp.consume_batch(b.numpy(), o.numpy().view(np.uint64))
        if resident:
            b, o = self.dev[0]
            hits = khmer.novel_batch(self.sketches[:1], self.sketches[1:], b.data_ptr(), (o.data_ptr(), o.numel() - 1, b.numel()),
                                     CASE_MIN, CTRL_MAX, where=khmer.MEM_DEVICE)[0]
        else:
            b, o = self.pinned[0]
            hits = khmer.novel_batch(self.sketches[:1], self.sketches[1:], b.numpy(), o.numpy().view(np.uint64), CASE_MIN,
                                     CTRL_MAX)[0]
        self.last_hits = hits
        return hits

    def step_sharded(self):
        # host buffers in both arms: hash locally, exchange, apply to the local bin ranges
        for sk in self.sketches:
            sk.local.clear()
        for i, sk in enumerate(self.sketches):
            b, o = self.pinned[i]
            sk.consume_batch(b.numpy(), o.numpy().view(np.uint64))
        b, o = self.pinned[0]
        hits, _, _ = self.multigpu.novel_batch_sharded(self.sketches[:1], self.sketches[1:], b.numpy(),
                                                       o.numpy().view(np.uint64), CASE_MIN, CTRL_MAX)
        self.last_hits = hits
        return hits

    def step(self, resident):
        if self.sharded:
            return self.step_sharded()
        if self.span:
            return self.step_span(resident)
        khmer = self.khmer
        for sk in self.sketches:
            sk.clear()
        if resident:
            batches = [(b.data_ptr(), (o.data_ptr(), o.numel() - 1)) for b, o in self.dev]
        else:
            batches = [(b.numpy(), o.numpy().view(np.uint64)) for b, o in self.pinned]
        # N = 1: tracked consume.  N > 1: untracked consume of this rank's shards, the exact n_unique_kmers of the
        # sharded stream (occupancy exchange + first-touch passes), then the merge
        self.multigpu.count_sharded(self.sketches, batches, how=self.args.merge, where=khmer.MEM_DEVICE if resident else khmer.MEM_HOST,
                                    exact_unique=not self.args.no_unique, overlap=not self.args.no_overlap)
        if resident:
            b, o = self.dev[0]
            hits, flags, _ = khmer.novel_batch(self.sketches[:1], self.sketches[1:], b.data_ptr(),
                                               (o.data_ptr(), o.numel() - 1, b.numel()), CASE_MIN, CTRL_MAX,
                                               where=khmer.MEM_DEVICE)
        else:
            b, o = self.pinned[0]
            hits, flags, _ = khmer.novel_batch(self.sketches[:1], self.sketches[1:], b.numpy(),
                                               o.numpy().view(np.uint64), CASE_MIN, CTRL_MAX)
        self.last_hits = hits
        return hits

    def timed(self, steps, resident, barrier):
        """K steps bracketed by barrier + synchronize; device time from CUDA events on the
        library's stream."""
        torch = self.torch
        barrier()
        torch.cuda.synchronize()
        self.lib.sync(self.device)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(self.stream)
        for _ in range(steps):
            self.step(resident)
        end.record(self.stream)
        self.lib.sync(self.device)
        torch.cuda.synchronize()
        barrier()
        return start.elapsed_time(end)   # ms


def load_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def load_traffic(kernel):
    """DRAM bytes per launch from the committed `ncu --set full` capture, if any."""
    path = os.path.join(REPO, 'profiles', 'traffic.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get(kernel)
    return None


def load_atomic_peak(table_mb=64):
    """Measured L2 atomic throughput for a table of this size (tools/atomic_microbench.cu, committed
    under profiles/): G updates/s for ATOM.ADD (the primitive the update kernel issues) and RED.ADD."""
    path = os.path.join(REPO, 'profiles', 'r01_atomic_microbench.csv')
    peaks = {}
    if os.path.exists(path):
        with open(path) as fh:
            for line in fh:
                f = line.strip().split(',')
                if len(f) >= 4 and f[1] == str(table_mb) and f[0] in ('atom_add', 'red_add'):
                    peaks[f[0]] = float(f[3])
    return peaks


def sketch_checksum(torch, sketch, local_only=False):
    """64-bit checksum of a sketch's table storage, computed on the device in 256 MB pieces.  Spanning
    sketches with `local_only`: only the pieces that live in THIS rank's HBM (all ranks together cover the
    sketch; nobody drags the whole thing over NVLink)."""
    import ctypes
    from kevlar_b200 import _lib, multigpu
    flat = multigpu.GpuSketchAdapter(sketch).flat_tensor()
    ranges = [(0, flat.numel())]
    if local_only:
        nt = len(sketch.hashsizes())
        rank, world = ctypes.c_int(), ctypes.c_int()
        piece = (ctypes.c_uint64 * nt)()
        _lib.check(_lib.lib().kv_sketch_span_info(sketch._h, ctypes.byref(rank), ctypes.byref(world), piece, None))
        ranges, off = [], 0
        for t in range(nt):
            ranges.append((off + rank.value * piece[t], off + (rank.value + 1) * piece[t]))
            off += piece[t] * world.value
    weights = torch.arange(1, 1025, dtype=torch.int64, device=flat.device)
    total, step = 0, 256 << 20
    for lo, hi in ranges:
        for a in range(lo, hi, step):
            part = flat[a:min(a + step, hi)]
            n8 = part.numel() // 8192 * 8192
            if n8:
                total ^= int((part[:n8].view(torch.int64).view(-1, 1024) * weights).sum().item()) & 0xffffffffffffffff
            if part.numel() > n8:
                total ^= int(part[n8:].sum(dtype=torch.int64).item())
    return total - (1 << 64) if total >= (1 << 63) else total   # as a signed 64-bit value (it travels in int64 tensors)


def multi_rank_parity(args, runner, rank, world, torch, multigpu, phase):
    """N>1: what did the merge produce?  (a) every rank holds byte-identical merged sketches (device
    checksums compared over the ranks); (b) rank 0 counts ALL ranks' shards with the multi-threaded oracle
    and compares its table bytes with rank 0's merged GPU sketches, and the oracle's novel hits with
    the hits gathered from all ranks.  Collective: every rank calls it."""
    td = torch.distributed
    sums = torch.tensor([sketch_checksum(torch, sk) for sk in runner.sketches], dtype=torch.int64, device='cuda')
    allsums = [torch.empty_like(sums) for _ in range(world)]
    td.all_gather(allsums, sums)
    identical = all(bool((s == allsums[0]).all()) for s in allsums)
    hits = multigpu.gather_hits(runner.last_hits, rank * args.reads_per_sample)
    if rank != 0:
        return None
    from oracle import khmer_oracle as ko
    from kevlar_b200 import simtrio
    cores = os.cpu_count() or 1
    name, buckets = sketch_shape(args)
    osk = [getattr(ko, name)(K, buckets, N_TABLES) for _ in range(3)]
    case_b, case_o = [], []
    for r in range(world):
        trio = simtrio.simulate_trio(1000000, reads_per_sample=args.reads_per_sample, seed_offset=1000 * r)
        for sk, (b, o) in zip(osk, trio):
            sk.consume_batch(b, o, threads=cores)
        case_b.append(trio[0][0])
        case_o.append(trio[0][1][:-1] + np.uint64(r * args.reads_per_sample * READ_LEN))
    case_o.append(np.array([world * args.reads_per_sample * READ_LEN], dtype=np.uint64))
    ohits, _ = ko.novel_batch(osk[:1], osk[1:], np.concatenate(case_b), np.concatenate(case_o), CASE_MIN, CTRL_MAX,
                              threads=cores)
    phase('oracle counted all {} shards'.format(world))
    same_tables = all(g.table_bytes(t) == c.table_bytes(t) for g, c in zip(runner.sketches, osk) for t in range(N_TABLES))
    hits = hits[np.lexsort((hits['offset'], hits['read']))]
    same_hits = len(hits) == len(ohits) and bool((hits['read'] == ohits['read']).all()) and \
        bool((hits['offset'] == ohits['offset']).all()) and bool((hits['abund'][:, :3] == ohits['abund'][:, :3]).all())
    # n_unique_kmers of the sharded stream (rank order = file order) against the single-threaded oracle -- the
    # order-dependent number cannot come from the multi-threaded count above; one sample, small worlds only (time)
    unique_note = ''
    same_unique = True
    if world <= 2 and not args.no_unique:
        seq = getattr(ko, name)(K, buckets, N_TABLES)
        for r in range(world):
            trio = simtrio.simulate_trio(1000000, reads_per_sample=args.reads_per_sample, seed_offset=1000 * r)
            seq.consume_batch(trio[0][0], trio[0][1])
        same_unique = runner.sketches[0].n_unique_kmers() == seq.n_unique_kmers()
        unique_note = ', n_unique_kmers {} == single-threaded oracle'.format(seq.n_unique_kmers())
        phase('oracle n_unique (single thread) done')
    ok = identical and same_tables and same_hits and same_unique
    text = ('bit-exact at N={}: merged sketches identical on all ranks (device checksums), rank 0 tables == oracle over all '
            '{} shards, {} gathered hits == oracle{}').format(world, world, len(ohits), unique_note) if ok else \
        'MISMATCH (ranks identical: {}, tables: {}, hits: {} [{} vs {}], n_unique: {})'.format(
            identical, same_tables, same_hits, len(hits), len(ohits), same_unique)
    return ok, text


# ----------------------------------------------------------------------------- config 3

def run_c3(args, rank, world, barrier, phase):
    """BASELINE config 3: 100 Mbp trio at 30x, 3 x 4 GB 8-bit Counttables (HBM-resident), reads drawn on
    the device and sharded over the ranks -- the same 90 M reads at every N (strong scaling).  Returns
    this rank's measurements; collective (every rank calls it)."""
    import torch
    from kevlar_b200 import _lib, khmer, multigpu, simtrio
    td = torch.distributed
    device = _lib.current_device()
    dev = torch.device('cuda', device)
    n_total = args.c3_reads
    lo, hi = multigpu.shard_bounds(n_total, rank, world)
    dtrio = simtrio.device_trio(args.c3_genome, n_total, rank, world, read_len=READ_LEN)
    phase('c3: {} reads x 3 samples drawn on the device'.format(hi - lo))
    buckets = args.c3_memory / N_TABLES
    sketches = [khmer.Counttable(K, buckets, N_TABLES) for _ in range(3)]
    for sk in sketches:
        sk.set_unique_tracking(False)
    stream = torch.cuda.ExternalStream(_lib.stream_ptr(device), device=dev)
    nk_sample = (hi - lo) * (READ_LEN - K + 1)
    state = {}

    def count(trio_slice, tracked=False):
        for sk in sketches:
            sk.clear()
        if world > 1 and not tracked:
            # each sample's merge runs on the merge lane while the next sample is being counted
            multigpu.count_sharded(sketches, [(b.data_ptr(), (o.data_ptr(), o.numel() - 1)) for b, o in trio_slice],
                                   how='p2p', where=khmer.MEM_DEVICE, exact_unique=False, overlap=not args.no_overlap)
            return
        for sk, (b, o) in zip(sketches, trio_slice):
            sk.set_unique_tracking(tracked)
            sk.consume_batch(b.data_ptr(), (o.data_ptr(), o.numel() - 1), where=khmer.MEM_DEVICE, wait=False)
        if world > 1:
            multigpu.merge_sketches(sketches, how='p2p')

    def scan(trio_slice):
        b, o = trio_slice[0]
        state['hits'] = khmer.novel_batch(sketches[:1], sketches[1:], b.data_ptr(), (o.data_ptr(), o.numel() - 1, b.numel()),
                                          CASE_MIN, CTRL_MAX, where=khmer.MEM_DEVICE)[0]

    def step():
        count(dtrio)
        scan(dtrio)

    def timed(fn, n):
        barrier()
        torch.cuda.synchronize()
        _lib.sync(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        _lib.sync(device)
        torch.cuda.synchronize()
        barrier()
        return e0.elapsed_time(e1) / n

    step()   # warm-up (allocates the scratch, maps the peers)
    ms_step = timed(step, args.c3_steps)
    phase('c3: timed steps done')
    _lib.profile(1)
    ms_count = timed(lambda: count(dtrio), 1)
    prof_count = _lib.profile(1)
    ms_scan = timed(lambda: scan(dtrio), 1)
    prof_scan = _lib.profile(0)
    ms_count_tracked = timed(lambda: count(dtrio, tracked=True), 1)
    for sk in sketches:
        sk.set_unique_tracking(False)
    count(dtrio)
    scan(dtrio)
    _lib.sync(device)
    n_hits = len(state['hits'])
    # size-independent properties at full size: no counter saturated => every table sums to the number of
    # k-mers counted by ALL ranks; all ranks hold the same merged sketch
    props = {}
    sums_ok, sat = True, False
    for sk in sketches:
        sat = sat or int(multigpu.GpuSketchAdapter(sk).flat_tensor().max().item()) == 255
    if not sat:
        for sk in sketches:
            total = int(multigpu.GpuSketchAdapter(sk).flat_tensor().sum(dtype=torch.int64).item())
            sums_ok = sums_ok and total == N_TABLES * n_total * (READ_LEN - K + 1)
    props['table_sums_equal_kmers_counted'] = (sums_ok if not sat else 'not applicable: a counter saturated')
    sums = torch.tensor([sketch_checksum(torch, sk) for sk in sketches], dtype=torch.int64, device=dev)
    if world > 1:
        allsums = [torch.empty_like(sums) for _ in range(world)]
        td.all_gather(allsums, sums)
        props['merged_sketches_identical_on_all_ranks'] = all(bool((s == allsums[0]).all()) for s in allsums)
    props['sketch_checksums'] = [int(x) for x in sums.tolist()]
    hits_total = torch.tensor([n_hits], dtype=torch.int64, device=dev)
    if world > 1:
        td.all_reduce(hits_total)
    props['novel_hits_all_ranks'] = int(hits_total.item())
    phase('c3: profiles and full-size properties done')

    # parity on a bounded sample: the first 1/32 of every rank's shard, counted into the same (cleared)
    # 4 GB sketches, merged, scanned; rank 0 redraws every rank's sample on its own device, copies it to the
    # host and runs the oracle on the union
    parity = None
    if not args.c3_no_parity:
        sub = max(1000, (hi - lo) // 32)
        sub_trio = [(b[:sub * READ_LEN], o[:sub + 1]) for b, o in dtrio]
        count(sub_trio, tracked=False)
        scan(sub_trio)
        _lib.sync(device)
        hits = multigpu.gather_hits(state['hits'], 0)   # read indices stay rank-local; the rank is added below
        sub_hits = state['hits']
        counts = torch.tensor([len(sub_hits)], dtype=torch.int64, device=dev)
        all_counts = [torch.empty_like(counts) for _ in range(world)]
        if world > 1:
            td.all_gather(all_counts, counts)
        else:
            all_counts = [counts]
        if rank == 0:
            from oracle import khmer_oracle as ko
            cores = os.cpu_count() or 1
            osk = [ko.Counttable(K, buckets, N_TABLES) for _ in range(3)]
            case_parts = []
            haps = simtrio.trio_haplotypes(args.c3_genome)
            import ctypes
            for r in range(world):
                rlo, rhi = multigpu.shard_bounds(n_total, r, world)
                rsub = max(1000, (rhi - rlo) // 32)
                for si, (hp, seed) in enumerate(zip(haps, simtrio.READ_SEEDS)):
                    dh = [torch.from_numpy(np.ascontiguousarray(h)).to(dev) for h in hp]
                    ptrs = (ctypes.c_void_p * len(dh))(*[t.data_ptr() for t in dh])
                    lens = (ctypes.c_uint64 * len(dh))(*[t.numel() for t in dh])
                    bases = torch.empty(rsub * READ_LEN, dtype=torch.uint8, device=dev)
                    torch.cuda.synchronize(dev)
                    _lib.check(_lib.lib().kv_synth_reads(device, ptrs, lens, len(dh), rsub, rlo, READ_LEN, 0.005, int(seed),
                                                         bases.data_ptr(), None))
                    hb = bases.cpu().numpy()
                    ho = np.arange(rsub + 1, dtype=np.uint64) * np.uint64(READ_LEN)
                    osk[si].consume_batch(hb, ho, threads=cores)
                    if si == 0:
                        case_parts.append((hb, ho))
                    del dh, bases
            same_tables = all(g.table_bytes(t) == c.table_bytes(t) for g, c in zip(sketches, osk) for t in range(N_TABLES))
            o_total, same_hits, pos = 0, True, 0
            for r, (hb, ho) in enumerate(case_parts):
                ohits, _ = ko.novel_batch(osk[:1], osk[1:], hb, ho, CASE_MIN, CTRL_MAX, threads=cores)
                n_r = int(all_counts[r].item())
                mine = hits[pos:pos + n_r]
                pos += n_r
                mine = mine[np.lexsort((mine['offset'], mine['read']))]
                same_hits = same_hits and len(mine) == len(ohits) and bool((mine['read'] == ohits['read']).all()) and \
                    bool((mine['offset'] == ohits['offset']).all()) and \
                    bool((mine['abund'][:, :3] == ohits['abund'][:, :3]).all())
                o_total += len(ohits)
            ok = same_tables and same_hits
            parity = {'ok': ok, 'sample': 'first 1/32 of every rank\'s shard ({} reads per sample in all) counted into 3 x {:.0f} GB '
                                          'sketches, merged over {} rank(s), scanned'.format(
                                              sum(max(1000, (multigpu.shard_bounds(n_total, r, world)[1] -
                                                             multigpu.shard_bounds(n_total, r, world)[0]) // 32)
                                                  for r in range(world)), args.c3_memory / 1e9, world),
                      'result': 'bit-exact: 12 tables of rank 0 == oracle, {} hits == oracle'.format(o_total) if ok else
                                'MISMATCH (tables {}, hits {})'.format(same_tables, same_hits)}
            del osk
            phase('c3: parity sample checked against the oracle')
        barrier()

    # whole-job numbers: max over ranks
    t = torch.tensor([ms_step, ms_count, ms_scan, ms_count_tracked, prof_count['merge'][0]], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    ms_step, ms_count, ms_scan, ms_count_tracked, ms_merge = t.tolist()
    flat_bytes = sketches[0].flat_device_buffer()[1]
    out = None
    if rank == 0:
        peak, peak_src = load_peaks()
        nk_all = n_total * (READ_LEN - K + 1)
        lfrac = READ_LEN / float(READ_LEN - K + 1)
        kern_count = {k: round(v[0], 3) for k, v in prof_count.items() if v[1]}
        kern_scan = {k: round(v[0], 3) for k, v in prof_scan.items() if v[1]}
        upd_ms = prof_count['increment'][0] or 1e-9      # rank 0, all three samples
        nov_ms = prof_scan['novel'][0] or 1e-9
        out = {
            'workload': 'C3: synthetic 100 Mbp trio at 30x = {} reads x {} bp per sample (drawn on the device), k={}, 3 x {:.0f} GB '
                        '8-bit Counttable with {} tables (HBM-resident); reads sharded over {} rank(s) (strong scaling), partial '
                        'sketches merged by the one-pass p2p all-reduce (each sample\'s merge under the next sample\'s count), '
                        'novel scan shard-local'.format(
                            n_total, READ_LEN, K, args.c3_memory / 1e9, N_TABLES, world),
            'scaling': 'strong', 'steps': args.c3_steps, 'kmers_per_step': 4 * nk_all,
            'value': 4 * nk_all / (ms_step / 1e3), 'unit': 'k-mers/s', 'ms_per_step': ms_step,
            'count': {'ms': ms_count, 'kmers_per_s': 3 * nk_all / (ms_count / 1e3),
                      'frac_of_hbm_model': 3 * nk_all / (ms_count / 1e3) * (64.0 * N_TABLES + lfrac) / (peak * 1e9),
                      'ms_with_exact_n_unique': ms_count_tracked,
                      'kmers_per_s_with_exact_n_unique': 3 * nk_all / (ms_count_tracked / 1e3),
                      'kernel_ms_rank0': kern_count},
            'novel': {'ms': ms_scan, 'kmers_per_s': nk_all / (ms_scan / 1e3),
                      'frac_of_hbm_model': nk_all / (ms_scan / 1e3) * (32.0 * 3 * N_TABLES + lfrac) / (peak * 1e9),
                      'kernel_ms_rank0': kern_scan},
            'roofline_update': {
                'kernel': 'kv_hash_kernel<scatter> + kv_tile_apply_kernel<8> (tiled update path, K3c)', 'bound': 'hbm',
                'achieved': 64.0 * N_TABLES * 3 * nk_sample / ((upd_ms + prof_count['hash'][0] + prof_count['partition'][0]) / 1e3) / 1e9,
                'peak': peak, 'unit': 'GB/s', 'peak_source': peak_src,
                'apply_only_GBps': 64.0 * N_TABLES * 3 * nk_sample / (upd_ms / 1e3) / 1e9,
                'note': 'algorithmic 64 B x 4 tables per k-mer (SURVEY 8d: one sector read + written per table touch) over '
                        'ALL update kernels of rank 0: the hash kernel, which also files every update as a 2-byte offset in '
                        'its region slab, and the apply kernel, which streams each region through shared memory once; '
                        '`apply_only_GBps` is the same bytes over the apply kernel alone (it beats the sector model because '
                        'a region is read and written once however many updates it takes)'},
            'roofline_novel': {
                'kernel': 'kv_novel_kernel', 'bound': 'hbm',
                'achieved': (32.0 * 3 * N_TABLES + lfrac) * nk_sample / (nov_ms / 1e3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                'peak_source': peak_src},
            'merge': {'ms_per_step': ms_merge, 'share_of_count': ms_merge / ms_count if ms_count else None,
                      'overlapped_with_counting': bool(world > 1 and not args.no_overlap),
                      'note': 'kernel time of the merge class on rank 0 (barriers + all-reduce kernels); with overlap each '
                              "sample's merge runs on the merge lane under the next sample's count, so only part of it is "
                              'exposed in count.ms',
                      'bytes_in_plus_out_per_rank': 2 * 3 * flat_bytes * (world - 1) / world,
                      'nvlink_GBps_per_rank': (2 * 3 * flat_bytes * (world - 1) / world) / (ms_merge / 1e3) / 1e9 if ms_merge else None},
            'properties_at_full_size': props,
            'parity_vs_oracle': parity,
        }
        for key in ('roofline_update', 'roofline_novel'):
            out[key]['frac'] = out[key]['achieved'] / peak
    barrier()
    return out, sketches


# ----------------------------------------------------------------------------- config 4 (shape)

def run_c4(args, rank, world, barrier, phase):
    """BASELINE config 4's SHAPE at a size one run can afford: 4-bit SmallCounttable sketches that do not
    fit one GPU, spread over the HBM of all ranks (spanning sketches), reads drawn on the device and sharded,
    updates exchanged inside the tiled apply kernel (2 B per update over NVLink), novel scan shard-local with
    plain loads that cross NVLink for remote pages.  Full-size checks are properties: the same reads counted
    under a different read-to-rank assignment give byte-identical sketches and the same number of hits."""
    import torch
    from kevlar_b200 import _lib, khmer, multigpu, simtrio
    td = torch.distributed
    device = _lib.current_device()
    dev = torch.device('cuda', device)
    n_total = args.c4_reads
    buckets = args.c4_gb_per_gpu * 1e9 * world * 2 / N_TABLES          # 4-bit: two buckets per byte
    spans = [multigpu.SpanningSketch(khmer.SmallCounttable, K, buckets, N_TABLES, chunk_positions=args.c4_chunk) for _ in range(3)]
    sketches = [sp.sketch for sp in spans]
    phase('c4: 3 spanning sketches of {:.0f} GB allocated ({:.0f} GB per GPU each)'.format(
        args.c4_gb_per_gpu * world, args.c4_gb_per_gpu))
    stream = torch.cuda.ExternalStream(_lib.stream_ptr(device), device=dev)
    state = {}

    def draw(shift):
        r = (rank + shift) % world
        return simtrio.device_trio(args.c3_genome, n_total, r, world, read_len=READ_LEN)

    def count(trio):
        for sp in spans:
            sp.clear()
        for sp, (b, o) in zip(spans, trio):
            sp.consume_batch(b.data_ptr(), (o.data_ptr(), o.numel() - 1), where=khmer.MEM_DEVICE, n_positions=b.numel())

    def scan(trio):
        b, o = trio[0]
        state['hits'] = khmer.novel_batch(sketches[:1], sketches[1:], b.data_ptr(), (o.data_ptr(), o.numel() - 1, b.numel()),
                                          CASE_MIN, CTRL_MAX, where=khmer.MEM_DEVICE)[0]

    def timed(fn):
        barrier()
        torch.cuda.synchronize()
        _lib.sync(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        _lib.sync(device)
        torch.cuda.synchronize()
        barrier()
        return e0.elapsed_time(e1)

    def local_checksums():
        out = []
        for sk in sketches:
            out.append(sketch_checksum(torch, sk, local_only=True))
        return out

    trio = draw(0)
    count(trio)
    scan(trio)          # warm-up
    _lib.profile(1)
    ms_count = timed(lambda: count(trio))
    prof_count = _lib.profile(1)
    ms_scan = timed(lambda: scan(trio))
    prof_scan = _lib.profile(0)
    sums_a = local_checksums()
    hits_a = torch.tensor([len(state['hits'])], dtype=torch.int64, device=dev)
    occ_a = [sp.n_occupied() for sp in spans]
    del trio
    trio = draw(1)      # every rank counts the NEXT rank's slice: same reads, different owners
    count(trio)
    scan(trio)
    sums_b = local_checksums()
    hits_b = torch.tensor([len(state['hits'])], dtype=torch.int64, device=dev)
    if world > 1:
        td.all_reduce(hits_a)
        td.all_reduce(hits_b)
    t = torch.tensor([ms_count, ms_scan], dtype=torch.float64, device=dev)
    differs = torch.tensor([0 if sums_a == sums_b else 1], dtype=torch.int64, device=dev)   # every rank checks its own pieces
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
        td.all_reduce(differs)
    same_bytes = int(differs.item()) == 0
    ms_count, ms_scan = t.tolist()
    nk = n_total * (READ_LEN - K + 1)
    out = None
    if rank == 0:
        out = {
            'workload': 'C4 shape: 4-bit SmallCounttable, {} tables, 3 sketches of {:.0f} GB each spanning {} GPU(s) ({:.0f} GB per '
                        'GPU and sketch), {} reads x {} bp per sample drawn on the device and sharded, k={}'.format(
                            N_TABLES, args.c4_gb_per_gpu * world, world, args.c4_gb_per_gpu, n_total, READ_LEN, K),
            'count': {'ms': ms_count, 'kmers_per_s': 3 * nk / (ms_count / 1e3), 'kernel_ms_rank0': {k: round(v[0], 3) for k, v in prof_count.items() if v[1]}},
            'novel': {'ms': ms_scan, 'kmers_per_s': nk / (ms_scan / 1e3), 'kernel_ms_rank0': {k: round(v[0], 3) for k, v in prof_scan.items() if v[1]}},
            'value': 4 * nk / ((ms_count + ms_scan) / 1e3), 'unit': 'k-mers/s',
            'nvlink_bytes_per_kmer': {
                'count': 'updates cross NVLink as 16-bit offsets pulled by the owner of the region: {} tables x 2 B x {}/{} = {:.1f} B per '
                         'k-mer (the hash all-gather of the first design moved 8 B x {} per position)'.format(
                             N_TABLES, world - 1, world, N_TABLES * 2.0 * (world - 1) / max(world, 1), world - 1),
                'novel': 'counter loads for remote pages: one 32 B sector per table touch x {}/{} of the touches'.format(world - 1, world)},
            'properties_at_full_size': {
                'same_sketch_bytes_under_a_rotated_read_assignment': same_bytes,
                'same_hit_count_under_a_rotated_read_assignment': int(hits_a.item()) == int(hits_b.item()),
                'novel_hits_all_ranks': int(hits_a.item()), 'n_occupied': occ_a},
        }
    del sketches[:]
    state.clear()
    for sp in spans:
        sp.close()
    barrier()
    return out


# ----------------------------------------------------------------------------- config 5 (banded chain)

def run_c5(args, rank, world, trio0, barrier, phase):
    """BASELINE config 5's chain through the FILE-based CLI path: the C2 trio of rank 0 written as FASTQ,
    `kevlar novel --num-bands 8 --band b` for every band (band b on rank (b-1) mod world, kevlar_b200.bands),
    `unband`, `filter` recount on rank 0.  Reported as wall-clock k-mer events per second: every band hashes all
    reads of the three samples (3 counts + 1 scan per band), as the reference's banding does."""
    import tempfile
    import torch
    import kevlar_b200
    from kevlar_b200 import bands
    td = torch.distributed
    shared = [tempfile.mkdtemp(prefix='kv_c5_') if rank == 0 else None]
    if world > 1:
        td.broadcast_object_list(shared, src=0)
    tmp = shared[0]
    names = ['proband', 'mother', 'father']
    if rank == 0:
        for name, (b, o) in zip(names, trio0):
            reads = b.reshape(-1, READ_LEN)
            qual = b'I' * READ_LEN
            with open(os.path.join(tmp, name + '.fq'), 'wb') as fh:
                fh.write(b''.join(b'@r%d\n%s\n+\n%s\n' % (i, reads[i].tobytes(), qual) for i in range(len(reads))))
    barrier()
    files = [os.path.join(tmp, n + '.fq') for n in names]
    ns = bands.parser().parse_args(['--case', files[0], '--control', files[1], '--control', files[2], '-k', str(K), '--memory',
                                    str(MEMORY / 8), '--case-min', str(CASE_MIN), '--ctrl-max', str(CTRL_MAX), '--num-bands', '8',
                                    '--filter-memory', '1M', '--out-prefix', os.path.join(tmp, 'c5')])
    saved = kevlar_b200.logstream
    kevlar_b200.logstream = open(os.devnull, 'w')
    barrier()
    t0 = time.perf_counter()
    result = bands.run(ns, rank, world, barrier=barrier)
    secs = time.perf_counter() - t0
    kevlar_b200.logstream = saved
    out = None
    if rank == 0:
        nk = kmers_per_step(trio0)
        n_reads = sum(1 for line in open(result['filter']) if line.startswith('@r'))
        out = {'workload': 'C5: C2 trio as FASTQ files, novel --num-bands 8 (sketch memory/8 per band, band b on rank (b-1) mod {}), '
                           'unband, filter recount; file-based CLI path'.format(world),
               'seconds': secs, 'kmer_events': 8 * nk, 'value': 8 * nk / secs, 'unit': 'k-mers/s',
               'reads_after_filter': n_reads}
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    barrier()
    return out


def run_files(args, trio0, phase):
    """The FILE-based path at N = 1 (SURVEY 8d "to last sketch byte on host"): the C2 trio written as three FASTQ files,
    `kevlar count` per sample (native reader -> GPU -> 64 MB sketch file), then `kevlar novel` from the case file and
    the three sketch files to an augmented FASTQ.  In-process calls of the CLI mains; wall clock."""
    import shutil
    import tempfile
    import kevlar_b200
    from kevlar_b200 import cli, count, fastx, novel
    tmp = tempfile.mkdtemp(prefix='kv_files_')
    names = ['proband', 'mother', 'father']
    try:
        for name, (b, o) in zip(names, trio0):
            reads = b.reshape(-1, READ_LEN)
            qual = b'I' * READ_LEN
            with open(os.path.join(tmp, name + '.fq'), 'wb') as fh:
                fh.write(b''.join(b'@r%d\n%s\n+\n%s\n' % (i, reads[i].tobytes(), qual) for i in range(len(reads))))
        files = [os.path.join(tmp, n + '.fq') for n in names]
        text_bytes = sum(os.path.getsize(f) for f in files)
        saved = kevlar_b200.logstream
        kevlar_b200.logstream = open(os.devnull, 'w')
        best_parse = None
        for _ in range(3):
            t = time.perf_counter()
            n = sum(len(batch) for batch in fastx.NativeFastxReader(files[0]).batches(64 << 20, prefetch=False))
            dt = time.perf_counter() - t
            best_parse = dt if best_parse is None else min(best_parse, dt)
        assert n == len(trio0[0][1]) - 1

        def once():
            t0 = time.perf_counter()
            for name, f in zip(names, files):
                a = cli.parser().parse_args(['count', '--memory', str(int(MEMORY)), '-k', str(K), os.path.join(tmp, name + '.ct'), f])
                count.main(a)
            t1 = time.perf_counter()
            a = cli.parser().parse_args(['novel', '--case', files[0], '--case-counts', os.path.join(tmp, 'proband.ct'),
                                         '--control-counts', os.path.join(tmp, 'mother.ct'), os.path.join(tmp, 'father.ct'),
                                         '-k', str(K), '--case-min', str(CASE_MIN), '--ctrl-max', str(CTRL_MAX),
                                         '-o', os.path.join(tmp, 'novel.augfastq')])
            novel.main(a)
            return t1 - t0, time.perf_counter() - t1
        once()
        t_count, t_novel = min((once() for _ in range(2)), key=sum)
        kevlar_b200.logstream = saved
        nk = kmers_per_step(trio0)
        out = {'workload': 'C2 trio as three FASTQ files ({:.0f} MB of text): kevlar count x3 (parse -> GPU -> 64 MB sketch file each), '
                           'kevlar novel (3 sketch files + case FASTQ -> augmented FASTQ); in-process CLI mains, files in {}'.format(
                               text_bytes / 1e6, tempfile.gettempdir()),
               'value': nk / (t_count + t_novel), 'unit': 'k-mers/s', 'count_x3_seconds': t_count, 'novel_seconds': t_novel,
               'reader': {'threads_env': os.environ.get('KV_READER_THREADS', 'default: min(8, cores / local ranks)'),
                          'cores': os.cpu_count(), 'seconds_per_file': best_parse,
                          'text_GBps': os.path.getsize(files[0]) / best_parse / 1e9,
                          'kmers_per_s': (nk / 4) / best_parse}}
        phase('file-based path done')
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_ours(args):
    t_start = time.perf_counter()
    import torch
    from kevlar_b200 import _lib, multigpu, simtrio
    rank, world = multigpu.init_from_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    if _lib.device_count() < 1:
        raise SystemExit('bench.py: no CUDA device; the GPU arm has no CPU fallback')
    torch.cuda.set_device(_lib.current_device())
    live = []   # every sketch a merge may have exported: released collectively in the teardown

    def barrier():
        if world > 1:
            torch.distributed.barrier()

    def phase(what):   # progress on stderr (rank 0): where the wall-clock of a run goes
        if rank == 0:
            print('[bench] +{:6.1f}s {}'.format(time.perf_counter() - t_start, what), file=sys.stderr, flush=True)

    try:
        line = measure(args, rank, world, barrier, phase, live)
        if rank == 0:
            print(json.dumps(line), flush=True)
    finally:
        # collective teardown on EVERY rank: unmap peers, meet, free, leave the group
        if world > 1:
            multigpu.shutdown(live)
        live.clear()


def measure(args, rank, world, barrier, phase, live):
    import torch
    from kevlar_b200 import _lib, multigpu, simtrio
    phase('rendezvous done ({} rank{})'.format(world, 's' if world > 1 else ''))
    trio = simtrio.simulate_trio(1000000, reads_per_sample=args.reads_per_sample, seed_offset=1000 * rank)
    nk_rank = kmers_per_step(trio)
    phase('synthetic trio generated')
    runner = GpuTrio(args, trio, world)
    live.extend(runner.plain_sketches())
    phase('sketches allocated, inputs staged')

    for _ in range(max(3, args.warmup)):
        runner.step(True)
    for _ in range(2):
        runner.step(False)
    phase('warm-up done')

    # headline: per-launch profiling OFF; clocks sampled during the timed region
    sampler = ClockSampler(_lib.current_device())
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ms_value = runner.timed(args.steps, True, barrier)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    hits_value = runner.last_hits.copy()
    ms_e2e = runner.timed(args.steps, False, barrier)
    hits_e2e = runner.last_hits.copy()
    # separate pass for the per-kernel-class split (event pair around every launch)
    prof_steps = min(args.steps, 5)
    _lib.profile(1)
    ms_profiled = runner.timed(prof_steps, True, barrier)
    prof = _lib.profile(0)
    phase('timed regions done')
    assert len(hits_value) == len(hits_e2e) and (hits_value['offset'] == hits_e2e['offset']).all()

    # whole-job numbers: max time over ranks, k-mers summed over ranks
    t = torch.tensor([ms_value, ms_e2e], dtype=torch.float64, device='cuda')
    n = torch.tensor([float(nk_rank)], dtype=torch.float64, device='cuda')
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(n, op=torch.distributed.ReduceOp.SUM)
    ms_value, ms_e2e = t.tolist()
    nk = n.item()

    parity_n = None
    if world > 1 and not runner.sharded and not args.no_cpu_baseline:
        parity_n = multi_rank_parity(args, runner, rank, world, torch, multigpu, phase)

    c3 = None
    if not args.no_c3 and not runner.sharded:
        c3, c3_sketches = run_c3(args, rank, world, barrier, phase)
        live.extend(c3_sketches)
        if world > 1:   # free 12 GB per rank before anything else is allocated
            multigpu.release_p2p(c3_sketches)
        for sk in c3_sketches:
            live.remove(sk)
        del c3_sketches

    c4 = None
    if args.c4 and not runner.sharded:
        c4 = run_c4(args, rank, world, barrier, phase)
    c5 = None
    if args.c5 or (world > 1 and not args.no_c5):
        trio0 = trio if rank == 0 else None
        c5 = run_c5(args, rank, world, trio0, barrier, phase)
        phase('c5: banded chain done')
    files = None
    if world == 1 and not args.no_files and not runner.sharded and not getattr(runner, 'span', False):
        files = run_files(args, trio, phase)
    if getattr(runner, 'span', False):
        runner.sketches = []
        for sp in runner.spans:
            sp.close()
    if world > 1 and args.merge in ('p2p', 'span'):
        multigpu.peer_sync_status()   # raises if a device-side barrier ever timed out
    if rank != 0:
        return None

    value = nk * args.steps / (ms_value / 1e3)
    e2e = nk * args.steps / (ms_e2e / 1e3)
    h2d = sum(b.nbytes + o.nbytes for b, o in trio) + trio[0][0].nbytes + trio[0][1].nbytes
    d2h = len(hits_e2e) * 24 + len(trio[0][1]) * 4 + 64

    # per-kernel-class breakdown (profiled pass), device time from CUDA events
    kmers_count = (nk_rank - int(np.maximum(np.diff(trio[0][1].astype(np.int64)) - K + 1, 0).sum())) // 3
    kmers_scan = nk_rank - 3 * kmers_count
    lfrac = READ_LEN / float(READ_LEN - K + 1)
    alg_bytes = {   # algorithmic bytes per STEP and kernel class (SURVEY.md 8d per-k-mer figure x k-mers per step)
        'increment': 64.0 * N_TABLES * kmers_count * 3,
        'hash': lfrac * kmers_count * 3,
        'novel': (32.0 * 3 * N_TABLES + lfrac) * kmers_scan,
    }
    peak, peak_src = load_peaks()
    total_kernel_ms = sum(ms for ms, _ in prof.values()) or 1.0
    kernels = {}
    for name, (ms, count) in prof.items():
        if not count:
            continue
        entry = {'launches_per_step': count / prof_steps, 'ms_per_launch': ms / count,
                 'share_of_kernel_time': ms / total_kernel_ms}
        entry['ms_per_step'] = ms / prof_steps
        if name in alg_bytes:
            entry['algorithmic_GBps'] = alg_bytes[name] / (ms / prof_steps / 1e3) / 1e9
        kernels[name] = entry
    dominant_any = max(kernels, key=lambda k_: kernels[k_]['share_of_kernel_time'])
    dominant = max((k for k in kernels if k in alg_bytes), key=lambda k_: kernels[k_]['share_of_kernel_time'])
    kname = {'increment': 'kv_increment_kernel<{}>'.format(args.counter_size), 'hash': 'kv_hash_kernel',
             'novel': 'kv_novel_kernel'}[dominant]
    achieved = kernels[dominant]['algorithmic_GBps']
    roofline = {'kernel': kname, 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': load_traffic(kname), 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': alg_bytes[dominant] / kernels[dominant]['launches_per_step'],
                'launches_per_step': kernels[dominant]['launches_per_step'],
                'avg_launch_ms': kernels[dominant]['ms_per_launch'],
                'largest_kernel_class': dominant_any}
    if dominant_any != dominant:
        roofline['note_class'] = ('the largest class by device time is `{}` ({:.0%}), bookkeeping with no algorithmic bytes in '
                                  'SURVEY 8(d); the roofline is quoted for the largest kernel that has them').format(
                                      dominant_any, kernels[dominant_any]['share_of_kernel_time'])
    if roofline['frac'] > 1.0:
        roofline['note'] = ('frac > 1: the 64 MB sketch is L2-resident, so most of the algorithmic sector traffic (64 B per '
                            'table touch) never reaches HBM -- `traffic` is the measured DRAM bytes per launch; the bound '
                            'that applies is roofline_l2 (L2 atomic throughput); the HBM-resident regime is in `c3`')
    # L2-resident case: update rate against the microbenchmarked L2 atomic peak for a 64 MB table
    atomic = load_atomic_peak(64)
    roofline_l2 = None
    if prof['increment'][1]:
        updates_per_s = N_TABLES * kmers_count * 3 / (prof['increment'][0] / prof_steps / 1e3)
        roofline_l2 = {'kernel': 'kv_increment_kernel<8>', 'bound': 'l2_atomic', 'achieved': updates_per_s / 1e9,
                       'unit': 'G updates/s', 'peak': atomic.get('atom_add'), 'peak_red_add': atomic.get('red_add'),
                       'frac': updates_per_s / 1e9 / atomic['atom_add'] if atomic.get('atom_add') else None,
                       'peak_source': 'tools/atomic_microbench.cu on this pool (profiles/r01_atomic_microbench.csv): random '
                                      'ATOM.ADD / RED.ADD over a 64 MB table'}
    # the count path (hash + unique + increment per sample) and the novel path against 8d's figures
    count_ms = sum(prof[c][0] for c in ('hash', 'unique', 'increment', 'fixup', 'partition')) / (3.0 * prof_steps)
    novel_ms = prof['novel'][0] / prof_steps
    paths = {
        'count': {'ms_per_sample': count_ms, 'kmers_per_s': kmers_count / (count_ms / 1e3),
                  'algorithmic_GBps': (64.0 * N_TABLES + lfrac) * kmers_count / (count_ms / 1e3) / 1e9},
        'novel': {'ms': novel_ms, 'kmers_per_s': kmers_scan / (novel_ms / 1e3),
                  'algorithmic_GBps': alg_bytes['novel'] / (novel_ms / 1e3) / 1e9},
    }
    for p in paths.values():
        p['frac_of_peak'] = p['algorithmic_GBps'] / peak

    line = {
        'metric': 'kmers_per_sec_count_plus_novel', 'value': value, 'unit': 'k-mers/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms_value / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': config_of(args, world),
        'kmers_per_step': nk, 'merge': args.merge if world > 1 else None,
        'e2e': {'value': e2e, 'unit': 'k-mers/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': roofline,
        'roofline_l2': roofline_l2,
        'kernels': kernels,
        'kernels_note': 'per-class split from a separate pass of {} steps with an event pair around every launch ({:.3f} ms/step); '
                        'the headline is timed with that switched off'.format(prof_steps, ms_profiled / prof_steps),
        'paths': paths,
        'novel_hits_per_step': int(len(hits_value)),
    }
    if c3 is not None:
        line['c3'] = c3
    if c4 is not None:
        line['c4'] = c4
    if c5 is not None:
        line['c5'] = c5
    if files is not None:
        line['e2e_file'] = files

    if world == 1 and not args.no_cpu_baseline:
        from oracle import khmer_oracle as ko
        cores = os.cpu_count() or 1
        secs, ohits, counted, osk = cpu_step(ko, trio, cores, sketch_shape(args))
        line['cpu_baseline'] = {
            'value': nk_rank / secs, 'unit': 'k-mers/s', 'cores': cores, 'kind': 'port',
            'sample': 'one full step (3 x {} reads): oracle count in {} pthreads + C novel scan in {} threads, {:.1f} s'
                      .format(args.reads_per_sample, cores, cores, secs)}
        # parity inside the bench: same hits, same sketch bytes as the oracle on this very workload
        same = len(ohits) == len(hits_value) and bool((ohits['read'] == hits_value['read']).all()) and \
            bool((ohits['offset'] == hits_value['offset']).all()) and \
            bool((ohits['abund'][:, :3] == hits_value['abund'][:, :3]).all())
        for g, c in zip(runner.sketches, osk):
            for tb in range(N_TABLES):
                same = same and g.table_bytes(tb) == c.table_bytes(tb)
        line['parity_vs_oracle'] = 'bit-exact (3 sketches, {} hits)'.format(len(ohits)) if same else 'MISMATCH'
        if not same:
            raise SystemExit('bench.py: GPU results differ from the oracle')
        # the reference's REAL novel path beside the multi-threaded C one -- a single-threaded Python loop
        py_rate, py_reads = reference_python_novel(ko, trio, sks=osk)
        line['cpu_baseline']['reference_python_novel_loop'] = {
            'value': py_rate, 'unit': 'k-mers/s', 'cores': 1,
            'sample': 'kevlar/novel.py:123-169 as the reference runs it (one get() per sample and k-mer from Python) over the '
                      'oracle sketches, first {} proband reads'.format(py_reads)}
    if parity_n is not None:
        line['parity_vs_oracle'] = parity_n[1]
        if not parity_n[0]:
            print(json.dumps(line), flush=True)
            raise SystemExit('bench.py: merged GPU results differ from the oracle')

    if world == 1 and not args.no_variants:
        # the same step with the other hasher / without n_unique tracking, for context
        variants = {}
        for label, hasher, no_unique, bits in (('countgraph_twobit', 'twobit', False, 8),
                                               ('counttable_no_unique', 'murmur', True, 8),
                                               ('smallcounttable_4bit', 'murmur', False, 4),
                                               ('nodetable_1bit', 'murmur', False, 1)):
            if hasher == args.hasher and no_unique == args.no_unique and bits == args.counter_size:
                continue
            a2 = argparse.Namespace(**vars(args))
            a2.hasher, a2.no_unique, a2.counter_size = hasher, no_unique, bits
            r2 = GpuTrio(a2, trio, world)
            for _ in range(3):
                r2.step(True)
            ms2 = r2.timed(args.steps, True, barrier)
            variants[label] = {'value': nk * args.steps / (ms2 / 1e3), 'ms_per_step': ms2 / args.steps}
            del r2
        line['variants'] = variants
    return line


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
