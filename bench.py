#!/usr/bin/env python
"""Headline benchmark: k-mers/s for `kevlar count` x3 + `kevlar novel` on a synthetic trio.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one trio batch: zero three sketches, count the
proband / mother / father reads into them (hash + exact-n_unique bookkeeping + saturating
increments), [N>1: merge the per-GPU partial sketches], then scan the proband reads against
all three sketches (kevlar novel).  k-mers per step = sum over the 3 samples of counted
k-mer positions + scanned k-mer positions.

Workload at N=1 = BASELINE.json configs[1]: 1 Mbp genome, 30x => 300,000 reads x 100 bp per
sample, k=31, 64 MB / 4-table 8-bit sketches.  Under torchrun every rank gets its own
300k-read shard per sample (weak scaling; reads sharded, sketches merged).

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
    ap.add_argument('--merge', default='p2p', choices=['allreduce', 'allgather', 'p2p', 'p2p_host', 'sharded'],
                    help="N>1: how per-GPU work is combined; 'sharded' = plan B, bin-range-sharded sketches")
    ap.add_argument('--reads-per-sample', type=int, default=READS_PER_SAMPLE)
    ap.add_argument('--no-unique', action='store_true', help='skip the exact n_unique_kmers bookkeeping')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-variants', action='store_true')
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


def kmers_per_step(trio):
    total = 0
    for bases, offs in trio:
        lens = np.diff(offs.astype(np.int64))
        total += int(np.maximum(lens - K + 1, 0).sum())
    lens = np.diff(trio[0][1].astype(np.int64))
    return total + int(np.maximum(lens - K + 1, 0).sum())


# ----------------------------------------------------------------------------- clocks

class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
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
    line = {
        'impl': 'reference', 'metric': 'kmers_per_sec_count_plus_novel', 'value': value, 'unit': 'k-mers/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'arm': 'CPU oracle port of the khmer path (khmer is not vendored)'},
        'cpu_baseline': {'value': value, 'unit': 'k-mers/s', 'cores': cores, 'kind': 'port', 'sample': sample},
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
        if self.sharded:   # plan B: every rank holds 1/world of every table; reads stay sharded
            self.sketches = [multigpu.ShardedSketch(cls, K, buckets, N_TABLES) for _ in range(3)]
        else:
            self.sketches = [cls(K, buckets, N_TABLES) for _ in range(3)]
        if args.no_unique and not self.sharded:
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

    def step_sharded(self):
        # host buffers in both arms: hash locally, all-gather the hash stream, apply to the local bin
        # ranges; novel from MIN-all-reduced partial abundances
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
        khmer = self.khmer
        for sk in self.sketches:
            sk.clear()
        for i, sk in enumerate(self.sketches):
            if resident:
                b, o = self.dev[i]
                sk.consume_batch(b.data_ptr(), (o.data_ptr(), o.numel() - 1), where=khmer.MEM_DEVICE, wait=False)
            else:
                b, o = self.pinned[i]
                sk.consume_batch(b.numpy(), o.numpy().view(np.uint64), wait=False)
        if self.world > 1:
            if self.args.merge != 'p2p':   # NCCL runs on torch's stream; 'p2p' stays on the library's own stream
                self.lib.sync(self.device)
            self.multigpu.merge_sketches(self.sketches, how=self.args.merge)
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

    def barrier():
        if world > 1:
            torch.distributed.barrier()

    def phase(what):   # progress on stderr (rank 0): where the wall-clock of a run goes
        if rank == 0:
            print('[bench] +{:6.1f}s {}'.format(time.perf_counter() - t_start, what), file=sys.stderr, flush=True)

    phase('rendezvous done ({} rank{})'.format(world, 's' if world > 1 else ''))
    trio = simtrio.simulate_trio(1000000, reads_per_sample=args.reads_per_sample, seed_offset=1000 * rank)
    nk_rank = kmers_per_step(trio)
    phase('synthetic trio generated')
    runner = GpuTrio(args, trio, world)
    phase('sketches allocated, inputs staged')

    for _ in range(max(3, args.warmup)):
        runner.step(True)
    for _ in range(2):
        runner.step(False)
    phase('warm-up done')

    sampler = ClockSampler(_lib.current_device())
    if rank == 0:
        sampler.start()
    _lib.profile(1)
    launches0 = _lib.launch_count()
    ms_value = runner.timed(args.steps, True, barrier)
    launches = _lib.launch_count() - launches0
    prof = _lib.profile(0)
    clocks = sampler.stop() if rank == 0 else None
    hits_value = runner.last_hits.copy()
    ms_e2e = runner.timed(args.steps, False, barrier)
    hits_e2e = runner.last_hits.copy()
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
    if rank != 0:
        return

    value = nk * args.steps / (ms_value / 1e3)
    e2e = nk * args.steps / (ms_e2e / 1e3)
    h2d = sum(b.nbytes + o.nbytes for b, o in trio) + trio[0][0].nbytes + trio[0][1].nbytes
    d2h = len(hits_e2e) * 24 + len(trio[0][1]) * 4 + 64

    # per-kernel-class breakdown of the timed (resident) region, device time from CUDA events
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
        entry = {'launches_per_step': count / args.steps, 'ms_per_launch': ms / count,
                 'share_of_kernel_time': ms / total_kernel_ms}
        entry['ms_per_step'] = ms / args.steps
        if name in alg_bytes:
            entry['algorithmic_GBps'] = alg_bytes[name] / (ms / args.steps / 1e3) / 1e9
        kernels[name] = entry
    dominant = max((k for k in kernels if k in alg_bytes), key=lambda k_: kernels[k_]['share_of_kernel_time'])
    kname = {'increment': 'kv_increment_kernel<{}>'.format(args.counter_size), 'hash': 'kv_hash_kernel',
             'novel': 'kv_novel_kernel'}[dominant]
    achieved = kernels[dominant]['algorithmic_GBps']
    roofline = {'kernel': kname, 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': load_traffic(kname), 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': alg_bytes[dominant] / kernels[dominant]['launches_per_step'],
                'launches_per_step': kernels[dominant]['launches_per_step'],
                'avg_launch_ms': kernels[dominant]['ms_per_launch']}
    if roofline['frac'] > 1.0:
        roofline['note'] = ('frac > 1: the 64 MB sketch is L2-resident, so most of the algorithmic sector traffic (64 B per '
                            'table touch) never reaches HBM -- `traffic` is the measured DRAM bytes per launch; the bound '
                            'that applies is roofline_l2 (L2 atomic throughput)')
    # L2-resident case: update rate against the microbenchmarked L2 atomic peak for a 64 MB table
    atomic = load_atomic_peak(64)
    updates_per_s = N_TABLES * kmers_count * 3 / (prof['increment'][0] / args.steps / 1e3)
    roofline_l2 = {'kernel': 'kv_increment_kernel<8>', 'bound': 'l2_atomic', 'achieved': updates_per_s / 1e9,
                   'unit': 'G updates/s', 'peak': atomic.get('atom_add'), 'peak_red_add': atomic.get('red_add'),
                   'frac': updates_per_s / 1e9 / atomic['atom_add'] if atomic.get('atom_add') else None,
                   'peak_source': 'tools/atomic_microbench.cu on this pool (profiles/r01_atomic_microbench.csv): random '
                                  'ATOM.ADD / RED.ADD over a 64 MB table'}
    # the count path (hash + unique + increment per sample) and the novel path against 8d's figures
    count_ms = sum(prof[c][0] for c in ('hash', 'unique', 'increment', 'fixup')) / (3.0 * args.steps)
    novel_ms = prof['novel'][0] / args.steps
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
        'config': {
            'workload': workload_name(args),
            'kmers_per_step': nk, 'reads_sharded': world > 1, 'merge': args.merge if world > 1 else None,
            'exact_n_unique_tracking': not args.no_unique,
            'l2': 'inputs larger than L2: per step 120 MB of reads + 192 MB of sketches + 240 MB of hash scratch '
                  'stream through a 126 MB L2; no explicit flush',
        },
        'e2e': {'value': e2e, 'unit': 'k-mers/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': roofline,
        'roofline_l2': roofline_l2,
        'kernels': kernels,
        'paths': paths,
        'novel_hits_per_step': int(len(hits_value)),
    }

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
    if world > 1 and args.merge == 'p2p':
        runner.multigpu.peer_sync_status()   # raises if a device-side barrier ever timed out
    print(json.dumps(line))
    if world > 1:
        if args.merge == 'p2p':
            runner.multigpu.release_peer_sync()
        torch.distributed.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
