#!/bin/bash
# 8-GPU session (one box): multi-rank parity tests at 8 ranks, the bench at N=8 (C2 weak scaling + C3 strong scaling +
# C5 chain), then the config-4-shaped spanning-sketch run
(time python -m pytest "tests/test_multigpu.py::test_merge_and_novel_match_oracle[8]" "tests/test_multigpu.py::test_spanning_sketches_shared_memory_apply[8]" -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02w_pytest_n8.log 2>&1; cat gpurun_out/r02w_pytest_n8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02w_bench_n8.json 2> gpurun_out/r02w_bench_n8.err; echo "n8 rc=$?"
grep "^\[bench\]" gpurun_out/r02w_bench_n8.err | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 8 --steps 10 --warmup 3 --merge span --no-c3 --no-c5 --no-variants --no-cpu-baseline --c4 --c4-gb-per-gpu 16 > gpurun_out/r02w_span_n8.json 2> gpurun_out/r02w_span_n8.err; echo "span rc=$?"
grep "^\[bench\]\|rror" gpurun_out/r02w_span_n8.err | tail -6
python - <<'PY'
import json
for f in ('r02w_bench_n8', 'r02w_span_n8'):
    try:
        for l in open('gpurun_out/%s.json' % f):
            if l.startswith('{'):
                d = json.loads(l)
                print(f, 'value %.4g ms %.2f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d.get('parity_vs_oracle'))
                print('  ', {k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
                c = d.get('c3')
                if c: print('  c3 value %.4g step %.1f count %.1f novel %.1f' % (c['value'], c['ms_per_step'], c['count']['ms'], c['novel']['ms']), c['count']['kernel_ms_rank0'], c['merge'], c['properties_at_full_size'], c['parity_vs_oracle'])
                if d.get('c4'): print('  c4', json.dumps(d['c4']['count']), d['c4']['novel']['kmers_per_s'], d['c4']['properties_at_full_size'])
                if d.get('c5'): print('  c5', d['c5'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
