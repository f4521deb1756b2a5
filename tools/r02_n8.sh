#!/bin/bash
# 8-GPU validation + scaling session (one box): multi-rank parity tests, then the bench at N=8 and N=4,
# then the config-4-shaped spanning-sketch run
python -m pytest tests/test_multigpu.py -m gpu -x -q -k "4 or 8" 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02o_bench_n8.json 2> gpurun_out/r02o_bench_n8.err; echo "n8 rc=$?"
grep "^\[bench\]" gpurun_out/r02o_bench_n8.err | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 4 --steps 20 --warmup 5 --c3-no-parity > gpurun_out/r02o_bench_n4.json 2> gpurun_out/r02o_bench_n4.err; echo "n4 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 8 --steps 10 --warmup 3 --merge span --no-c3 --no-c5 --c4 --c4-gb-per-gpu 16 > gpurun_out/r02o_span_n8.json 2> gpurun_out/r02o_span_n8.err; echo "span rc=$?"
grep "^\[bench\]\|rror" gpurun_out/r02o_span_n8.err | tail -6
python - <<'PY'
import json
for f in ('r02o_bench_n8', 'r02o_bench_n4', 'r02o_span_n8'):
    try:
        for l in open('gpurun_out/%s.json' % f):
            if l.startswith('{'):
                d = json.loads(l)
                print(f, 'value %.3g ms %.2f e2e %.3g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d.get('parity_vs_oracle'))
                c = d.get('c3')
                if c: print('  c3 value %.3g count %.3g novel %.3g merge' % (c['value'], c['count']['kmers_per_s'], c['novel']['kmers_per_s']), c['merge'], c['properties_at_full_size'], c['parity_vs_oracle'])
                if d.get('c4'): print('  c4', json.dumps(d['c4']['count']), d['c4']['novel']['kmers_per_s'], d['c4']['properties_at_full_size'])
                if d.get('c5'): print('  c5', d['c5'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
