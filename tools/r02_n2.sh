#!/bin/bash
# 2-GPU validation of the round-2 multi-GPU changes: in-flight peer loads in the merge kernel, n_unique shares from the
# hashes left on the device, merges on the (high-priority) merge lane; then the bench at N=2, and the lane width
python -m pytest "tests/test_multigpu.py::test_merge_and_novel_match_oracle[2]" -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02v_bench_n2.json 2> gpurun_out/r02v_bench_n2.err; echo "n2 rc=$?"
grep "^\[bench\]\|rror" gpurun_out/r02v_bench_n2.err | tail -3
for c in 1 4 8; do
KV_MERGE_LANE_CTAS=$c python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$c bench.py --gpus 2 --steps 10 --warmup 3 --c3-no-parity --no-c5 --no-variants --no-cpu-baseline > gpurun_out/r02v_bench_n2_lane$c.json 2> gpurun_out/r02v_bench_n2_lane$c.err; echo "lane $c rc=$?"
done
python - <<'PY'
import json
for f in ('r02v_bench_n2', 'r02v_bench_n2_lane1', 'r02v_bench_n2_lane4', 'r02v_bench_n2_lane8'):
    try:
        for l in open('gpurun_out/%s.json' % f):
            if l.startswith('{'):
                d = json.loads(l)
                print(f, 'value %.4g ms %.2f e2e %.4g (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), d.get('parity_vs_oracle'))
                print('  ', {k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
                c = d.get('c3')
                if c: print('  c3 value %.4g step %.1f count %.1f novel %.1f' % (c['value'], c['ms_per_step'], c['count']['ms'], c['novel']['ms']), c['count']['kernel_ms_rank0'], c['properties_at_full_size'], (c['parity_vs_oracle'] or {}).get('result'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
