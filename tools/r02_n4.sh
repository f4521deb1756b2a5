#!/bin/bash
# 4-GPU check of the merge kernel's 3-peer instantiation and the merge lane: parity test + bench
(time python -m pytest "tests/test_multigpu.py::test_merge_and_novel_match_oracle[4]" -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r02x_pytest_n4.log 2>&1; cat gpurun_out/r02x_pytest_n4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02x_bench_n4.json 2> gpurun_out/r02x_bench_n4.err; echo "n4 rc=$?"
grep "^\[bench\]\|rror" gpurun_out/r02x_bench_n4.err | tail -3
python - <<'PY'
import json
for l in open('gpurun_out/r02x_bench_n4.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print('value %.4g ms %.2f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d.get('parity_vs_oracle'))
        print('  ', {k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
        c = d.get('c3')
        if c: print('  c3 value %.4g step %.1f count %.1f novel %.1f' % (c['value'], c['ms_per_step'], c['count']['ms'], c['novel']['ms']), c['count']['kernel_ms_rank0'], c['properties_at_full_size'], c['parity_vs_oracle'])
PY
