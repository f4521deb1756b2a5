#!/bin/bash
out=gpurun_out/r02r_l2fetch.jsonl
: > $out
for g in 32 64 128; do
  KV_L2_FETCH=$g timeout 300 python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 2 --label "l2fetch $g" >> $out 2>> gpurun_out/r02r.err
  KV_L2_FETCH=$g KV_UPDATE_PATH=direct timeout 300 python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 2 --label "l2fetch $g direct" >> $out 2>> gpurun_out/r02r.err
  KV_L2_FETCH=$g timeout 300 python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 2 --track --label "l2fetch $g tracked" >> $out 2>> gpurun_out/r02r.err
  KV_L2_FETCH=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-c3 --no-cpu-baseline > gpurun_out/r02r_c2_$g.json 2>> gpurun_out/r02r.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02r_l2fetch.jsonl'):
    d = json.loads(l); print(d['config']['label'], 'count', d['count_x3']['G_kmers_per_s'], d['count_x3']['kernel_ms'], 'novel', d['novel']['G_kmers_per_s'], d['novel']['kernel_ms'])
for g in (32, 64, 128):
    d = json.load(open('gpurun_out/r02r_c2_%d.json' % g))
    print('C2 l2fetch', g, d['value'], d['ms_per_step'], {k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()}, d['variants'])
PY
