#!/bin/bash
out=gpurun_out/r02e.jsonl
: > $out
run() { label="$1"; shift; env "$@" timeout 300 python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 2 --label "$label" >> $out 2>> gpurun_out/r02e.err; }
run "default(1cta)"
run "scatter 5 ctas" KV_LIB_PATH=$PWD/gpurun_exp_scatter5.so
run "scatter 6 ctas" KV_LIB_PATH=$PWD/gpurun_exp_scatter6.so
env timeout 300 python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 2 --track --label "tracked v2" >> $out 2>> gpurun_out/r02e.err
cat $out | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['label'], d['count_x3']['G_kmers_per_s'], d['count_x3']['kernel_ms'], d['novel']['ms'])
"
