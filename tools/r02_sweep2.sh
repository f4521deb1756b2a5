#!/bin/bash
out=gpurun_out/r02c_layout.jsonl
: > $out
run() { label="$1"; shift; env "$@" KV_UPDATE_PATH=tile timeout 300 python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 2 --label "$label" >> $out 2>> gpurun_out/r02c.err; }
run "blk-1 chunk256M" KV_TILE_BLOCK_LOG2=-1
run "blk4 chunk256M" KV_TILE_BLOCK_LOG2=4
run "blk6 chunk256M" KV_TILE_BLOCK_LOG2=6
run "blk8 chunk256M" KV_TILE_BLOCK_LOG2=8
run "blk6 chunk64M" KV_TILE_BLOCK_LOG2=6 KV_TILE_CHUNK_BASES=67108864
run "blk6 chunk128M" KV_TILE_BLOCK_LOG2=6 KV_TILE_CHUNK_BASES=134217728
run "blk6 chunk512M" KV_TILE_BLOCK_LOG2=6 KV_TILE_CHUNK_BASES=536870912
run "blk6 chunk1G" KV_TILE_BLOCK_LOG2=6 KV_TILE_CHUNK_BASES=1073741824
run "blk6 rb14" KV_TILE_BLOCK_LOG2=6 KV_TILE_RB=14
run "blk6 rb16" KV_TILE_BLOCK_LOG2=6 KV_TILE_RB=16
cat $out | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['label'], d['count_x3']['G_kmers_per_s'], d['count_x3']['kernel_ms'])
"
KV_UPDATE_PATH=tile timeout 600 ncu --set full --import-source on --clock-control none -k regex:'kv_hash_kernel|kv_tile_apply' -s 6 -c 4 -o gpurun_out/r02c_tile_4g python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 1 > gpurun_out/r02c_ncu.log 2>&1
tail -3 gpurun_out/r02c_ncu.log
