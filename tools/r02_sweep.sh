#!/bin/bash
# large-sketch sweep: update path x sketch size (count x3 + novel, reads drawn on the device)
out=gpurun_out/r02b_bigsketch.jsonl
: > $out
for mem in 1e9 4e9 16e9; do
  for path in tile part direct; do
    KV_UPDATE_PATH=$path timeout 300 python tools/bigsketch_bench.py --memory $mem --genome 30000000 --reads 9000000 --steps 2 --label "$path" >> $out 2>> gpurun_out/r02b_bigsketch.err
  done
done
KV_UPDATE_PATH=tile KV_TILE_RB=16 timeout 300 python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 2 --label "tile rb16" >> $out 2>> gpurun_out/r02b_bigsketch.err
KV_UPDATE_PATH=tile KV_TILE_RB=14 timeout 300 python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 2 --label "tile rb14" >> $out 2>> gpurun_out/r02b_bigsketch.err
KV_UPDATE_PATH=tile KV_TILE_CHUNK_BASES=67108864 timeout 300 python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 2 --label "tile chunk64M" >> $out 2>> gpurun_out/r02b_bigsketch.err
KV_UPDATE_PATH=tile timeout 300 python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 2 --track --label "tile tracked" >> $out 2>> gpurun_out/r02b_bigsketch.err
KV_UPDATE_PATH=tile timeout 300 python tools/bigsketch_bench.py --memory 2e9 --bits 4 --genome 30000000 --reads 9000000 --steps 2 --label "tile 4bit" >> $out 2>> gpurun_out/r02b_bigsketch.err
KV_UPDATE_PATH=direct timeout 300 python tools/bigsketch_bench.py --memory 2e9 --bits 4 --genome 30000000 --reads 9000000 --steps 2 --label "direct 4bit" >> $out 2>> gpurun_out/r02b_bigsketch.err
cat $out
