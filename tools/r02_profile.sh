#!/bin/bash
# ncu evidence for the round-2 kernels (one GPU).  C2 (L2-resident sketches) from bench.py, C3 (4 GB sketches)
# from tools/bigsketch_bench.py.  Launch lists: gpu__time_duration only; full captures: a few launches each.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02p_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-c3 --no-variants --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02p_launches_c3.csv python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 1 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:'kv_hash_kernel|kv_first_compact|kv_first_min_list|kv_first_own_list|kv_increment_kernel|kv_novel_kernel' -s 60 -c 16 -o gpurun_out/r02p_c2 python bench.py --steps 2 --warmup 3 --no-c3 --no-variants --no-cpu-baseline > gpurun_out/r02p_c2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'kv_hash_kernel|kv_tile_apply|kv_novel_kernel' -s 8 -c 6 -o gpurun_out/r02p_c3 python tools/bigsketch_bench.py --memory 4e9 --genome 30000000 --reads 9000000 --steps 1 > gpurun_out/r02p_c3.log 2>&1
ls -la gpurun_out/r02p*
