#!/bin/bash
# ncu evidence after the last kernel changes of round 2 (C2, one GPU): launch list + full captures of the step's kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02zd_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-c3 --no-variants --no-cpu-baseline --no-files > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:'kv_hash_kernel|kv_first_compact|kv_first_min_list|kv_first_own_list|kv_increment_kernel|kv_novel_kernel|kv_occ_rebuild' -s 60 -c 12 -o gpurun_out/r02zd_c2 python bench.py --steps 2 --warmup 3 --no-c3 --no-variants --no-cpu-baseline --no-files > gpurun_out/r02zd_c2.log 2>&1
ncu -i gpurun_out/r02zd_c2.ncu-rep --page raw --csv > gpurun_out/r02zd_ncu_raw_c2.csv 2>/dev/null
ls -la gpurun_out/r02zd*
