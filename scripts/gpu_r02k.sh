#!/bin/bash
# Round 2, call K (1 GPU): ncu capture of one steady-state pair incl. the TSDF kernels (traffic.json), pairs-in-flight A/B
mkdir -p gpurun_out
T=gpurun_out/r02k
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], d['ms_per_step'])" || tail -5 $1; }
GSB_PAIRS_IN_FLIGHT=3 timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_pif3.log 2>&1; show ${T}_bench_pif3.log pif3
GSB_PAIRS_IN_FLIGHT=1 timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_pif1.log 2>&1; show ${T}_bench_pif1.log pif1
timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_pif2.log 2>&1; show ${T}_bench_pif2.log pif2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file ${T}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > ${T}_ncu_bench.log 2>&1
grep -c . ${T}_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'render_table_kernel|preprocess_kernel|emit_sorted_kernel|scan_tiles_kernel|mark_bricks_kernel|radix_pass_kernel|radix_histogram_kernel|integrate_kernel|to_u8_kernel|prepare_depth_kernel|init_ranges_kernel' -s 640 -c 56 \
    -o ${T}_prof python bench.py --steps 4 --warmup 3 --no-cpu-baseline > ${T}_ncu_full.log 2>&1
ls -la gpurun_out | grep r02k
