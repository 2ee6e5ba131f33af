#!/bin/bash
# Round 2, call C (1 GPU): ncu of the new kernels (table blend, fused pair preprocess, emit+scan, hash mark) at C1
mkdir -p gpurun_out
T=gpurun_out/r02c
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file ${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > ${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'render_table_kernel|preprocess_kernel|emit_scan_kernel|mark_bricks_kernel|radix_pass_kernel' -s 120 -c 24 \
    -o ${T}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline > ${T}_ncu_full.log 2>&1
ls -la gpurun_out | grep r02c
