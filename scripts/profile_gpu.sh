#!/usr/bin/env bash
# Runs on the B200 box (under gpurun): launch list + one full ncu capture of the hot kernels.
# Outputs land in gpurun_out/; summaries are copied into profiles/ by hand afterwards.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
STEPS=${2:-3}
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
# full capture of our kernels, a few launches each, skipping warm-up launches
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'render_compact_kernel|render_warp_kernel|preprocess_kernel|integrate_kernel|emit_sorted|radix_pass|radix_hist|sorted_offsets|sorted_block' -s 69 -c 46 \
    -o gpurun_out/${TAG}_prof python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out
