#!/bin/bash
# Last single-GPU gpurun script of round 1 (final-state verification + A/B of the tile-sort digit split); run from the repo root:
#   gpurun --timeout 600 -- "bash scripts/gpu_run_last.sh"
mkdir -p gpurun_out
T=gpurun_out/run25
timeout 200 python -m pytest tests -m gpu -x -q > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], {k:round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()})" || tail -3 $1; }
timeout 200 python bench.py --cpu-budget 8 > ${T}_bench_final.log 2>&1; show ${T}_bench_final.log final
GSB_RADIX_SPLIT=even timeout 200 python -m pytest tests/test_gpu_raster.py tests/test_gpu_pipeline.py -m gpu -x -q > ${T}_tests_even.log 2>&1
echo "tests[even] exit $? : $(tail -1 ${T}_tests_even.log)"
GSB_RADIX_SPLIT=even timeout 100 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_even.log 2>&1; show ${T}_bench_even.log even
GSB_RADIX_R_ITEMS=8 timeout 100 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_r8.log 2>&1; show ${T}_bench_r8.log r8
