#!/bin/bash
# Round 2, call U (1 GPU): blend CTA granularity -- 4 / 2 / 1 warps (8x8 blocks) per CTA on C1 and C3
mkdir -p gpurun_out
T=gpurun_out/r02u
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('$2', d['value'], d['e2e']['value'], 'render', k['render']['avg_ms'], 'serial', d['kernels_note'].split('(')[1].split(' ms')[0])" || tail -5 $1; }
for w in 4 2 1; do
  GSB_BLEND_WARPS=$w timeout 300 python bench.py --steps 100 --no-cpu-baseline > ${T}_bench_C1_warps$w.log 2>&1; show ${T}_bench_C1_warps$w.log C1_warps$w
done
for w in 4 2 1; do
  GSB_BLEND_WARPS=$w timeout 300 python bench.py --steps 40 --no-cpu-baseline --config C3 > ${T}_bench_C3_warps$w.log 2>&1; show ${T}_bench_C3_warps$w.log C3_warps$w
done
GSB_BLEND_WARPS=1 timeout 600 python -m pytest tests/test_gpu_raster.py tests/test_gpu_pair.py -m gpu -q -x > ${T}_tests_warps1.log 2>&1; echo "tests(warps=1) exit $? : $(tail -1 ${T}_tests_warps1.log)"
