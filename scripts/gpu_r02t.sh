#!/bin/bash
# Round 2, call T (1 GPU): heavy tiles first in the blend launch -- threshold sweep on C1 and C3, raster tests on the new launch shape
mkdir -p gpurun_out
T=gpurun_out/r02t
timeout 600 python -m pytest tests/test_gpu_raster.py tests/test_gpu_pair.py tests/test_gpu_pipeline.py -m gpu -q -x > ${T}_tests.log 2>&1; echo "tests exit $? : $(tail -1 ${T}_tests.log)"
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('$2', d['value'], d['e2e']['value'], 'render', k['render']['avg_ms'], 'serial', d['kernels_note'].split('(')[1].split(' ms')[0])" || tail -5 $1; }
for m in 0 2 3 4 6 8; do
  GSB_HEAVY_TILE_MULT=$m timeout 300 python bench.py --steps 100 --no-cpu-baseline > ${T}_bench_C1_mult$m.log 2>&1; show ${T}_bench_C1_mult$m.log C1_mult$m
done
for m in 0 3 6; do
  GSB_HEAVY_TILE_MULT=$m timeout 300 python bench.py --steps 40 --no-cpu-baseline --config C3 > ${T}_bench_C3_mult$m.log 2>&1; show ${T}_bench_C3_mult$m.log C3_mult$m
done
