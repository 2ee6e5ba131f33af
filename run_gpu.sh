#!/bin/bash
# scratch: GPU run 15 - tight candidate rectangles
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/run15_tests.log 2>&1
tail -5 gpurun_out/run15_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/run15_bench.log 2>&1
grep -h '^{"metric' gpurun_out/run15_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['config']['per_view'], {k:v['avg_ms'] for k,v in d['kernels'].items()})"
