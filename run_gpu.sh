#!/bin/bash
# scratch: GPU run 16 - verify HEAD after session restart: full gpu tests, bench (default), ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/run16_smi.log 2>&1
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/run16_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/run16_tests.log
tail -6 gpurun_out/run16_tests.log
timeout 600 python bench.py > gpurun_out/run16_bench.log 2>&1
grep -h '^{"metric' gpurun_out/run16_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['config']['per_view'], {k:v['avg_ms'] for k,v in d['kernels'].items()}, d['cpu_baseline'])"
timeout 800 bash scripts/profile_gpu.sh r01e 3 > gpurun_out/run16_profile.log 2>&1
ls -la gpurun_out | tail -12
