#!/bin/bash
# scratch: GPU run 22 - sorted tiles delivery + hoisted histogram loads: tests, bench, sanitizers on the new kernels
mkdir -p gpurun_out
T=gpurun_out/run22
timeout 400 python -m pytest tests -m gpu -x -q > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench.log 2>&1
grep -h '^{"metric' ${T}_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})" || tail -3 ${T}_bench.log
SEL='blend_kernel_variants or sh_staging or tma_staging or undersized or sort'
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_raster.py tests/test_gpu_tsdf.py -m gpu -x -q -k "$SEL or plane or repeat" > ${T}_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> ${T}_sanitizer_memcheck.log; tail -3 ${T}_sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_raster.py -m gpu -x -q -k "blend_kernel_variants or tma_staging" > ${T}_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> ${T}_sanitizer_racecheck.log; tail -3 ${T}_sanitizer_racecheck.log
