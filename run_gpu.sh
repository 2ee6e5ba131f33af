#!/bin/bash
# scratch: GPU run 23 - solo (one warp per CTA) blend variant: parity + A/B
mkdir -p gpurun_out
T=gpurun_out/run23
timeout 300 python -m pytest tests/test_gpu_raster.py tests/test_gpu_pipeline.py -m gpu -x -q > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
GSB_RENDER_IMPL=s timeout 300 python -m pytest tests/test_gpu_raster.py tests/test_gpu_pipeline.py -m gpu -x -q > ${T}_tests_solo.log 2>&1
echo "tests[solo] exit $? : $(tail -1 ${T}_tests_solo.log)"
run_bench() {  # name, extra bench args (quoted), env...
  local name=$1; local extra=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline $extra > ${T}_bench_$name.log 2>&1
  grep -h '^{"metric' ${T}_bench_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench[$name]', d['value'], d['e2e']['value'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})" || tail -3 ${T}_bench_$name.log
}
run_bench dual "" GSB_RENDER_IMPL=d
run_bench solo "" GSB_RENDER_IMPL=s
run_bench C3_dual "--config C3 --steps 60" GSB_RENDER_IMPL=d
run_bench C3_solo "--config C3 --steps 60" GSB_RENDER_IMPL=s
