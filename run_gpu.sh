#!/bin/bash
# scratch: GPU run 19 - new defaults (dual render, 2 pairs in flight, radix phase-1 rewrite): tests + benches
mkdir -p gpurun_out
T=gpurun_out/run19
run_tests() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m pytest tests -m gpu -x -q > ${T}_tests_$name.log 2>&1
  echo "tests[$name] exit $? : $(tail -1 ${T}_tests_$name.log)"
}
run_bench() {  # name, extra bench args (quoted), env...
  local name=$1; local extra=$2; shift; shift
  env "$@" timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline $extra > ${T}_bench_$name.log 2>&1
  grep -h '^{"metric' ${T}_bench_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench[$name]', d['value'], d['e2e']['value'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})" || tail -3 ${T}_bench_$name.log
}
run_tests default
run_bench default ""
run_bench pif1 "" GSB_PAIRS_IN_FLIGHT=1
run_bench pif3 "" GSB_PAIRS_IN_FLIGHT=3
run_bench compact "" GSB_RENDER_IMPL=c
run_bench i16 "" GSB_RADIX_P_ITEMS=16
run_bench C2 "--config C2 --steps 49"
run_bench C3 "--config C3 --steps 60"
run_bench C4 "--config C4 --steps 40"
