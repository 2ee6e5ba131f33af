#!/bin/bash
# scratch: GPU run 18 - matrix: render compact/dual x pairs in flight x depth-sort block size
mkdir -p gpurun_out
T=gpurun_out/run18
run_tests() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m pytest tests -m gpu -x -q > ${T}_tests_$name.log 2>&1
  echo "tests[$name] exit $? : $(tail -1 ${T}_tests_$name.log)"
}
run_bench() {  # name, extra bench args (quoted), env...
  local name=$1; local extra=$2; shift; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline $extra > ${T}_bench_$name.log 2>&1
  grep -h '^{"metric' ${T}_bench_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench[$name]', d['value'], d['e2e']['value'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})" || tail -3 ${T}_bench_$name.log
}
run_tests d_pif2_i4 GSB_RENDER_IMPL=d GSB_PAIRS_IN_FLIGHT=2 GSB_RADIX_P_ITEMS=4
run_tests c_pif3_i8 GSB_RENDER_IMPL=c GSB_PAIRS_IN_FLIGHT=3 GSB_RADIX_P_ITEMS=8
run_bench c_pif1 "" GSB_RENDER_IMPL=c GSB_PAIRS_IN_FLIGHT=1
run_bench d_pif1 "" GSB_RENDER_IMPL=d GSB_PAIRS_IN_FLIGHT=1
run_bench c_pif2 "" GSB_RENDER_IMPL=c GSB_PAIRS_IN_FLIGHT=2
run_bench d_pif2 "" GSB_RENDER_IMPL=d GSB_PAIRS_IN_FLIGHT=2
run_bench c_pif3 "" GSB_RENDER_IMPL=c GSB_PAIRS_IN_FLIGHT=3
run_bench d_pif3 "" GSB_RENDER_IMPL=d GSB_PAIRS_IN_FLIGHT=3
run_bench d_pif2_i8 "" GSB_RENDER_IMPL=d GSB_PAIRS_IN_FLIGHT=2 GSB_RADIX_P_ITEMS=8
run_bench d_pif2_i4 "" GSB_RENDER_IMPL=d GSB_PAIRS_IN_FLIGHT=2 GSB_RADIX_P_ITEMS=4
run_bench c_pif1_i4 "" GSB_RENDER_IMPL=c GSB_PAIRS_IN_FLIGHT=1 GSB_RADIX_P_ITEMS=4
run_bench C3_c_pif2 "--config C3 --steps 60" GSB_RENDER_IMPL=c GSB_PAIRS_IN_FLIGHT=2 GSB_RADIX_P_ITEMS=4
run_bench C3_d_pif2 "--config C3 --steps 60" GSB_RENDER_IMPL=d GSB_PAIRS_IN_FLIGHT=2 GSB_RADIX_P_ITEMS=4
run_bench C3_w_pif1 "--config C3 --steps 60" GSB_RENDER_IMPL=w GSB_PAIRS_IN_FLIGHT=1 GSB_RADIX_P_ITEMS=16 GSB_PRE_SH=s GSB_RADIX_LOOKBACK=s
