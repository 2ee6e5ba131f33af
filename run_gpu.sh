#!/bin/bash
# scratch: GPU run 24 - ncu capture of the TSDF-side kernels + C2/C3/C4 with the final defaults
mkdir -p gpurun_out
T=gpurun_out/run24
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'integrate_kernel|mark_bricks|prepare_depth|to_u8_kernel' -s 24 -c 16 \
    -o gpurun_out/r01m_prof_tsdf python bench.py --steps 3 --warmup 3 --no-cpu-baseline > ${T}_ncu_tsdf.log 2>&1
tail -2 ${T}_ncu_tsdf.log | cut -c1-200
run_bench() {  # name, extra bench args
  local name=$1; local extra=$2
  timeout 400 python bench.py --warmup 5 --no-cpu-baseline $extra > ${T}_bench_$name.log 2>&1
  grep -h '^{"metric' ${T}_bench_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench[$name]', d['value'], d['e2e']['value'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})" || tail -3 ${T}_bench_$name.log
}
run_bench C2 "--config C2 --steps 49"
run_bench C3 "--config C3 --steps 100"
run_bench C4 "--config C4 --steps 60"
