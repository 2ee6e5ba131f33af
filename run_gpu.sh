#!/bin/bash
# scratch: GPU run 20 - final defaults: tests, bench (+cpu baseline), caller-stream priority A/B, reference arm, ncu
mkdir -p gpurun_out
T=gpurun_out/run20
timeout 400 python -m pytest tests -m gpu -x -q > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], d.get('cpu_baseline'), {k:round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()})" || tail -3 $1; }
timeout 600 python bench.py > ${T}_bench_default.log 2>&1; show ${T}_bench_default.log default
BENCH_MAIN_PRIORITY=0 timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_prio0.log 2>&1; show ${T}_bench_prio0.log prio0
GSB_PAIRS_IN_FLIGHT=1 timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_pif1.log 2>&1; show ${T}_bench_pif1.log pif1
GSB_PAIRS_IN_FLIGHT=3 timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_pif3.log 2>&1; show ${T}_bench_pif3.log pif3
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > ${T}_bench_reference.log 2>&1; tail -1 ${T}_bench_reference.log | cut -c1-600
timeout 600 bash scripts/profile_gpu.sh r01i 3 > ${T}_profile.log 2>&1
ls -la gpurun_out | tail -8
