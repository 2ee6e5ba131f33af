#!/bin/bash
# Round 2, call Z (1 GPU): deep lists finished by render_deep_kernel (rest of the list split over 16 warps) -- tests, cap sweep
mkdir -p gpurun_out
T=gpurun_out/r02z
timeout 900 python -m pytest tests -m gpu -q -x > ${T}_tests.log 2>&1; echo "tests exit $? : $(tail -1 ${T}_tests.log)"
grep -E "^(FAILED|ERROR)|Error|assert" ${T}_tests.log | head -12
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('$2', d['value'], d['e2e']['value'], 'render', k['render']['avg_ms'], 'serial', d['kernels_note'].split('(')[1].split(' ms')[0])" || tail -5 $1; }
for c in 0 512 1024 2048; do
  GSB_DEEP_CAP=$c timeout 300 python bench.py --steps 100 --no-cpu-baseline > ${T}_bench_C1_cap$c.log 2>&1; show ${T}_bench_C1_cap$c.log C1_cap$c
done
for c in 0 1024; do
  GSB_DEEP_CAP=$c timeout 300 python bench.py --steps 40 --no-cpu-baseline --config C3 > ${T}_bench_C3_cap$c.log 2>&1; show ${T}_bench_C3_cap$c.log C3_cap$c
done
