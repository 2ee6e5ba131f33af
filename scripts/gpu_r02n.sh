#!/bin/bash
# Round 2, call N (1 GPU): final tree: whole GPU suite, smoke, default bench (deeper e2e pipeline), C2 / C3 / C3-1080p / C4 single GPU
mkdir -p gpurun_out
T=gpurun_out/r02n
timeout 1200 python -m pytest tests -m gpu -q > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
grep -E "^(FAILED|ERROR)" ${T}_tests.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; tail -1 ${T}_smoke.log
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], d['ms_per_step'], {k:round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()})" || tail -5 $1; }
timeout 400 python bench.py --steps 200 > ${T}_bench_C1.log 2>&1; show ${T}_bench_C1.log C1
for C in C2 C3 C4; do timeout 400 python bench.py --no-cpu-baseline --steps 49 --config $C > ${T}_bench_$C.log 2>&1; show ${T}_bench_$C.log $C; done
