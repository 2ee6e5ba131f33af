#!/bin/bash
# Round 2, call 3B (1 GPU): final tree (three pairs in flight by default) -- whole GPU suite, smoke, default bench line + reference arm, ncu launch list and full-set
# capture of one steady-state pair (traffic.json)
mkdir -p gpurun_out
T=gpurun_out/r03b
rm -f gpurun_out/parity_observed.json gpurun_out/parity_session_observed.json
timeout 1200 python -m pytest tests -m gpu -q > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
cp gpurun_out/parity_observed.json ${T}_parity_observed.json 2>/dev/null; cp gpurun_out/parity_session_observed.json ${T}_parity_session_observed.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; tail -1 ${T}_smoke.log
timeout 400 python bench.py > ${T}_bench_default.log 2>&1
grep -h '^{"metric' ${T}_bench_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'], {k: v['avg_ms'] for k, v in d['kernels'].items()})" || tail -5 ${T}_bench_default.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > ${T}_bench_reference.log 2>&1
grep -h '^{"impl' ${T}_bench_reference.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('reference', d['value'])" || tail -5 ${T}_bench_reference.log
ls -la gpurun_out | grep r03b
