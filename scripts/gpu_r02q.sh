#!/bin/bash
# Round 2, call Q (1 GPU): whole GPU suite on the final tree (incl. the ragged-input pair tests), default bench line with CPU baseline
mkdir -p gpurun_out
T=gpurun_out/r02q
rm -f gpurun_out/parity_observed.json gpurun_out/parity_session_observed.json
timeout 1200 python -m pytest tests -m gpu -q > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
grep -E "^(FAILED|ERROR)" ${T}_tests.log | head -20
cp gpurun_out/parity_observed.json ${T}_parity_observed.json 2>/dev/null; cp gpurun_out/parity_session_observed.json ${T}_parity_session_observed.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; tail -1 ${T}_smoke.log
timeout 400 python bench.py > ${T}_bench_default.log 2>&1
grep -h '^{"metric' ${T}_bench_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'])" || tail -5 ${T}_bench_default.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > ${T}_bench_reference.log 2>&1
grep -h '^{"impl' ${T}_bench_reference.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('reference', d['value'], d['config'], d['details']['reference_rasterizer_ms_per_view'])" || tail -5 ${T}_bench_reference.log
