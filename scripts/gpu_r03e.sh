#!/bin/bash
# Round 2, last call (1 GPU, ~3.5 GPU-minutes left): the final tree once more -- whole GPU suite, smoke, default bench line.
# Every step stamps its wall time so that a call cut short by the budget still says how far it got.
mkdir -p gpurun_out
T=gpurun_out/r03e
S=$(date +%s)
stamp() { echo "[+$(( $(date +%s) - S )) s] $*" | tee -a ${T}_timeline.log; }
rm -f gpurun_out/parity_observed.json gpurun_out/parity_session_observed.json
stamp start
timeout 170 python -m pytest tests -m gpu -x -q > ${T}_tests.log 2>&1
stamp "tests exit $? : $(tail -1 ${T}_tests.log)"
cp gpurun_out/parity_observed.json ${T}_parity_observed.json 2>/dev/null; cp gpurun_out/parity_session_observed.json ${T}_parity_session_observed.json 2>/dev/null
timeout 90 python bench.py --cpu-budget 8 > ${T}_bench_default.log 2>&1
stamp "bench exit $?"
grep -h '^{"metric' ${T}_bench_default.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], {k: v['avg_ms'] for k, v in d['kernels'].items()})" || tail -5 ${T}_bench_default.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; stamp "smoke: $(tail -1 ${T}_smoke.log)"
