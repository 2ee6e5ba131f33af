#!/bin/bash
# Round 2, call E (1 GPU): CTA-level heavy-first tile queue A/B, full-size parity C1..C4 with the round-2 kernels
mkdir -p gpurun_out
T=gpurun_out/r02e
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], {k:round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()})" || tail -5 $1; }
timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_default.log 2>&1; show ${T}_bench_default.log default
GSB_RENDER_QUEUE=0 timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_noqueue.log 2>&1; show ${T}_bench_noqueue.log noqueue
timeout 300 python bench.py --no-cpu-baseline --steps 60 --config C3 > ${T}_bench_C3.log 2>&1; show ${T}_bench_C3.log C3
GSB_RENDER_QUEUE=0 timeout 300 python bench.py --no-cpu-baseline --steps 60 --config C3 > ${T}_bench_C3_noqueue.log 2>&1; show ${T}_bench_C3_noqueue.log C3noqueue
rm -f gpurun_out/parity_observed.json
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -k "full_size" > ${T}_parity.log 2>&1
echo "parity exit $? : $(tail -1 ${T}_parity.log)"
cp gpurun_out/parity_observed.json ${T}_parity_observed.json 2>/dev/null
