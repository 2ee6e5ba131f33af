#!/bin/bash
# Round 2, call P (1 GPU): CTA-aggregated work-list append in mark_bricks; blend skips empty chunks / even-count sentinels
mkdir -p gpurun_out
T=gpurun_out/r02p
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_pipeline.py::test_full_size_parity_vs_reference_binary_and_oracle > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
grep -E "^(FAILED|ERROR)" ${T}_tests.log | head -20
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], d['ms_per_step'], {k:round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()})" || tail -5 $1; }
timeout 400 python bench.py --no-cpu-baseline --steps 200 > ${T}_bench_C1.log 2>&1; show ${T}_bench_C1.log C1
timeout 400 python bench.py --no-cpu-baseline --steps 200 > ${T}_bench_C1_b.log 2>&1; show ${T}_bench_C1_b.log C1again
