#!/bin/bash
# Round 2, call G (1 GPU): warp-deduplicated mark_bricks, preprocess at 5 CTAs/SM (default), blend at 8 CTAs/SM (A/B)
mkdir -p gpurun_out
T=gpurun_out/r02g
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_pipeline.py::test_full_size_parity_vs_reference_binary_and_oracle > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
grep -E "^(FAILED|ERROR)" ${T}_tests.log | head -20
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], {k:round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()})" || tail -5 $1; }
timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_default.log 2>&1; show ${T}_bench_default.log default
GSB_RENDER_OCC=8 timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_occ8.log 2>&1; show ${T}_bench_occ8.log renderocc8
GSB_RENDER_OCC=8 timeout 300 python bench.py --no-cpu-baseline --steps 60 --config C3 > ${T}_bench_C3_occ8.log 2>&1; show ${T}_bench_C3_occ8.log C3occ8
timeout 300 python bench.py --no-cpu-baseline --steps 60 --config C3 > ${T}_bench_C3.log 2>&1; show ${T}_bench_C3.log C3
