#!/bin/bash
# Round 2, call B (1 GPU): first hardware run of the round-2 kernels (table blend, fused pair preprocess + shared depth sort,
# emit+scan, unbounded TSDF brick pool): GPU tests (all, not -x), smoke, bench A/Bs.
mkdir -p gpurun_out
T=gpurun_out/r02b
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_pipeline.py::test_full_size_parity_vs_reference_binary_and_oracle > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
grep -E "^(FAILED|ERROR)" ${T}_tests.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; tail -2 ${T}_smoke.log
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], {k:round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()})" || tail -5 $1; }
timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_default.log 2>&1; show ${T}_bench_default.log default
GSB_RENDER_IMPL=dual timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_dual.log 2>&1; show ${T}_bench_dual.log dual
GSB_PAIR_MODE=separate timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_separate.log 2>&1; show ${T}_bench_separate.log separate
GSB_PAIR_MODE=noshare timeout 300 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_noshare.log 2>&1; show ${T}_bench_noshare.log noshare
