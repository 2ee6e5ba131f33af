#!/bin/bash
# Round 2, call A (1 GPU): Open3D acquisition attempt on the box, pipe micro-benchmark, full-size parity on C1..C4 of the
# round-1 kernels (observed outlier numbers), baseline bench on today's box.
#   gpurun --timeout 1500 -- "bash scripts/gpu_r02a.sh"
mkdir -p gpurun_out
T=gpurun_out/r02a
bash scripts/try_open3d.sh gpurun_out/r02_open3d_attempt_gpubox.log > /dev/null 2>&1
tail -1 gpurun_out/r02_open3d_attempt_gpubox.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${T}_smi.log 2>&1
timeout 120 scripts/ubench_pipes > ${T}_ubench.log 2>&1; cat ${T}_ubench.log
rm -f gpurun_out/parity_observed.json
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -k "full_size_parity" > ${T}_parity.log 2>&1
echo "parity exit $? : $(tail -1 ${T}_parity.log)"
cp gpurun_out/parity_observed.json ${T}_parity_observed.json 2>/dev/null
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], {k:round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()})" || tail -3 $1; }
timeout 200 python bench.py --no-cpu-baseline --steps 100 > ${T}_bench_C1.log 2>&1; show ${T}_bench_C1.log C1
