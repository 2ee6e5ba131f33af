#!/bin/bash
# Round 2, call O (1 GPU): brick discovery with overlapped probes; compute-sanitizer memcheck + racecheck on the round-2 kernels;
# final default bench line
mkdir -p gpurun_out
T=gpurun_out/r02o
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_pipeline.py::test_full_size_parity_vs_reference_binary_and_oracle > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
grep -E "^(FAILED|ERROR)" ${T}_tests.log | head -20
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], d['ms_per_step'], {k:round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()})" || tail -5 $1; }
timeout 400 python bench.py --no-cpu-baseline --steps 200 > ${T}_bench_C1.log 2>&1; show ${T}_bench_C1.log C1
SEL="test_fused_pair_equals_two_forward_calls or test_scratch_overflow_is_repaired or test_brick_volume_bit_exact_vs_oracle or test_pool_exhaustion or test_mesh_matches_open3d_restatement or test_blend_kernel_variants_agree"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pair.py tests/test_gpu_tsdf.py tests/test_gpu_mesh.py tests/test_gpu_raster.py -m gpu -q -k "$SEL" > ${T}_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" | tee -a ${T}_sanitizer_memcheck.log; tail -4 ${T}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pair.py tests/test_gpu_tsdf.py -m gpu -q -k "test_fused_pair_equals_two_forward_calls or test_brick_volume_bit_exact_vs_oracle" > ${T}_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" | tee -a ${T}_sanitizer_racecheck.log; tail -4 ${T}_sanitizer_racecheck.log
