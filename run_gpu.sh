set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python tests/golden/make_raster_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_ours.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2>&1
echo "bench ref exit $?" >> gpurun_out/bench_ref.log
tail -5 gpurun_out/pytest_gpu.log
