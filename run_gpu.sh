set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_warp2.log 2>&1
tail -1 gpurun_out/bench_warp2.log | cut -c1-200
