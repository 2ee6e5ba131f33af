set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_ours.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_ours.log
tail -3 gpurun_out/bench_ours.log | cut -c1-300
