set -x
mkdir -p gpurun_out
for c in C2 C3 C4; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.log 2>&1
  echo "exit $?" >> gpurun_out/bench_$c.log
  tail -2 gpurun_out/bench_$c.log | cut -c1-400
done
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1
tail -1 gpurun_out/bench_ref.log | cut -c1-900
