set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check.log 2>&1
echo "dist check exit $?" >> gpurun_out/dist_check.log
tail -2 gpurun_out/dist_check.log
for mode in "" "--dense-reduce"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 $mode > gpurun_out/bench_2gpu$mode.log 2>&1
tail -1 gpurun_out/bench_2gpu$mode.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['volume_merge'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config C4 --steps 30 --warmup 3 $mode > gpurun_out/bench_2gpu_C4$mode.log 2>&1
tail -1 gpurun_out/bench_2gpu_C4$mode.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['volume_merge'])"
done
