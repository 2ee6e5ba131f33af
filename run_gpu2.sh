set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check.log 2>&1
echo "dist check exit $?" >> gpurun_out/dist_check.log
tail -3 gpurun_out/dist_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1
echo "bench2 exit $?" >> gpurun_out/bench_2gpu.log
tail -2 gpurun_out/bench_2gpu.log | cut -c1-400
