set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check4.log 2>&1
echo "dist check exit $?" >> gpurun_out/dist_check4.log
tail -3 gpurun_out/dist_check4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --config C4 --steps 62 --warmup 3 > gpurun_out/bench_8gpu_C4_sparse.log 2>&1
echo "exit $?" >> gpurun_out/bench_8gpu_C4_sparse.log
tail -2 gpurun_out/bench_8gpu_C4_sparse.log | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/bench_8gpu_C1_sparse.log 2>&1
echo "exit $?" >> gpurun_out/bench_8gpu_C1_sparse.log
tail -2 gpurun_out/bench_8gpu_C1_sparse.log | cut -c1-250
