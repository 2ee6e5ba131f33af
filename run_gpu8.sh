set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/ngpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/bench_8gpu_C1.log 2>&1
echo "exit $?" >> gpurun_out/bench_8gpu_C1.log
tail -2 gpurun_out/bench_8gpu_C1.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/bench_4gpu_C1.log 2>&1
echo "exit $?" >> gpurun_out/bench_4gpu_C1.log
tail -2 gpurun_out/bench_4gpu_C1.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --config C4 --steps 62 --warmup 3 > gpurun_out/bench_8gpu_C4.log 2>&1
echo "exit $?" >> gpurun_out/bench_8gpu_C4.log
tail -2 gpurun_out/bench_8gpu_C4.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --config C3 --steps 75 --warmup 3 > gpurun_out/bench_4gpu_C3.log 2>&1
echo "exit $?" >> gpurun_out/bench_4gpu_C3.log
tail -2 gpurun_out/bench_4gpu_C3.log | cut -c1-300
