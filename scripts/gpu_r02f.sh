#!/bin/bash
# Round 2, call F (2 GPUs): C-ABI merge (gsb_tsdf_reduce over NCCL) against sequential fusion; 2-GPU bench (weak + strong)
mkdir -p gpurun_out
T=gpurun_out/r02f
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > ${T}_dist_check.log 2>&1
echo "dist check exit $?"; tail -3 ${T}_dist_check.log
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], d['ms_per_step'], d['details']['volume_merge'])" || tail -8 $1; }
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > ${T}_bench_2gpu_weak20.log 2>&1; show ${T}_bench_2gpu_weak20.log weak20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline --scaling strong > ${T}_bench_2gpu_strong200.log 2>&1; show ${T}_bench_2gpu_strong200.log strong200
