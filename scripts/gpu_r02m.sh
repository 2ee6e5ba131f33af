#!/bin/bash
# Round 2, call M (2 GPUs): merge phase breakdown
mkdir -p gpurun_out
T=gpurun_out/r02m
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > ${T}_bench_2gpu_weak20.log 2>&1
grep -h '^{"metric' ${T}_bench_2gpu_weak20.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['details']['volume_merge'])" || tail -8 ${T}_bench_2gpu_weak20.log
