#!/bin/bash
# Round 2, call R (N GPUs): two-shot merge (reduce-scatter + gather to the root) vs the single ncclReduce; raw NCCL yardstick
mkdir -p gpurun_out
N=${1:-2}
T=gpurun_out/r02r
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], d['ms_per_step'], d['details']['volume_merge'], d['details']['tsdf_volume']['bricks_open_after_timed_region'])" || tail -8 $1; }
run 29511 tests/dist_gpu_check.py > ${T}_dist_check_${N}.log 2>&1; echo "dist check exit $?"; tail -1 ${T}_dist_check_${N}.log
GSB_MERGE_ALGO=reduce run 29516 tests/dist_gpu_check.py > ${T}_dist_check_${N}_onecall.log 2>&1; echo "dist check (one call) exit $?"; tail -1 ${T}_dist_check_${N}_onecall.log
NCCL_DEBUG=INFO run 29517 scripts/nccl_probe.py > ${T}_nccl_probe_${N}.log 2>&1; grep -h '^{"world' ${T}_nccl_probe_${N}.log; grep -h -i "nvls\|channels\|via P2P" ${T}_nccl_probe_${N}.log | sort | uniq -c | sort -rn | head -8
run 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > ${T}_bench_${N}gpu_C1_weak20.log 2>&1; show ${T}_bench_${N}gpu_C1_weak20.log twoshot
GSB_MERGE_ALGO=reduce run 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > ${T}_bench_${N}gpu_C1_weak20_onecall.log 2>&1; show ${T}_bench_${N}gpu_C1_weak20_onecall.log onecall
