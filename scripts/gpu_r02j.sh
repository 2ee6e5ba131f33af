#!/bin/bash
# Round 2, call J (8 GPUs): C1 weak (driver shape: 20 pairs per rank) and strong (200 pairs in total), C4 (3M Gaussians, 1920x1080,
# voxel 2/1024) with the merge timed; multi-rank merge check at 8 ranks
mkdir -p gpurun_out
T=gpurun_out/r02j
N=${1:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 tests/dist_gpu_check.py > ${T}_dist_check_${N}.log 2>&1; echo "dist check exit $?"; tail -1 ${T}_dist_check_${N}.log
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], d['ms_per_step'], d['details']['volume_merge'], d['details']['tsdf_volume'])" || tail -8 $1; }
run 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > ${T}_bench_${N}gpu_C1_weak20.log 2>&1; show ${T}_bench_${N}gpu_C1_weak20.log C1weak20
run 29513 bench.py --gpus $N --steps 200 --warmup 5 --no-cpu-baseline --scaling strong > ${T}_bench_${N}gpu_C1_strong200.log 2>&1; show ${T}_bench_${N}gpu_C1_strong200.log C1strong200
run 29514 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --config C4 > ${T}_bench_${N}gpu_C4.log 2>&1; show ${T}_bench_${N}gpu_C4.log C4
