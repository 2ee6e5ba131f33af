#!/bin/bash
# Round 2, call L (N GPUs): final-tree scaling numbers with the merge timed (scratch pre-sized): C1 weak 20 pairs/rank (driver shape),
# C1 strong 200, reduce vs all-reduce, C4 (N=8) / C3 (N=4)
mkdir -p gpurun_out
N=${1:-8}
T=gpurun_out/r02l
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], d['ms_per_step'], d['details']['volume_merge'], d['details']['tsdf_volume']['bricks_open_after_timed_region'])" || tail -8 $1; }
run 29511 tests/dist_gpu_check.py > ${T}_dist_check_${N}.log 2>&1; echo "dist check exit $?"; tail -1 ${T}_dist_check_${N}.log
run 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > ${T}_bench_${N}gpu_C1_weak20.log 2>&1; show ${T}_bench_${N}gpu_C1_weak20.log C1weak20
BENCH_MERGE=allreduce run 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > ${T}_bench_${N}gpu_C1_weak20_allreduce.log 2>&1; show ${T}_bench_${N}gpu_C1_weak20_allreduce.log C1weak20allreduce
run 29514 bench.py --gpus $N --steps 200 --warmup 5 --no-cpu-baseline --scaling strong > ${T}_bench_${N}gpu_C1_strong200.log 2>&1; show ${T}_bench_${N}gpu_C1_strong200.log C1strong200
if [ "$N" = "8" ]; then C=C4; else C=C3; fi
run 29515 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --config $C > ${T}_bench_${N}gpu_${C}.log 2>&1; show ${T}_bench_${N}gpu_${C}.log $C
