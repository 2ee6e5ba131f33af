#!/bin/bash
# Round 2, call S (N GPUs): where the merge's time goes -- repeated merges with phase timers, raw NCCL on the same bytes (cold / pipelined)
mkdir -p gpurun_out
N=${1:-2}
T=gpurun_out/r02s
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29521 scripts/merge_probe.py > ${T}_merge_probe_${N}_twoshot.log 2>&1; grep -h '^{"world' ${T}_merge_probe_${N}_twoshot.log || tail -20 ${T}_merge_probe_${N}_twoshot.log
GSB_MERGE_ALGO=reduce run 29522 scripts/merge_probe.py > ${T}_merge_probe_${N}_onecall.log 2>&1; grep -h '^{"world' ${T}_merge_probe_${N}_onecall.log || tail -20 ${T}_merge_probe_${N}_onecall.log
GSB_MERGE_ALGO=reduce NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING,COLL run 29523 scripts/merge_probe.py > ${T}_merge_probe_${N}_debug.log 2>&1; grep -c "NCCL INFO" ${T}_merge_probe_${N}_debug.log; grep -h "Reduce:\|ncclReduce\|Reduce " ${T}_merge_probe_${N}_debug.log | head -12
