#!/bin/bash
# Round 2, call 3A (1 GPU): more hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS) x pairs in flight, C1 and C3
mkdir -p gpurun_out
T=gpurun_out/r03a
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], d['ms_per_step'], d['details'].get('host_enqueue_ms_per_step'))" || tail -5 $1; }
for conn in 8 32; do for pif in 2 3 4; do
  CUDA_DEVICE_MAX_CONNECTIONS=$conn GSB_PAIRS_IN_FLIGHT=$pif timeout 300 python bench.py --steps 100 --no-cpu-baseline > ${T}_C1_conn${conn}_pif$pif.log 2>&1; show ${T}_C1_conn${conn}_pif$pif.log C1_conn${conn}_pif$pif
done; done
for pif in 2 3 4; do
  CUDA_DEVICE_MAX_CONNECTIONS=32 GSB_PAIRS_IN_FLIGHT=$pif timeout 300 python bench.py --steps 40 --no-cpu-baseline --config C3 > ${T}_C3_conn32_pif$pif.log 2>&1; show ${T}_C3_conn32_pif$pif.log C3_conn32_pif$pif
done
