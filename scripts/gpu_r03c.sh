#!/bin/bash
# Round 2, call 3C (1 GPU): pairs in flight 3 / 2 / 3 / 2 on one box, default step count (e2e is host / PCIe sensitive: compare within a box)
mkdir -p gpurun_out
T=gpurun_out/r03c
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', d['value'], d['e2e']['value'], d['ms_per_step'])" || tail -5 $1; }
n=0
for pif in 3 2 3 2; do n=$((n+1))
  GSB_PAIRS_IN_FLIGHT=$pif timeout 200 python bench.py --no-cpu-baseline > ${T}_C1_pif${pif}_run$n.log 2>&1; show ${T}_C1_pif${pif}_run$n.log pif${pif}_run$n
done
