#!/bin/bash
# Round 2 (1 GPU, ~2.4 GPU-minutes left): host-side pipelining switches of the stage class, all OFF by default --
#   GSB_COPY_STREAMS=1       uint8 frame D2H copies on their own streams (the render streams are free for the next pair at once)
#   GSB_SPARE_BUFFER_SETS=1  one more buffer set, so a pair's renders are gated on the call before the previous one
#   BENCH_E2E_ORDER=overlap  bench e2e loop: depth read-back enqueued before the next render call, fused after it
# whole GPU suite with the renderer switches ON, then bench lines: all off / all on / loop order alone.
mkdir -p gpurun_out
T=gpurun_out/r03f
S=$(date +%s)
stamp() { echo "[+$(( $(date +%s) - S )) s] $*" | tee -a ${T}_timeline.log; }
line() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$2', 'value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], 'enqueue', d['details']['host_enqueue_ms_per_step'])" || tail -3 $1; }
stamp start
GSB_COPY_STREAMS=1 GSB_SPARE_BUFFER_SETS=1 timeout 100 python -m pytest tests -m gpu -x -q > ${T}_tests_switches_on.log 2>&1
stamp "tests (switches on) exit $? : $(tail -1 ${T}_tests_switches_on.log)"
timeout 40 python bench.py --no-cpu-baseline > ${T}_bench_off.log 2>&1; stamp "bench off $?"; line ${T}_bench_off.log off
GSB_COPY_STREAMS=1 GSB_SPARE_BUFFER_SETS=1 BENCH_E2E_ORDER=overlap timeout 40 python bench.py --no-cpu-baseline > ${T}_bench_on.log 2>&1; stamp "bench on $?"; line ${T}_bench_on.log on
BENCH_E2E_ORDER=overlap timeout 40 python bench.py --no-cpu-baseline > ${T}_bench_order.log 2>&1; stamp "bench order $?"; line ${T}_bench_order.log order
