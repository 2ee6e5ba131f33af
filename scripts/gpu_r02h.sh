#!/bin/bash
# Round 2, call H (1 GPU): whole GPU suite incl. full-size parity, smoke, final-shape bench (both arms), ncu of every kernel
mkdir -p gpurun_out
T=gpurun_out/r02h
rm -f gpurun_out/parity_observed.json gpurun_out/parity_session_observed.json
timeout 1200 python -m pytest tests -m gpu -q > ${T}_tests.log 2>&1
echo "tests exit $? : $(tail -1 ${T}_tests.log)"
grep -E "^(FAILED|ERROR)" ${T}_tests.log | head -20
cp gpurun_out/parity_observed.json ${T}_parity_observed.json 2>/dev/null; cp gpurun_out/parity_session_observed.json ${T}_parity_session_observed.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${T}_smoke.log 2>&1; tail -1 ${T}_smoke.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > ${T}_bench_reference.log 2>&1; tail -c 600 ${T}_bench_reference.log
timeout 400 python bench.py --steps 200 > ${T}_bench_final.log 2>&1
grep -h '^{"metric' ${T}_bench_final.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('final', d['value'], d['e2e']['value'], d['roofline'], d['cpu_baseline'])" || tail -5 ${T}_bench_final.log
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'render_table_kernel|preprocess_kernel|emit_sorted_kernel|scan_tiles_kernel|mark_bricks_kernel|radix_pass_kernel|radix_histogram_kernel|integrate_kernel|to_u8_kernel|prepare_depth_kernel|init_ranges_kernel' -s 150 -c 40 \
    -o ${T}_prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline > ${T}_ncu_full.log 2>&1
ls -la gpurun_out | grep r02h
