#!/bin/bash
# Round 2, call X (1 GPU): software-pipelined blend loop (alpha of the next pair in flight during this pair's T chain), record
# indices prefetched two chunks ahead -- tests, C1 / C3 bench, per-warp trace
mkdir -p gpurun_out
T=gpurun_out/r02x
timeout 900 python -m pytest tests -m gpu -q -x > ${T}_tests.log 2>&1; echo "tests exit $? : $(tail -1 ${T}_tests.log)"
grep -E "^(FAILED|ERROR)" ${T}_tests.log | head
show() { grep -h '^{"metric' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']; print('$2', d['value'], d['e2e']['value'], 'render', k['render']['avg_ms'], 'serial', d['kernels_note'].split('(')[1].split(' ms')[0])" || tail -5 $1; }
timeout 300 python bench.py --steps 100 --no-cpu-baseline > ${T}_bench_C1.log 2>&1; show ${T}_bench_C1.log C1
timeout 300 python bench.py --steps 40 --no-cpu-baseline --config C3 > ${T}_bench_C3.log 2>&1; show ${T}_bench_C3.log C3
timeout 300 python scripts/blend_trace.py 2> ${T}.err | grep '^{"config' > ${T}_blend_trace_C1.json
timeout 300 python scripts/blend_trace.py --config C3 2>> ${T}.err | grep '^{"config' > ${T}_blend_trace_C3.json
python - <<'PY'
import json
for c in ("C1","C3"):
    d=json.load(open(f"gpurun_out/r02x_blend_trace_{c}.json"))
    v=d["views"][0]
    print(c,"span",v["span_us"],"sum",v["sum_warp_us"],"cycles b/b/w",v["cycles_build_blend_wait_all_warps"])
    for x in v["longest_chains"][:3]: print("  long:",x["list"],x["walked"],x["dur_us"],x["blended"],x["cycles_per_chunk_build_blend_wait"])
PY
