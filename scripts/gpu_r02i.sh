#!/bin/bash
mkdir -p gpurun_out
T=gpurun_out/r02i
timeout 300 python bench.py --no-cpu-baseline --steps 200 > ${T}_bench.log 2>&1
grep -h '^{"metric' ${T}_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('final', d['value'], d['e2e']['value'], d['ms_per_step'], d['details']['host_enqueue_ms_per_step'])" || tail -5 ${T}_bench.log
timeout 300 python -X importtime -c "pass" 2>/dev/null
timeout 300 python - <<'PY' > ${T}_hostprof.log 2>&1
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from gs2mesh_b200 import scene
from gs2mesh_b200.renderer import Renderer
from gs2mesh_b200.tsdf import TSDF
import bench
cfg = scene.CONFIGS["C1"]
cloud, rigs, baseline = bench.build_scene(cfg, 60)
bargs = bench.BenchArgs(cfg)
r = Renderer.from_scene(rigs, baseline, cloud, args=bargs, device="cuda:0")
r.prepare_renderer()
st = TSDF(r, None, bargs, "b", window_resolution=512, device="cuda:0")
st.volume = st._make_volume()
def step(i):
    out = r.render_image_pair(i, to_host=False)
    st.integrate(out["depth"], out["left_u8"], rigs[i]["left"], final_T=out["final_T"])
for i in range(10): step(i)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for i in range(10, 60): step(i)
pr.disable()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue ms/step", 1e3 * (t1 - t0) / 50)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
PY
tail -60 ${T}_hostprof.log
