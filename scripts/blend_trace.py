"""Per-warp timeline of the table blend on one C1 (or --config) view: who ends last, how long the longest serial chains are,
how many warp slots are busy over the kernel's span.  Profiling aid (gsb_debug_blend_trace); prints one JSON object."""
import argparse, ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gs2mesh_b200 import _lib, scene, rasterizer as rast
from gs2mesh_b200.renderer import Renderer

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--config", default="C1"); ap.add_argument("--views", default="0,57")
    a = ap.parse_args()
    cfg = scene.CONFIGS[a.config]
    cloud, rigs, baseline = bench.build_scene(cfg, 200)
    r = Renderer.from_scene(rigs, baseline, cloud, output_dir_root=None, args=bench.BenchArgs(cfg), device="cuda:0")
    r.prepare_renderer()
    W, H = cfg["width"], cfg["height"]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    L = _lib.lib(); L.gsb_debug_blend_trace.argtypes = [C.c_void_p]; L.gsb_debug_blend_trace.restype = C.c_int
    trace = torch.zeros(gx * gy * 4 * 8, dtype=torch.int64, device="cuda:0")
    out = {"config": a.config, "views": []}
    def render(i):
        return rast.rasterize_forward(means3D=r.means3D, opacities=r.opacity, viewmatrix=r._camera_table[i, 0, 0:16], projmatrix=r._camera_table[i, 0, 16:32],
                                      campos=r._camera_table[i, 0, 32:35], bg=r.background, width=W, height=H, tan_fovx=r._views[i][0].tan_fovx,
                                      tan_fovy=r._views[i][0].tan_fovy, shs=r.shs, scales=r.scales, rotations=r.rotations, sh_degree=r.sh_degree)
    for i in [int(v) for v in a.views.split(",")]:
        render(i); render(i); torch.cuda.synchronize()
        trace.zero_(); _lib.check(L.gsb_debug_blend_trace(trace.data_ptr()))
        render(i); torch.cuda.synchronize()
        _lib.check(L.gsb_debug_blend_trace(None))
        t = trace.cpu().numpy().astype(np.uint64).reshape(gy * gx, 4, 8)
        cyc = t[..., 3:6].astype(np.int64).reshape(-1, 3); hits = t[..., 6].astype(np.int64).reshape(-1)
        st, en, meta = t[..., 0].astype(np.int64), t[..., 1].astype(np.int64), t[..., 2]
        total, done = (meta >> np.uint64(32)).astype(np.int64), (meta & np.uint64(0xffffffff)).astype(np.int64)
        ok = st > 0
        t0, t1 = st[ok].min(), en[ok].max(); span = float(t1 - t0)
        dur = np.where(ok, en - st, 0).astype(np.float64)
        order = np.argsort(-en.reshape(-1))[:8]
        last = [{"tile_xy": [int((k // 4) % gx), int((k // 4) // gx)], "block": int(k % 4), "list": int(total.reshape(-1)[k]), "walked": int(done.reshape(-1)[k]),
                 "start_us": round((st.reshape(-1)[k] - t0) / 1e3, 1), "end_us": round((en.reshape(-1)[k] - t0) / 1e3, 1)} for k in order]
        longest = np.argsort(-dur.reshape(-1))[:8]
        chains = [{"tile_xy": [int((k // 4) % gx), int((k // 4) // gx)], "list": int(total.reshape(-1)[k]), "walked": int(done.reshape(-1)[k]),
                   "start_us": round((st.reshape(-1)[k] - t0) / 1e3, 1), "dur_us": round(dur.reshape(-1)[k] / 1e3, 1), "blended": int(hits[k]),
                   "cycles_per_chunk_build_blend_wait": [int(c // max((done.reshape(-1)[k] + 31) // 32, 1)) for c in cyc[k]]} for k in longest]
        bins = 20; edges = np.linspace(t0, t1, bins + 1); busy = []
        for b in range(bins):  # mean number of resident blend warps in each twentieth of the span
            lo, hi = edges[b], edges[b + 1]
            busy.append(round(float(np.clip(np.minimum(en[ok], hi) - np.maximum(st[ok], lo), 0, None).sum() / (hi - lo)), 0))
        nz = total > 0
        out["views"].append({"view": i, "span_us": round(span / 1e3, 1), "warps": int(ok.sum()), "warps_with_work": int(nz.sum()),
                             "sum_warp_us": round(dur.sum() / 1e3, 0), "mean_resident_warps": round(dur.sum() / span, 0), "slots": 148 * 28,
                             "walked_fraction_of_lists": round(float(done.sum() / max(total.sum(), 1)), 3),
                             "blended_fraction_of_walked": round(float(hits.sum() / max(done.sum(), 1)), 3),
                             "cycles_build_blend_wait_all_warps": [int(c) for c in cyc.sum(0)],
                             "resident_warps_per_twentieth": busy, "last_to_end": last, "longest_chains": chains,
                             "dur_us_percentiles_50_90_99_max": [round(float(np.percentile(dur[nz], q)) / 1e3, 1) for q in (50, 90, 99, 100)]})
    print(json.dumps(out))

if __name__ == "__main__":
    main()
