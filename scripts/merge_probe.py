"""Where does the time of gsb_tsdf_reduce go?  (torchrun, one process per GPU.)  Every rank fuses three synthetic sphere views
of its own into a C1-sized lattice (voxel 2/512), then: (a) repeated merges with the phase timers on, (b) raw NCCL reduces of
the same payload through torch.distributed, one cold call at a time (device idle before, like inside the merge) and pipelined."""
import os, sys, json, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def look_at(pos):
    pos = np.asarray(pos, float); f = -pos / np.linalg.norm(pos)
    r = np.cross(f, [0, 0, 1.0]); r /= np.linalg.norm(r); d = np.cross(f, r)
    c2w = np.eye(4); c2w[:3, :3] = np.stack([r, d, f], 1); c2w[:3, 3] = pos
    return np.linalg.inv(c2w)

def sphere_depth(w2c, W, H, fx, fy, cx, cy, dev, radius=0.6):
    c2w = torch.tensor(np.linalg.inv(w2c), dtype=torch.float64, device=dev)
    ys, xs = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float64), torch.arange(W, device=dev, dtype=torch.float64), indexing="ij")
    dirs = torch.stack([(xs - cx) / fx, (ys - cy) / fy, torch.ones_like(xs)], -1)
    dw = dirs @ c2w[:3, :3].T; o = c2w[:3, 3]
    a = (dw * dw).sum(-1); b = 2 * (dw @ o); c = o @ o - radius ** 2
    disc = b * b - 4 * a * c
    t = torch.where(disc > 0, (-b - torch.sqrt(disc.clamp_min(0))) / (2 * a), torch.zeros_like(a))
    return t.float().contiguous()

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from gs2mesh_b200 import _lib
    from gs2mesh_b200.tsdf import TSDFVolume
    W, H, fx = 1600, 1200, 1400.0
    vol = TSDFVolume(2.0 / 512, 4 * 2.0 / 512, with_color=True, device=dev)
    rgb = torch.full((H, W, 3), 128, dtype=torch.uint8, device=dev)
    def fuse():
        vol.reset()
        for k in range(3):
            az = 2 * np.pi * (3 * rank + k) / (3 * world)
            w2c = look_at(1.6 * np.array([np.cos(az), np.sin(az), 0.4 * np.sin(2 * az)]))
            vol.integrate(sphere_depth(w2c, W, H, fx, fx, W / 2, H / 2, dev), rgb, W, H, fx, fx, W / 2, H / 2, w2c)
        vol.ensure_capacity()
    out = {"world": world, "merges": []}
    _lib.profile_collect()
    for it in range(5):
        fuse(); own = vol.num_bricks()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        _lib.profile_enable(True)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        t0 = time.perf_counter(); e0.record(); vol.reduce_across_ranks(dst=0); e1.record(); host = 1e3 * (time.perf_counter() - t0)
        torch.cuda.synchronize(); _lib.profile_enable(False)
        ph = {k: round(ms, 4) for k, (ms, n) in _lib.profile_collect().items() if k.startswith("merge_") and n}
        out["merges"].append({"own": own, "union": vol.num_bricks() if rank == 0 else None, "device_ms": round(e0.elapsed_time(e1), 4), "host_ms": round(host, 4), "phases": ph})
    # raw NCCL on the same bytes
    n = (out["merges"][-1]["union"] or 0)
    t = torch.tensor([n], device=dev); dist.broadcast(t, 0); n = int(t.item())
    a = torch.ones(n * 4096 * 2, device=dev); b = torch.ones(n * 4096 * 4, device=dev)  # 16^3 voxels per brick
    def one(fn, cold):
        if cold: torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize(); time.sleep(0.002)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    def both():
        dist.reduce(a, 0); dist.reduce(b, 0)
    def coalesced():
        with dist._coalescing_manager(device=dev, async_ops=False):
            dist.reduce(a, 0); dist.reduce(b, 0)
    for _ in range(3): both()
    out["raw_two_reduces_cold_ms"] = [round(one(both, True), 4) for _ in range(5)]
    out["raw_big_reduce_cold_ms"] = [round(one(lambda: dist.reduce(b, 0), True), 4) for _ in range(5)]
    def piped():
        for _ in range(10): both()
    out["raw_two_reduces_pipelined_ms"] = round(one(piped, True) / 10, 4)
    out["payload_MB"] = round((a.numel() + b.numel()) * 4 / 1e6, 1)
    if rank == 0: print(json.dumps(out))
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
