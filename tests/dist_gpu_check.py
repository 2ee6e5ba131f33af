"""Multi-GPU check (run under torchrun on a box with >= 2 GPUs; not collected by pytest):
view-sharded integration + ONE NCCL sum-reduce of the brick volume == sequential single-GPU fusion.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gs2mesh_b200 import scene
    from gs2mesh_b200.renderer import Renderer
    from gs2mesh_b200.tsdf import TSDF, shard_views

    class A:
        GS_white_background = False
        TSDF_voxel = 8
        TSDF_sdf_trunc = 0.06
        TSDF_scale = 1.0
        TSDF_min_depth_baselines = 4
        TSDF_max_depth_baselines = 20

    W, H, NV = 320, 240, 6
    cloud = scene.make_gaussians(8000, seed=3)
    rigs, baseline = scene.make_stereo_cameras(NV, W, H)
    r = Renderer.from_scene(rigs, baseline, cloud, args=A(), device=f"cuda:{local}")
    r.prepare_renderer()

    def fuse(views):
        st = TSDF(r, None, A(), "d", window_resolution=128, device=f"cuda:{local}")
        st.volume = st._make_volume()
        for i in views:
            out = r.render_image_pair(i, to_host=False)
            st.integrate(out["depth"], out["left_u8"], rigs[i]["left"], final_T=out["final_T"])
        return st.volume

    mine = shard_views(NV, rank, world)
    vol = fuse(mine)
    vol.reduce_across_ranks(dst=None, chunk_bytes=1 << 20)  # sparse merge (touched bricks only), many chunks
    dense = fuse(mine)
    dense.reduce_across_ranks(dst=None, chunk_bytes=1 << 22, sparse=False)  # whole-volume merge
    torch.cuda.synchronize()
    # same sums, but NCCL may associate them differently for the two buffer shapes: equal to rounding, same weights
    assert torch.equal(vol.tsdf_weight.view(-1, 2)[:, 1], dense.tsdf_weight.view(-1, 2)[:, 1]), "sparse != dense merge (weights)"
    assert float((vol.tsdf_weight - dense.tsdf_weight).abs().max()) <= 1e-6, "sparse != dense merge (tsdf)"
    assert float((vol.color - dense.color).abs().max()) <= 1e-3, "sparse != dense merge (colour)"
    del dense
    seq = fuse(range(NV))
    torch.cuda.synchronize()
    a = vol.bricks().cpu().numpy()
    b = seq.bricks().cpu().numpy()
    assert np.array_equal(a[..., 1], b[..., 1]), "merged weights differ from sequential fusion"
    err = float(np.abs(a[..., 0] - b[..., 0]).max())
    assert err <= 2e-6, err
    ca = vol.color.view(-1, 4).cpu().numpy()
    cb = seq.color.view(-1, 4).cpu().numpy()
    cerr = float(np.abs(ca - cb).max())
    assert cerr <= 1e-3, cerr
    dist.barrier()
    if rank == 0:
        print(f"dist_gpu_check ok: world={world} views/rank={[len(shard_views(NV, q, world)) for q in range(world)]} "
              f"touched voxels={int((b[..., 1] > 0).sum())} max|dtsdf|={err:.2e} max|dcolor|={cerr:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
