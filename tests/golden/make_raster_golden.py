"""Generates tests/golden/raster_ref_*.npz: outputs of the UNMODIFIED reference rasterizer
(oracle/_ref/libref_dgr.so, built by oracle/build_ref.sh from /root/reference) run on a B200.
Must run on a GPU box:  gpurun -- python tests/golden/make_raster_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/.  Inputs are stored with the outputs so the
CPU-only test (tests/test_oracle_raster.py) needs nothing else."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gs2mesh_b200 import camera as cam  # noqa: E402
from gs2mesh_b200 import scene  # noqa: E402
from oracle import oracle as orc  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(out_dir, exist_ok=True)
dev = torch.device("cuda", 0)
cases = [("a", 1500, 160, 120, 11, 0, 3, (0, 0, 0), 1.0), ("b", 900, 200, 104, 12, 1, 2, (1, 1, 1), 0.7)]
for name, n, W, H, seed, view, deg, bg, smod in cases:
    g = scene.make_gaussians(n, seed=seed)
    rigs, _ = scene.make_stereo_cameras(4, W, H)
    vt = cam.view_transforms_from_camera(rigs[view]["right" if name == "b" else "left"])
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    res = orc.ref_forward_torch(t(g.xyz), t(g.opacity).reshape(-1), t(vt.world_view), t(vt.full_proj), t(vt.cam_center), W, H,
                                vt.tan_fovx, vt.tan_fovy, t(np.array(bg, np.float32)), shs=t(g.features), scales=t(g.scaling),
                                rotations=t(g.rotation), sh_degree=deg, scale_modifier=smod)
    np.savez_compressed(os.path.join(out_dir, f"raster_ref_{name}.npz"), means3D=g.xyz, opacities=g.opacity, shs=g.features,
                        scales=g.scaling, rotations=g.rotation, view=vt.world_view, proj=vt.full_proj, campos=vt.cam_center,
                        W=W, H=H, tan_fovx=vt.tan_fovx, tan_fovy=vt.tan_fovy, bg=np.array(bg, np.float32), sh_degree=deg,
                        scale_modifier=smod, color=res["color"].cpu().numpy(), radii=res["radii"].cpu().numpy(),
                        num_rendered=res["num_rendered"], final_T=res["final_T"].cpu().numpy())
    print("wrote", name, res["num_rendered"])
