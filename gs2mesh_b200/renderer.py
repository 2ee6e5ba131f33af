"""Stereo-pair renderer stage: B200-native replacement for
gs2mesh_utils/renderer_utils.py `Renderer` (reference), same class surface:

    Renderer(base_dir, colmap_dir, output_dir_root, args, dataset=, splatting=, experiment_name=, device=)
    .prepare_renderer()  .render_image_pair(i, visualize=False)  .render_folder_name(i)  len()
    .cameras  .left_cameras  .baseline  .output_dir_root
plus `render(i)` (the name BASELINE.json's north_star uses) and `Renderer.from_scene(...)` for
in-memory scenes.

What changes relative to the reference loop (renderer_utils.py:363-395), by design:
* activations / SH concatenation are done once in prepare_renderer(), not per view
  (gaussian_renderer/__init__.py:53-80 recomputes them for every camera);
* all camera matrices are uploaded once as a device table; no per-view `Camera` object, no
  dummy `torch.rand(3,h,w)` image upload (renderer_utils.py:386, cameras.py:39);
* the float->uint8 conversion (x255, saturate, HWC) runs on the GPU, so the D2H copy is 3 bytes
  per pixel instead of 12 (renderer_utils.py:389);
* the left view's expected depth (sum z*alpha*T) and transmittance are produced in the same
  blend loop and stay on the device for the TSDF stage.
PNG files are still written when `write_images` is on, so downstream reference stages (Stereo,
Masker) keep working.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from . import camera as cam
from . import rasterizer as rast


DEFAULT_PAIRS_IN_FLIGHT = 2


class _PairReady:
    """Completion handle of an asynchronously rendered pair (render_image_pair(..., wait=False))."""

    def __init__(self, events):
        self._events = events

    def synchronize(self):
        for e in self._events:
            e.synchronize()


class Renderer:
    def __init__(self, base_dir=None, colmap_dir=None, output_dir_root=None, args=None, dataset="custom", splatting="custom",
                 experiment_name=None, device="cuda", *, cameras=None, baseline=None, gaussians=None):
        self.args = args
        self.base_dir = base_dir
        self.colmap_dir = colmap_dir
        self.output_dir_root = output_dir_root
        self.device = device
        self.dataset = dataset
        self.render_name = getattr(args, "colmap_name", None)
        self.white_background = bool(getattr(args, "GS_white_background", False))
        self.splatting_iteration = getattr(args, "GS_iterations", 30000)
        self._cloud = gaussians
        if cameras is None:
            # reference path: COLMAP text model + trained point_cloud.ply on disk (renderer_utils.py:127-216)
            from .io import load_reference_scene

            cameras, baseline, self._cloud, self.poses, self.sorted_camera_indices = load_reference_scene(
                base_dir, colmap_dir, args, splatting)
        self.cameras = cameras
        self.baseline = baseline
        self.left_cameras = [c["left"] for c in self.cameras]
        print(f"num views: {len(self.cameras)}")
        print(f"baseline: {self.baseline}")
        if output_dir_root is not None and getattr(args, "renderer_save_json", False):
            self.save_camera_data()
        self.write_images = output_dir_root is not None
        # blend with ex2.approx instead of full-precision expf (GSB_RASTER_FAST_EXP): ~2e-7 relative on
        # alpha, parity-tested against the reference rasterizer inside the 1e-4 budget
        self.fast_exp = True
        # render the left and the right eye on two side streams (each with its own scratch) so that the
        # small latency-bound binning kernels of one eye overlap the other eye's blend kernel
        self.overlap_eyes = True
        # how many stereo pairs may be in flight at once (stream sets); GSB_PAIRS_IN_FLIGHT overrides (A/B switch)
        self.pairs_in_flight = int(os.environ.get("GSB_PAIRS_IN_FLIGHT", DEFAULT_PAIRS_IN_FLIGHT))
        self.keep_frames = False
        self._frames = {}
        self._ready = False

    @classmethod
    def from_scene(cls, cameras, baseline, gaussians, output_dir_root=None, args=None, device="cuda"):
        return cls(None, None, output_dir_root, args, device=device, cameras=cameras, baseline=baseline, gaussians=gaussians)

    def __len__(self):
        return len(self.cameras)

    def render_folder_name(self, render_number):
        return os.path.join(self.output_dir_root or ".", f"{render_number:03}")

    def save_camera_data(self):
        """renderer_utils.py:298-314"""
        import copy
        import json

        os.makedirs(self.output_dir_root, exist_ok=True)
        out = copy.deepcopy(self.cameras)
        for pair in out:
            for side in ("left", "right"):
                pair[side]["intrinsic"] = np.asarray(pair[side]["intrinsic"]).tolist()
                pair[side]["extrinsic"] = np.asarray(pair[side]["extrinsic"]).tolist()
        with open(os.path.join(self.output_dir_root, "camera_data.json"), "w") as f:
            json.dump(out, f, indent=4)

    # ------------------------------------------------------------------------------------------
    def prepare_renderer(self):
        """Upload the Gaussian cloud and every camera once (reference: renderer_utils.py:316-361)."""
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError("Renderer needs a CUDA device (gs2mesh_b200 has no CPU path)")
        g = self._cloud
        if g is None:
            raise RuntimeError("Renderer has no Gaussian cloud")
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
        self.means3D = up(g.xyz)
        self.shs = up(g.features)  # [P,16,3] == cat(features_dc, features_rest) (gaussian_model.py:107-111)
        self.opacity = up(g.opacity).reshape(-1)
        self.scales = up(g.scaling)
        self.rotations = up(g.rotation)
        self.sh_degree = int(g.sh_degree)
        bg = [1, 1, 1] if self.white_background else [0, 0, 0]  # renderer_utils.py:359
        self.background = torch.tensor(bg, dtype=torch.float32, device=dev)
        # camera table: [n_views, 2 (left,right), 36] floats
        recs = np.zeros((len(self.cameras), 2, cam.CAMERA_RECORD_FLOATS), dtype=np.float32)
        self._views = []
        for i, pair in enumerate(self.cameras):
            vts = []
            for s, side in enumerate(("left", "right")):  # dict order of the reference: left then right (:378)
                vt = cam.view_transforms_from_camera(pair[side])
                recs[i, s] = vt.packed()
                vts.append(vt)
            self._views.append(vts)
        self._camera_table = torch.as_tensor(recs).to(dev)
        self._bufs = {}
        # per-(view, side) rasterizer status words, written by the GPU straight into pinned host memory
        self._status = torch.zeros(len(self.cameras), 2, 4, dtype=torch.int64).pin_memory()
        self._min_instances = 0
        self._ready = True
        self.calibrate_capacity()

    def calibrate_capacity(self, sample_views=3, slack=1.5):
        """Size the binning scratch once from a few synchronous renders so that the hot loop can run
        without ever waiting for the GPU (the reference re-sizes -- and blocks -- inside every frame,
        rasterizer_impl.cu:281-285)."""
        n = len(self.cameras)
        worst = 0
        for i in sorted(set(int(round(k * (n - 1) / max(sample_views - 1, 1))) for k in range(sample_views))):
            for s in range(2):
                out = self.render_view(i, s, want_counts=True)
                worst = max(worst, int(out["counts"][0].item()))
        self._min_instances = int(worst * slack) + 4096

    def check_status(self, views=None):
        """After a synchronisation: raise if any asynchronously rendered frame overflowed the scratch."""
        st = self._status if views is None else self._status[list(views)]
        bad = (st[..., 2] != 0).nonzero()
        if len(bad):
            need = int(st[..., 0].max().item())
            self._min_instances = max(self._min_instances, int(need * 1.5))
            raise RuntimeError(f"{len(bad)} frame(s) needed more binning scratch than calibrated ({need} instances); "
                               "capacity has been raised, render them again")

    def _buffers(self, w, h, parity=0):
        key = (w, h, parity)
        if key not in self._bufs:
            dev = self._camera_table.device
            f32 = dict(dtype=torch.float32, device=dev)
            self._bufs[key] = dict(
                color=[torch.empty(3, h, w, **f32) for _ in range(2)],
                u8=[torch.empty(h, w, 3, dtype=torch.uint8, device=dev) for _ in range(2)],
                depth=torch.empty(h, w, **f32), final_T=torch.empty(h, w, **f32),
                host_u8=[torch.empty(h, w, 3, dtype=torch.uint8).pin_memory() for _ in range(2)],
                host_depth=torch.empty(h, w, dtype=torch.float32).pin_memory(),
            )
        return self._bufs[key]

    def render_view(self, camera_number, side, *, want_depth=False, out_color=None, out_depth=None, out_final_T=None,
                    flags=None, want_counts=False, async_mode=False):
        """One forward rasterization of view `camera_number`, side 0 (left) / 1 (right)."""
        if not self._ready:
            raise RuntimeError("call prepare_renderer() first")
        if flags is None:
            flags = rast.DEFAULT_FLAGS | (rast._lib.RASTER_FAST_EXP if self.fast_exp else 0)
        vt = self._views[camera_number][side]
        rec = self._camera_table[camera_number, side]
        return rast.rasterize_forward(
            means3D=self.means3D, opacities=self.opacity, viewmatrix=rec[0:16], projmatrix=rec[16:32], campos=rec[32:35],
            bg=self.background, width=vt.width, height=vt.height, tan_fovx=vt.tan_fovx, tan_fovy=vt.tan_fovy, shs=self.shs,
            scales=self.scales, rotations=self.rotations, sh_degree=self.sh_degree, scale_modifier=1.0, flags=flags,
            want_depth=want_depth, want_final_T=want_depth, want_radii=False, want_counts=want_counts,
            out_color=out_color, out_depth=out_depth, out_final_T=out_final_T, async_mode=async_mode,
            counts_out=self._status[camera_number, side] if async_mode else None, min_instances=self._min_instances)

    def render_image_pair(self, camera_number, visualize=False, *, to_host: Optional[bool] = None, wait: bool = True):
        """Render the stereo-aligned left/right pair of view `camera_number`
        (reference: renderer_utils.py:363-395).  Returns a dict of DEVICE tensors
        (left/right float CHW, left_u8/right_u8 HWC, depth = left sum(z*alpha*T), final_T) and, when
        `to_host` (default: whenever PNGs are written), pinned host copies `host_left_u8`,
        `host_right_u8`.  With `wait=False` the call only enqueues: `result["ready"]` is a CUDA event to
        synchronise on before touching the host copies (lets the caller overlap the next pair's rendering
        with the host-side consumption of this one)."""
        with torch.no_grad():
            vt = self._views[camera_number][0]
            dev = self._camera_table.device
            main = torch.cuda.current_stream(dev)
            if self.overlap_eyes:
                # outputs rotate over pairs_in_flight + 2 buffer sets: the tensors returned by call n stay valid until
                # call n + pairs_in_flight + 1 returns, which covers a caller that consumes pair n only after
                # enqueuing the next pairs_in_flight pairs (wait=False)
                self._call_index = getattr(self, "_call_index", -1) + 1
                depth = max(1, int(self.pairs_in_flight))
                nslots = depth + 2
                slot = self._call_index % nslots
                if not hasattr(self, "_side_streams") or len(self._side_streams) != depth:
                    # `pairs_in_flight` sets of (left, right) streams, used round-robin: consecutive pairs run on
                    # different stream sets (each stream has its own scratch), so pair n+1's latency-bound binning
                    # kernels fill the gaps of pair n's blend kernels
                    self._side_streams = [[torch.cuda.Stream(dev), torch.cuda.Stream(dev)] for _ in range(depth)]
                if not hasattr(self, "_entry_events") or len(self._entry_events) != nslots:
                    main.synchronize()  # (re)configuration only: nothing may still use the old rotation
                    self._call_index = 0
                    slot = 0
                    self._entry_events = [None] * nslots
                    self._slot_done = [[] for _ in range(nslots)]
                b = self._buffers(vt.width, vt.height, slot)
                entry = torch.cuda.Event()
                entry.record(main)
                # everything that read this buffer set (handed out nslots calls ago) was enqueued on the caller's
                # stream before the PREVIOUS call was entered
                gate = self._entry_events[(self._call_index - 1) % nslots]
                self._entry_events[slot] = entry
                streams = self._side_streams[self._call_index % depth]
                for st in streams:
                    if gate is not None:
                        st.wait_event(gate)
                    else:
                        st.wait_stream(main)
                    for done in self._slot_done[slot]:  # the renders that last wrote this buffer set (another stream set)
                        st.wait_event(done)
            else:
                b = self._buffers(vt.width, vt.height)
                streams = [main, main]
            for s in range(2):
                with torch.cuda.stream(streams[s]):
                    self.render_view(camera_number, s, want_depth=(s == 0), out_color=b["color"][s],
                                     out_depth=b["depth"] if s == 0 else None, out_final_T=b["final_T"] if s == 0 else None,
                                     async_mode=True)
                    rast.image_to_u8(b["color"][s], out=b["u8"][s])
            result = dict(left=b["color"][0], right=b["color"][1], left_u8=b["u8"][0], right_u8=b["u8"][1],
                          depth=b["depth"], final_T=b["final_T"])
            to_host = self.write_images if to_host is None else to_host
            if to_host:
                for s in range(2):
                    with torch.cuda.stream(streams[s]):
                        b["host_u8"][s].copy_(b["u8"][s], non_blocking=True)
            if self.overlap_eyes:
                self._slot_done[slot] = [st.record_event() for st in streams]
            asynchronous = to_host and not wait and not self.write_images and self.overlap_eyes
            if self.overlap_eyes and not asynchronous:
                for st in streams:
                    main.wait_stream(st)  # device-side dependency only: later work on the caller's stream sees both eyes
            if to_host:
                result["host_left_u8"], result["host_right_u8"] = b["host_u8"]
                if asynchronous:
                    result["ready"] = _PairReady([st.record_event() for st in streams])
                else:
                    main.synchronize()
                    self.check_status([camera_number])
                    result["ready"] = _PairReady([])
            if self.write_images:
                import cv2

                out_dir = self.render_folder_name(camera_number)
                os.makedirs(out_dir, exist_ok=True)
                for s, name in enumerate(("left", "right")):
                    cv2.imwrite(os.path.join(out_dir, f"{name}.png"), cv2.cvtColor(b["host_u8"][s].numpy(), cv2.COLOR_RGB2BGR))
            if self.keep_frames:
                self._frames[camera_number] = dict(left_u8=b["u8"][0].clone(), depth=self.expected_depth(b["depth"], b["final_T"]))
            return result

    # north_star alias
    def render(self, camera_number, **kw):
        return self.render_image_pair(camera_number, **kw)

    @staticmethod
    def expected_depth(depth_sum, final_T, alpha_min=0.5):
        """Expected depth D/alpha where alpha = 1 - T > alpha_min, else 0."""
        alpha = 1.0 - final_T
        return torch.where(alpha > alpha_min, depth_sum / alpha.clamp_min(1e-12), torch.zeros_like(depth_sum))

    def get_frame(self, camera_number):
        return self._frames.get(camera_number)
