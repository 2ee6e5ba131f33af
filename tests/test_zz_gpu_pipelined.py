"""Stage-class pipelining (runs last in the GPU suite): asynchronous stereo pairs on rotating stream sets and buffer sets."""
import pytest

from tests.test_gpu_pipeline import Args, small_scene  # noqa: F401  (fixture + reference flag defaults)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("in_flight", [1, 2, 3, 4])
def test_pipelined_pairs_equal_synchronous_renders(gsb_lib, cuda_device, small_scene, in_flight):
    """`render_image_pair(..., wait=False)` with `pairs_in_flight` stream sets, consumed the way bench.py's e2e loop
    does (pair n after the next `pairs_in_flight` pairs were enqueued), returns exactly what synchronous calls return:
    buffer rotation, stream gates and completion handles hide no race."""
    import torch

    from gs2mesh_b200.renderer import Renderer

    cloud, rigs, baseline = small_scene
    r = Renderer.from_scene(rigs, baseline, cloud, output_dir_root=None, args=Args(), device=str(cuda_device))
    r.pairs_in_flight = in_flight
    r.prepare_renderer()
    views = [0, 1, 2, 3, 0, 1, 2, 3, 1, 3, 2]  # more calls than buffer sets, every view several times
    keys = ("host_left_u8", "host_right_u8", "depth", "final_T", "left", "right")
    want = {}
    for v in sorted(set(views)):
        out = r.render_image_pair(v, to_host=True)  # synchronous
        want[v] = {k: out[k].detach().cpu().clone() for k in keys}
    lag = max(1, in_flight)
    pending = [r.render_image_pair(v, to_host=True, wait=False) for v in views[:lag]]
    for n, v in enumerate(views):
        out = pending.pop(0)
        if n + lag < len(views):
            pending.append(r.render_image_pair(views[n + lag], to_host=True, wait=False))
        out["ready"].synchronize()
        for k in keys:
            assert torch.equal(out[k].detach().cpu(), want[v][k]), (in_flight, n, v, k)
    torch.cuda.synchronize()
    r.check_status(sorted(set(views)))
