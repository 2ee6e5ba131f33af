"""Shared checker: a gs2mesh_b200 TSDFVolume (unbounded brick pool on the GPU) against the CPU oracle's volume
(Open3D ScalableTSDFVolume restatement), unit by unit, keyed by the integer lattice index of the 16^3 block."""
import numpy as np


def assert_units_equal(vol, ovol, *, color=True, color_atol=1e-3, color_rtol=1e-5):
    """Same SET of opened units (Open3D's block discovery, T1), bit-identical (tsdf, weight) in every one of them (T2) and
    colours within fp32-vs-fp64 running-mean rounding.  Returns the number of units."""
    got = vol.export_units()
    idx = ovol.unit_indices()
    want_keys = {tuple(int(v) for v in k) for k in idx}
    assert set(got) == want_keys, (f"unit sets differ: {len(got)} on the GPU, {len(want_keys)} in the oracle; "
                                   f"only GPU {sorted(set(got) - want_keys)[:5]}, only oracle {sorted(want_keys - set(got))[:5]}")
    for i, k in enumerate(idx):
        t, w, c = ovol.unit_data(i)
        tw, col = got[tuple(int(v) for v in k)]
        np.testing.assert_array_equal(tw[:, 1], w, err_msg=f"weights of unit {tuple(k)}")
        np.testing.assert_array_equal(tw[:, 0], t, err_msg=f"tsdf of unit {tuple(k)}")
        if color and c is not None and col is not None:
            m = w > 0
            np.testing.assert_allclose(col[m, :3], c[m], rtol=color_rtol, atol=color_atol, err_msg=f"colour of unit {tuple(k)}")
    return len(idx)
