"""Multi-GPU path, host side, on CPU with gloo (world_size 2): round-robin view sharding and the
single sum-form reduce of the volume reproduce the sequential fusion."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

W, H, FX, FY, CX, CY = 64, 48, 60.0, 60.0, 32.0, 24.0
VL, TRUNC = 1.0 / 64, 0.05
B0, NB = (-4, -4, -2), (8, 8, 8)
N_VIEWS = 5


def _views():
    out = []
    for k in range(N_VIEWS):
        a = np.radians(-10 + 5 * k)
        e = np.eye(4)
        e[0, 0], e[0, 2], e[2, 0], e[2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
        e[2, 3] = 0.05 * k
        out.append((np.full((H, W), 1.0 + 0.01 * k, np.float32), e))
    return out


def _fuse(orc, idxs):
    views = _views()
    vol = orc.OracleTSDFVolume(VL, TRUNC, with_color=False)
    for i in idxs:
        vol.integrate(views[i][0], None, W, H, FX, FY, CX, CY, views[i][1])
    tw, _, outside = vol.export_bricks(B0, NB)
    assert outside == 0
    return tw


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gs2mesh_b200.tsdf import reduce_sum_chunked, shard_views
        from oracle import oracle as orc

        mine = shard_views(N_VIEWS, rank, world)
        tw = _fuse(orc, mine)
        sums = torch.from_numpy(np.stack([tw[..., 0] * tw[..., 1], tw[..., 1]], -1).copy())  # (mean,w) -> (sum,w)
        reduce_sum_chunked([sums, None], chunk_bytes=1 << 16)  # many small chunks on purpose
        s = sums.numpy()
        mean = np.where(s[..., 1] > 0, s[..., 0] / np.maximum(s[..., 1], 1e-30), 0).astype(np.float32)
        q.put((rank, mine, mean, s[..., 1].copy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_views_round_robin():
    from gs2mesh_b200.tsdf import shard_views

    assert shard_views(7, 0, 2) == [0, 2, 4, 6] and shard_views(7, 1, 2) == [1, 3, 5]
    got = sorted(v for r in range(8) for v in shard_views(500, r, 8))
    assert got == list(range(500))
    assert max(len(shard_views(500, r, 8)) for r in range(8)) - min(len(shard_views(500, r, 8)) for r in range(8)) <= 1
    assert shard_views(3, 5, 8) == []


@pytest.mark.timeout(300)
def test_two_rank_gloo_merge_equals_sequential(oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seq = _fuse(oracle, range(N_VIEWS))
    shards = {r: m for r, m, _, _ in results}
    assert sorted(shards[0] + shards[1]) == list(range(N_VIEWS))
    for _, _, mean, weight in results:  # all-reduce: every rank holds the merged volume
        np.testing.assert_array_equal(weight, seq[..., 1])
        np.testing.assert_allclose(mean, seq[..., 0], atol=1e-6)
