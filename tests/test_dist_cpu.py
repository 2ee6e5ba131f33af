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
    """{(bx,by,bz): (tsdf, weight)[4096,2]} of the units the CPU oracle opens for these views."""
    views = _views()
    vol = orc.OracleTSDFVolume(VL, TRUNC, with_color=False)
    for i in idxs:
        vol.integrate(views[i][0], None, W, H, FX, FY, CX, CY, views[i][1])
    out = {}
    for i, k in enumerate(vol.unit_indices()):
        t, w, _ = vol.unit_data(i)
        out[tuple(int(v) for v in k)] = np.stack([t, w], -1)
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gs2mesh_b200.tsdf import merge_units_protocol, shard_views
        from oracle import oracle as orc

        mine = shard_views(N_VIEWS, rank, world)
        units = _fuse(orc, mine)
        merged_all = merge_units_protocol(units, dst=None)   # all-reduce: every rank ends with the union
        merged_root = merge_units_protocol(units, dst=1)     # reduce to rank 1: rank 0's volume is untouched
        q.put((rank, mine, sorted(units), merged_all, merged_root))
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
    by_rank = {r: rest for r, *rest in results}
    assert sorted(by_rank[0][0] + by_rank[1][0]) == list(range(N_VIEWS))
    own0, own1 = set(by_rank[0][1]), set(by_rank[1][1])
    assert own0 != own1 and own0 | own1 == set(seq), "the two shards open different unit sets whose union is the sequential one"

    def same(merged):
        assert set(merged) == set(seq)
        for k in seq:
            np.testing.assert_array_equal(merged[k][:, 1], seq[k][:, 1])
            np.testing.assert_allclose(merged[k][:, 0], seq[k][:, 0], atol=1e-6)

    for r in (0, 1):  # all-reduce: every rank holds the merged volume
        same(by_rank[r][2])
    same(by_rank[1][3])  # reduce to rank 1
    assert set(by_rank[0][3]) == own0  # ... leaves rank 0 with what it had
    own_units = _fuse(oracle, by_rank[0][0])
    for k in own0:
        np.testing.assert_array_equal(by_rank[0][3][k], own_units[k])
