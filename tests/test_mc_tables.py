"""Internal-consistency checks of the marching-cubes case table (include/gsb_mc_tables.h)."""
import itertools
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_tables():
    text = open(os.path.join(ROOT, "include", "gsb_mc_tables.h")).read()

    def grab(name):
        m = re.search(name + r"\[[^\]]*\](?:\[[^\]]*\])?\s*=\s*\{(.*?)\};", text, re.S)
        return [int(v) for v in re.findall(r"-?\d+", m.group(1))]

    shift = np.array(grab("kMcShift")).reshape(8, 3)
    edge_shift = np.array(grab("kMcEdgeShift")).reshape(12, 4)
    e2v = np.array(grab("kMcEdgeToVert")).reshape(12, 2)
    tri = np.array(grab("kMcTriTable")).reshape(256, 16)
    return shift, edge_shift, e2v, tri


SHIFT, EDGE_SHIFT, E2V, TRI = load_tables()


def test_geometry_tables_are_consistent():
    for e in range(12):
        a, b = E2V[e]
        axis = EDGE_SHIFT[e][3]
        lo = np.minimum(SHIFT[a], SHIFT[b])
        assert (lo == EDGE_SHIFT[e][:3]).all()  # edge origin = its lower corner
        d = np.abs(SHIFT[a] - SHIFT[b])
        assert d.sum() == 1 and d[axis] == 1  # unit edge along `axis`


def test_rows_are_well_formed_and_use_exactly_the_sign_changing_edges():
    for c in range(256):
        row = TRI[c]
        n = int((row >= 0).sum())
        assert n % 3 == 0 and (row[:n] >= 0).all() and (row[n:] == -1).all() and n <= 15
        used = set(row[:n].tolist())
        crossing = {e for e in range(12) if ((c >> E2V[e][0]) & 1) != ((c >> E2V[e][1]) & 1)}
        assert used == crossing, (c, sorted(used), sorted(crossing))
        for t in range(0, n, 3):
            assert len(set(row[t:t + 3].tolist())) == 3


def _mesh(field):
    """Triangle list over welded edge-vertices for a sign field (Open3D's winding: edges 0,2,1 of a triple)."""
    n = field.shape[0]
    tris = []
    for x, y, z in itertools.product(range(n - 1), repeat=3):
        c = 0
        for i in range(8):
            if field[x + SHIFT[i][0], y + SHIFT[i][1], z + SHIFT[i][2]] < 0:
                c |= 1 << i
        row = TRI[c]
        for t in range(0, 15, 3):
            if row[t] < 0:
                break
            keys = []
            for e in (row[t], row[t + 2], row[t + 1]):
                es = EDGE_SHIFT[e]
                keys.append((x + es[0], y + es[1], z + es[2], es[3]))
            tris.append(tuple(keys))
    return tris


def test_surfaces_are_closed_and_consistently_oriented():
    """Inside the grid every mesh edge is shared by exactly two triangles that traverse it in opposite
    directions (the negative region is padded away from the boundary, so the surface must be closed)."""
    rng = np.random.default_rng(0)
    for trial in range(12):
        n = 7
        f = np.ones((n, n, n))
        f[1:-1, 1:-1, 1:-1] = rng.normal(size=(n - 2,) * 3) + (0.3 if trial % 2 else -0.3)
        tris = _mesh(f)
        assert tris
        directed = {}
        for a, b, c in tris:
            for u, v in ((a, b), (b, c), (c, a)):
                directed[(u, v)] = directed.get((u, v), 0) + 1
        for (u, v), cnt in directed.items():
            assert cnt == 1, "an edge is traversed twice in the same direction"
            assert directed.get((v, u), 0) == 1, "open edge: the surface has a hole"


def test_complement_cases_use_the_same_edges():
    for c in range(256):
        a = set(TRI[c][TRI[c] >= 0].tolist())
        b = set(TRI[255 - c][TRI[255 - c] >= 0].tolist())
        assert a == b
