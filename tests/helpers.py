"""Shared test helpers: golden fixtures, oracle networks, native (CUDA) networks."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
CASES = ["tiny_sphere_h8", "small_sphere", "small_torus"]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


def oracle_net(g):
    from oracle.trinet import NetParams
    return NetParams.from_npz_dict({k[4:]: v for k, v in g.items() if k.startswith("net_")})


def native_net(P):
    """CUDA-resident copy of an oracle NetParams (plain arrays only cross over)."""
    from tropical._native import NativeNet
    return NativeNet(P.levels, P.n_feat, P.log2_T, P.n_min, P.per_level_scale, P.num_layers,
                     P.num_hidden, P.table, P.mlp, P.marks, P.eps, P.scale)


def random_net(seed, levels=4, n_min=2, n_max=32, log2_T=19, num_layers=3, num_hidden=16,
               table_amp=0.5, marks=None):
    """Random-weight oracle network (no fitting): enough for evaluation parity."""
    from oracle.trinet import NetParams, grid_layout
    rng = np.random.default_rng(seed)
    b = float(np.exp2(np.log2(n_max / n_min) / (levels - 1))) if levels > 1 else 1.0
    total = grid_layout(levels, log2_T, n_min, b)[4]
    table = (rng.random(total * 2, dtype=np.float32) * 2 - 1) * np.float32(table_amp)
    nodes = [levels * 2] + [num_hidden] * (num_layers - 1) + [2]
    ws = [(rng.standard_normal((nodes[i + 1], nodes[i])) / np.sqrt(nodes[i])).astype(np.float32)
          for i in range(num_layers)]
    bs = [(rng.standard_normal(nodes[i + 1]) * 0.1).astype(np.float32) for i in range(num_layers)]
    if marks is None:
        marks = np.linspace(0, 1, 9, dtype=np.float32)
    return NetParams(levels, 2, log2_T, n_min, b, num_layers, num_hidden, table, ws, bs, marks)


def canonical_polygons(rows):
    from oracle.subpoly_ref import canonical_polygons as cp
    return cp(rows)


def canonical_triangles(vertices, triangles):
    """Mesh as a sorted multiset of triangles given by their corner POSITION bits (cyclic order
    kept, rotation normalised): independent of vertex numbering, exact on positions."""
    v = np.ascontiguousarray(np.asarray(vertices, np.float32)).view(np.int32).astype(np.int64)
    t = np.asarray(triangles, np.int64)
    if t.shape[0] == 0:
        return np.zeros((0, 9), np.int64)
    p = v[t]                                   # [T,3,3]
    key = (p[..., 0] << 42) ^ (p[..., 1] << 21) ^ p[..., 2]
    first = np.argmin(key, axis=1)
    idx = (first[:, None] + np.arange(3)[None, :]) % 3
    p = np.take_along_axis(p, idx[:, :, None], axis=1).reshape(-1, 9)
    order = np.lexsort(p.T[::-1])
    return p[order]


def canonical_vertices(vertices):
    v = np.ascontiguousarray(np.asarray(vertices, np.float32)).view(np.int32).reshape(-1, 3)
    return v[np.lexsort(v.T[::-1])]
