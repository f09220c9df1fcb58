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
