"""Manual diagnosis of slab sharding on one device: prints where a slab run departs from the
single-GPU run (sizes per slab, vertex / triangle multiset differences)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tropical-nerf.pytorch_b200"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)
from helpers import canonical_triangles, canonical_vertices, load_golden, native_net, oracle_net  # noqa: E402
from tropical import parallel  # noqa: E402


def rows_diff(a, b):
    sa = {tuple(r) for r in a.tolist()}
    sb = {tuple(r) for r in b.tolist()}
    return len(sa - sb), len(sb - sa)


for case in sys.argv[1:] or ["tiny_sphere_h8", "small_sphere", "small_torus"]:
    N = native_net(oracle_net(load_golden(case)))
    v1, _, t1, _, _ = N.subpoly(size=1.2, eps=1e-4, force=True).read()
    v1, t1 = v1.cpu().numpy(), t1.cpu().numpy()
    for slabs in (2, 3, 4, 5, 6, 8):
        if slabs > N.n_marks - 1:
            continue
        try:
            v, t, stats = parallel.subpoly_slabs_local(N, slabs)
        except Exception as e:  # noqa: BLE001
            print(case, slabs, "FAILED", repr(e)[:300])
            continue
        v, t = v.cpu().numpy(), t.cpu().numpy()
        ok_v = v.shape == v1.shape and np.array_equal(canonical_vertices(v), canonical_vertices(v1))
        ok_t = t.shape == t1.shape and np.array_equal(canonical_triangles(v, t), canonical_triangles(v1, t1))
        print(case, "slabs", slabs, "V", v.shape[0], "/", v1.shape[0], "T", t.shape[0], "/", t1.shape[0],
              "vertices", "OK" if ok_v else ("DIFF %d/%d" % rows_diff(canonical_vertices(v), canonical_vertices(v1))),
              "triangles", "OK" if ok_t else ("DIFF %d/%d" % rows_diff(canonical_triangles(v, t), canonical_triangles(v1, t1))),
              stats.get("slab_vertices"), "shared", stats.get("shared_vertices"), "near_plane", stats.get("near_plane"))
