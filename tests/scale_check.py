"""One-off scale check (not collected by pytest): a medium/large fitted network, CUDA path
against the oracle, every array bit-exact.  Minutes of CPU time for the oracle.

    python tests/scale_check.py medium_sphere
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tropical-nerf.pytorch_b200"), os.path.join(ROOT, "tests")]
import bench  # noqa: E402
from oracle import subpoly_ref as R  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "medium_sphere"
w = bench.load_workload(name)
N = bench.make_native(w)
t = time.time()
mesh = N.subpoly()
v, e, tri, f, p = [a.cpu().numpy() for a in mesh.read()]
print(f"{name}: device path {time.time() - t:.3f}s (first call) -> {mesh.sizes()}", flush=True)
P = bench.oracle_params(w)
t = time.time()
faces, vo, to, inter = R.subpoly(P, return_intermediate=True)
print(f"oracle {time.time() - t:.1f}s -> V {vo.shape[0]} T {to.shape[0]}", flush=True)
ok = (np.array_equal(v, vo) and np.array_equal(e, inter["surface_edges"]) and np.array_equal(tri, to)
      and np.array_equal(f, faces))
print("BIT-EXACT" if ok else "MISMATCH", flush=True)
sys.exit(0 if ok else 1)
