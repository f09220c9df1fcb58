"""compute-sanitizer target: one small extraction through each step route (persistent kernel, device-driven stream,
curve path) plus the stage-level entry points.  Usage: compute-sanitizer --tool memcheck python profiles/r2_memcheck.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tropical-nerf.pytorch_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from helpers import load_golden, native_net, oracle_net  # noqa: E402
from tropical._native import lib  # noqa: E402

torch.cuda.set_device(0)
for case in ("tiny_sphere_h8", "small_torus"):
    N = native_net(oracle_net(load_golden(case)))
    print(case, "persistent", N.subpoly().sizes())
    before = lib().tnb_set_fused_max_items(0)
    print(case, "stream", N.subpoly().sizes())
    lib().tnb_set_fused_max_items(before)
    print(case, "curve", N.subpoly(force=False).sizes())
    c = N.skeleton(17)   # chunked skeleton: overlap duplicates, long partner lists, long face rows
    H = N.num_hidden
    c.steps([(l, h) for l in range(N.num_layers - 1) for h in range(H)] + [(N.num_layers - 2, H)])
    print(case, "unit 17", c.extract_mesh().sizes())
# the gradient-descent repair: stage-level call on the golden inputs, and inside the step kernels (persistent and
# multi-launch) at eps = 1e-5, where it runs its 500 steps and the extraction ends (tests/test_repair.py)
import numpy as np  # noqa: E402
from tropical import _native  # noqa: E402
g = load_golden("gd_stage")
N = native_net(oracle_net(load_golden("small_torus")))
for name in ("nat", "syn0"):
    eps = np.float32(g[f"{name}_eps"])
    gd = ~g[f"{name}_gg"] & ((np.abs(g[f"{name}_d_new"]) > eps).sum(-1) > 0)
    x, d, n, ok = N.gradient_descent(torch.from_numpy(g[f"{name}_e"][gd]).cuda(), torch.from_numpy(g[f"{name}_ints"][gd]).cuda(),
                                     torch.from_numpy(g[f"{name}_plane"][gd]).cuda(), int(g[f"{name}_idx"]), float(eps))
    print("repair", name, n, ok)
H = N.num_hidden
steps = [(l, h) for l in range(N.num_layers - 1) for h in range(H)] + [(N.num_layers - 2, H)]
for fused in (True, False):
    c = N.skeleton(128)
    try:
        for (l, h) in steps:
            if fused:
                c.steps([(l, h)], eps=1e-5, force=False)
            else:
                c.step(l, h, eps=1e-5, force=False)
            c.num_vertices
    except _native.NativeError as e:
        print("curve path at eps 1e-5,", "persistent" if fused else "multi-launch", "ended at", (l, h), str(e)[:60])
# several objects in one call
T = native_net(oracle_net(load_golden("tiny_sphere_h8")))
meshes = _native.subpoly_batch([T, N, T, T, N], size=1.2, eps=1e-4, force=True, in_flight=3)
print("batch", [m.sizes()["V"] for m in meshes])
del meshes
lib().tnb_release_cached_blocks()
torch.cuda.synchronize()
print("memcheck target done")
