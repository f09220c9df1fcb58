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
torch.cuda.synchronize()
print("memcheck target done")
