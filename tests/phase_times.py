"""Manual: device time of the three phases of one extraction (skeleton | 33 steps | faces)."""
import os, sys, time
import numpy as np, torch
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tropical-nerf.pytorch_b200"), HERE):
    sys.path.insert(0, p)
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "large_sphere"
force = not (len(sys.argv) > 2 and sys.argv[2] == "curve")  # second argument "curve": force=False
w = bench.load_workload(name)
net = bench.make_native(w)
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(4):
    e = [ev() for _ in range(4)]
    t0 = time.perf_counter()
    e[0].record(); c = net.skeleton(128); e[1].record()
    lh = [(l, h) for l in range(net.num_layers - 1) for h in range(net.num_hidden)] + [(net.num_layers - 2, net.num_hidden)]
    c.steps(lh, force=force); e[2].record()
    m = c.extract_mesh(); e[3].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(name, "skeleton %.3f steps %.3f faces %.3f total %.3f ms (host %.3f) V,E after steps" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), e[0].elapsed_time(e[3]), (t1 - t0) * 1e3), c.num_vertices, c.num_edges, m.sizes())
# per-step sizes (syncs each step: only for the trace)
c = net.skeleton(128)
print("skeleton V,E", c.num_vertices, c.num_edges)
for l in range(net.num_layers - 1):
    for h in range(net.num_hidden):
        a = ev(); b = ev(); a.record(); c.step(l, h, force=force); b.record(); torch.cuda.synchronize()
        print("step", l, h, "%.3f ms" % a.elapsed_time(b), c.num_vertices, c.num_edges)
