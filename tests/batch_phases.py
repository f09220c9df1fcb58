"""Manual: which phase of a small extraction overlaps with itself when K host threads (one stream each) run it?
usage: python tests/batch_phases.py [workload]"""
import os, sys, threading, time
import torch
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tropical-nerf.pytorch_b200"), HERE):
    sys.path.insert(0, p)
import bench
from tropical import _native

name = sys.argv[1] if len(sys.argv) > 1 else "small_sphere"
w = bench.load_workload(name)
net = bench.make_native(w)
lh = [(l, h) for l in range(net.num_layers - 1) for h in range(net.num_hidden)] + [(net.num_layers - 2, net.num_hidden)]


def run(label, K, M, prepare, body):
    streams = [torch.cuda.Stream() for _ in range(K)]
    bar = threading.Barrier(K + 1)

    def work(i):
        with torch.cuda.stream(streams[i]):
            st = prepare()
            body(st)
            streams[i].synchronize()
            bar.wait()
            for _ in range(M):
                body(st)
            streams[i].synchronize()
    th = [threading.Thread(target=work, args=(i,)) for i in range(K)]
    for t in th: t.start()
    bar.wait()
    t0 = time.perf_counter()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(f"{label:28s} K={K:2d}: {dt / M * 1e3:7.3f} ms per round of K  ({dt / M / K * 1e3:6.3f} ms per object)", flush=True)


def full_complex():
    c = net.skeleton(128)
    c.steps(lh)
    return c


for cluster in (0, 200000):
    _native.lib().tnb_set_cluster_max_items(cluster)
    for K in (1, 4, 8):
        run(f"whole (cluster={cluster})", K, 16, lambda: None, lambda st: net.subpoly())
        run(f"skeleton", K, 16, lambda: None, lambda st: net.skeleton(128))
        run(f"skeleton+steps (cluster={cluster})", K, 16, lambda: None, lambda st: net.skeleton(128).steps(lh))
        run(f"faces", K, 16, full_complex, lambda c: c.extract_mesh())
_native.lib().tnb_set_cluster_max_items(0)
