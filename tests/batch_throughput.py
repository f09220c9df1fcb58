"""Manual: throughput of K concurrent extractions on ONE GPU (K host threads, one CUDA stream each).
usage: python tests/batch_throughput.py [workload] [cluster_max_items]"""
import os, sys, threading, time
import torch
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tropical-nerf.pytorch_b200"), HERE):
    sys.path.insert(0, p)
import bench
from tropical import _native

name = sys.argv[1] if len(sys.argv) > 1 else "small_sphere"
cluster = int(sys.argv[2]) if len(sys.argv) > 2 else 0
w = bench.load_workload(name)
net = bench.make_native(w)
_native.lib().tnb_set_cluster_max_items(cluster)
for _ in range(3):
    m = net.subpoly(); nv = m.sizes()["V"]; del m
torch.cuda.synchronize()
for K in (1, 2, 4, 8, 12, 16):
    M = 24
    streams = [torch.cuda.Stream() for _ in range(K)]
    bar = threading.Barrier(K + 1)
    def work(i):
        with torch.cuda.stream(streams[i]):
            m = net.subpoly(); del m          # warm the thread
            streams[i].synchronize()
            bar.wait()
            for _ in range(M):
                m = net.subpoly()
                del m
            streams[i].synchronize()
    th = [threading.Thread(target=work, args=(i,)) for i in range(K)]
    for t in th: t.start()
    bar.wait()
    t0 = time.perf_counter()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(f"{name} cluster_max={cluster} K={K}: {K*M/dt:8.1f} objects/s  {K*M*nv/dt/1e6:7.2f} M vertices/s  ({dt/M*1e3:.3f} ms per round of K)")
