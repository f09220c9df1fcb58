"""One extraction (and one dense sign sweep) inside a cudaProfilerStart/Stop range, after warm-up:
the target of the ncu captures (`ncu --profile-from-start off ...`).

    python profiles/one_extraction.py [workload] [planar|curve] [sweep_n]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    name = sys.argv[1] if len(sys.argv) > 1 else "large_sphere"
    planar = (sys.argv[2] if len(sys.argv) > 2 else "planar") == "planar"
    sweep_n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    torch.cuda.set_device(0)
    w = bench.load_workload(name)
    net = bench.make_native(w)
    for _ in range(3):
        net.subpoly(size=1.2, eps=w["eps"], force=planar)
    buf = torch.empty((sweep_n ** 3, 2), dtype=torch.int64, device="cuda") if sweep_n > 0 else None
    if buf is not None:
        net.sweep_signs((-1, -1, -1), (1, 1, 1), (sweep_n,) * 3, out=buf)
    torch.cuda.synchronize()
    rt = ctypes.CDLL("libcudart.so")
    rt.cudaProfilerStart()
    mesh = net.subpoly(size=1.2, eps=w["eps"], force=planar)
    if buf is not None:
        net.sweep_signs((-1, -1, -1), (1, 1, 1), (sweep_n,) * 3, out=buf)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
    print(mesh.sizes())


if __name__ == "__main__":
    main()
