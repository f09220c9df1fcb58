"""SDF training data (reference: tropical/stanford/dataset.py).

`StanfordDataset(name)` keeps the reference's interface (R, resample(), X, Y, len 50000).  The
Stanford scans need `trimesh` + `cubvh` (BVH signed distance) and the mesh files, none of
which exist offline; the names "sphere" and "torus" provide analytic SDFs with the same sampling
scheme so that `python -m tropical.stanford.train -d sphere -e` runs end to end.
"""
import os

import torch
from torch.utils.data import Dataset

ANALYTIC = ("sphere", "torus")


def analytic_sdf(name, x):
    """Inside-positive signed distance (the reference's convention, dataset.py:94)."""
    if name == "sphere":
        return 0.6 - x.norm(dim=-1)
    q = (x[:, 0] ** 2 + x[:, 1] ** 2).sqrt() - 0.55
    return 0.22 - (q ** 2 + x[:, 2] ** 2).sqrt()


class StanfordDataset(Dataset):
    def __init__(self, name: str = "dragon"):
        self.R = .8
        self.name = name
        self.init()
        self.resample()

    def __len__(self):
        return 50000

    def init(self):
        if self.name.lower() in ANALYTIC:
            self.mesh = None
            return
        try:
            import cubvh
            import trimesh
        except ImportError as e:
            raise RuntimeError(
                f"dataset '{self.name}' needs trimesh + cubvh and the Stanford scan files "
                f"(dataset.py:36-78 of the reference); offline use -d sphere or -d torus") from e
        base = os.path.dirname(__file__)
        files = {"bunny": "bunny/reconstruction/bun_zipper.ply", "armadillo": "armadillo/Armadillo.ply",
                 "drill": "drill/reconstruction/drill_shaft_vrip.ply", "lucy": "lucy/lucy_res10.ply"}
        path = files.get(self.name.lower(), f"{self.name}_recon/{self.name}_vrip_res3.ply")
        print(f"Loading {os.path.basename(path)} ...")
        self.mesh = trimesh.load(os.path.join(base, path))
        v = torch.Tensor(self.mesh.vertices)
        v = v / (v.max(dim=0)[0] - v.min(dim=0)[0]).max() * 2
        v -= (v.max(dim=0)[0] + v.min(dim=0)[0]) / 2
        self.mesh.vertices = v.numpy()
        self.BVH = cubvh.cuBVH(self.mesh.vertices, self.mesh.faces)
        print("BVH initialized.", flush=True)

    def resample(self):
        n = len(self)
        if self.mesh is None:
            # points near the analytic surface, jittered like the reference jitters mesh vertices
            u = torch.randn(n, 3)
            u = u / u.norm(dim=-1, keepdim=True)
            if self.name.lower() == "sphere":
                base = 0.6 * u
            else:
                phi = torch.rand(n) * 2 * torch.pi
                ring = torch.stack([0.55 * torch.cos(phi), 0.55 * torch.sin(phi), torch.zeros(n)], -1)
                base = ring + 0.22 * u
            d = 0.4
            self.X = base + (torch.rand(n, 3) * d - d / 2)
            self.Y = analytic_sdf(self.name.lower(), self.X)
            return
        vertices = torch.Tensor(self.mesh.vertices)
        if "lucy" != self.name.lower():
            vertices = vertices.repeat(10, 1)
        d = 0.4
        if vertices.shape[0] < n:
            vertices = torch.Tensor(self.mesh.vertices).repeat(30, 1)
            d = 0.2
        points = vertices[torch.randperm(vertices.shape[0])[:n]] + (torch.rand(n, 3) * d - d / 2)
        distances, _, _ = self.BVH.signed_distance(points)
        self.X, self.Y = points, distances.cpu()

    def __getitem__(self, idx):
        return self.X[idx], self.Y[idx]
