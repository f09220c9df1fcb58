"""`python -m tropical.stanford.train` -- the reference's entry point (tropical/stanford/train.py).

Same command line (-d/-s/-c/-m/-e/-f with the same defaults and the same inverted -c / -f
switches), same phases: fit the SDF (or load the cached model), extract the polyhedral mesh
with `tropical.subpoly.subpoly`, write it as .ply, optionally evaluate.  The mesh extraction
runs on the sm_100a path.  Offline additions: `-d sphere|torus` (analytic SDFs), and the
evaluation falls back to what is computable without mcubes / cubvh / trimesh.
"""
import argparse
import os
import random
import time

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import DataLoader

import tropical.subpoly as sp
from tropical.stanford.dataset import ANALYTIC, StanfordDataset, analytic_sdf
from tropical.stanford.model import Net

DIM = 3
CANVAS_SIZE = 1.2
BATCH_SIZE = 1000


def parse(argv=None):
    p = argparse.ArgumentParser(prog="python -m tropical.stanford.train",
                                description="Polyhedral complex derivation from piecewise trilinear networks")
    p.add_argument("-d", "--dataset", default="dragon",
                   choices=["bunny", "dragon", "happy", "armadillo", "drill", "lucy", "bunny_npy", *ANALYTIC],
                   help="Stanford 3D scanning model name (sphere/torus: analytic SDFs, usable offline)")
    p.add_argument("-s", "--seed", default=45, type=int, help="Seed")
    p.add_argument("-c", "--cache", default=True, action="store_false", help="Cache the trained SDF?")
    p.add_argument("-m", "--model_size", default="small", choices=["small", "medium", "large"], help="Model size")
    p.add_argument("-e", "--eval", default=False, action="store_true", help="Run evaluation?")
    p.add_argument("-f", "--force", default=True, action="store_false",
                   help="Force flat assumption to skip curve approximation.")
    p.add_argument("--epochs", default=None, type=int, help="(addition) override the number of epochs")
    return p.parse_args(argv)


def write_ply(path, vertices, triangles):
    """ASCII .ply of a triangle mesh (the reference goes through trimesh.Trimesh.export)."""
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\n")
        f.write(f"element vertex {len(vertices)}\nproperty float x\nproperty float y\nproperty float z\n")
        f.write(f"element face {len(triangles)}\nproperty list uchar int vertex_indices\nend_header\n")
        np.savetxt(f, vertices, fmt="%.8g")
        np.savetxt(f, np.concatenate([np.full((len(triangles), 1), 3), triangles], 1), fmt="%d")


def main(argv=None):
    args = parse(argv)
    print(args)
    seed = args.seed
    torch.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)
    epochs = args.epochs if args.epochs is not None else (6 if "drill" == args.dataset else 10)
    r_min, r_max = {"small": (2, 32), "medium": (4, 64), "large": (8, 128)}[args.model_size]
    T = 21 if ("large" == args.model_size and "bunny" in args.dataset.lower()) else 19  # train.py:80
    net = Net(num_layers=3, num_hidden=16, levels=4, r_min=r_min, r_max=r_max, T=T).cuda()
    data = StanfordDataset(args.dataset)
    loader = DataLoader(data, batch_size=BATCH_SIZE, shuffle=True)
    criterion = nn.L1Loss()
    optimizer = torch.optim.Adam(net.parameters(), lr=1e-3)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, epochs * len(data) / BATCH_SIZE)
    model_path = os.path.join(os.path.dirname(__file__),
                              f"models/{args.dataset}/{args.dataset}_sdf_{args.model_size}_{seed}.pth")

    def extract():
        t = time.time()
        faces, vertices, tri = sp.subpoly(net, DIM, CANVAS_SIZE, force=args.force)
        dt = time.time() - t
        print(f" take {dt:.2f}")
        return faces, vertices, tri, dt

    result = None
    for epoch in range(epochs):
        if args.cache and os.path.isfile(model_path):
            net.load_state_dict(torch.load(model_path, map_location=net.device()))
            print(f"The pretrained model loaded from {model_path}")
            result = extract()
            break
        if epoch == 0:
            print(f"warning: cannot find a pretrained model for seed ({seed})! This training code does not "
                  f"guarantee convergence nor a reliable SDF.", flush=True)
        running = 0.0
        data.resample()
        for i, (inputs, labels) in enumerate(loader):
            inputs, labels = inputs.cuda(), labels.cuda()
            optimizer.zero_grad()
            pts = inputs.clone().requires_grad_(True)
            sdf = net.sdf(pts)
            l1 = criterion(torch.clamp(sdf[:, 0], -0.2, 0.2), torch.clamp(labels, -0.2, 0.2))
            J = torch.autograd.grad(sdf.sum(), pts, create_graph=True)[0]            # eikonal, train.py:194-197
            loss = l1 + 1e-2 * (J.norm(p=2) - 1).pow(2) / BATCH_SIZE
            loss = loss + 1e-1 * sum((1 - fc.weight.norm(p=2, dim=1)).pow(2).mean() for fc in net.fc) / len(net.fc)
            loss.backward()
            optimizer.step()
            scheduler.step()
            running += loss.item()
            if i % 10 == 9:
                print(f"[{epoch + 1}, {i + 1:5d}] lr: {scheduler.get_last_lr()[0]:.4f}, loss: {running / 10:.5f} "
                      f"l1: {l1.item() / 10:.5f}", end="")
                running = 0.0
                it = len(data) * epoch // BATCH_SIZE // 10 + (i + 1) // 10
                if 5 * epochs > it:
                    print(" mesh calculation skipped.")
                    continue
                result = extract()
    print("Finished training.", flush=True)
    if result is None:
        result = extract()
    if args.cache:
        os.makedirs(os.path.dirname(model_path), exist_ok=True)
        torch.save(net.state_dict(), model_path)

    faces, vertices, triangles, our_t = result
    vertices = vertices.cpu().numpy() / data.R
    print(f"Ours: {vertices.shape}/{np.asarray(triangles).shape}")
    os.makedirs(f"meshes/{args.dataset}", exist_ok=True)
    out = os.path.join(f"meshes/{args.dataset}", f"our_mesh_{args.model_size}_{seed}.ply")
    write_ply(out, vertices, np.asarray(triangles))
    print(f"mesh written to {out}")
    if not args.eval:
        return
    # Evaluation.  The reference compares against marching cubes at 12 resolutions through
    # pymcubes + cubvh ray casting (train.py:276-355); those packages are not part of the
    # extraction path.  What can be stated without them:
    if vertices.shape[0] == 0:
        print("Ours,     0 vertices: the fitted SDF has no zero level set inside the grid (try another seed).")
        return
    with torch.no_grad():
        v = torch.from_numpy(vertices * data.R).cuda()
        residual = net.sdf(v)[:, 0].abs()
        print(f"Ours, {vertices.shape[0]:5d} vertices, |sdf| at the mesh vertices: max {residual.max():.2e}, "
              f"mean {residual.mean():.2e}, extraction {our_t:.2f} s")
        if args.dataset in ANALYTIC:
            err = analytic_sdf(args.dataset, v.cpu()).abs()
            print(f"distance of the mesh vertices to the analytic {args.dataset}: mean {err.mean():.4f}, max {err.max():.4f}")
    try:
        import cubvh  # noqa: F401
        import mcubes  # noqa: F401
    except ImportError:
        print("Marching Cubes comparison skipped: pymcubes / cubvh are not installed.")


if __name__ == "__main__":
    main()
