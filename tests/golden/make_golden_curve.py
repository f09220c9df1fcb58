"""Golden fixtures of the curve-approximation path (force=False), produced by the UNMODIFIED
reference from the networks already stored in tests/golden/<case>.npz.

    python tests/golden/make_golden_curve.py        # writes tests/golden/<case>_curve.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refenv  # noqa: E402

R_MAX = {2: 8, 4: 32}


def reference_net(g):
    import torch
    _, sp, Net = refenv.import_reference()
    L = int(g["net_levels"])
    net = Net(num_layers=int(g["net_num_layers"]), num_hidden=int(g["net_num_hidden"]), levels=L,
              r_min=int(g["net_n_min"]), r_max=R_MAX[L], T=int(g["net_log2_T"]))
    sd = {"enc.module.params": torch.from_numpy(g["net_table"])}
    for i in range(int(g["net_num_layers"])):
        sd[f"fc.{i}.weight"] = torch.from_numpy(g[f"net_w{i}"])
        sd[f"fc.{i}.bias"] = torch.from_numpy(g[f"net_b{i}"])
    net.load_state_dict(sd)
    assert np.array_equal(net.enc.marks.numpy(), g["net_marks"])
    return net, sp


if __name__ == "__main__":
    import torch
    for case in ["tiny_sphere_h8", "small_sphere", "small_torus"]:
        g = dict(np.load(os.path.join(HERE, f"{case}.npz")))
        net, sp = reference_net(g)
        out = {}
        with torch.no_grad():
            v, e = net.enc.skeleton(net)
            outputs, sizes = None, []
            H = net.num_hidden
            for (l, h) in [(l, h) for l in range(net.num_layers - 1) for h in range(H)] + [(net.num_layers - 2, H)]:
                v, e, outputs = sp.subpoly_(v, e, net, l, h, 1e-4, outputs, force=False)
                sizes.append((v.shape[0], e.shape[0]))
        out["step_sizes"] = np.array(sizes, np.int64)
        faces, vertices, tri = sp.subpoly(net, 3, 1.2, force=False)
        out["surface_vertices"] = vertices.numpy()
        out["triangles"] = np.asarray(tri).astype(np.int32)
        np.savez_compressed(os.path.join(HERE, f"{case}_curve.npz"), **out)
        print(f"\n{case}: curve path -> {vertices.shape[0]} vertices, {len(tri)} triangles", flush=True)
