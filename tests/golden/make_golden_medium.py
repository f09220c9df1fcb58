"""Golden fixture of BASELINE.json configs[2]: the MEDIUM model (r = 4..64, marks grid 98^3) fitted to
the analytic torus, extracted by the UNMODIFIED reference on CPU with the curve-approximation path
(force=False) and, for the same network, with the planar path.

    python tests/golden/make_golden_medium.py        # writes tests/golden/medium_torus.npz

The hash table of the fitted network is rounded to fp16-representable values BEFORE the reference
runs (the network IS those values), so the fixture stores it in 2 bytes per entry.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refenv  # noqa: E402
from make_golden import fit_fast  # noqa: E402


def main(name="medium_torus", sdf_name="torus", steps=300, seed=7):
    import torch
    _, sp, Net = refenv.import_reference()
    from oracle.trinet import NetParams
    torch.manual_seed(seed)
    net = Net(num_layers=3, num_hidden=16, levels=4, r_min=4, r_max=64, T=19)
    sdf_fn = refenv.sphere_sdf if sdf_name == "sphere" else refenv.torus_sdf
    loss = fit_fast(net, sdf_fn, steps, seed)
    with torch.no_grad():
        p = net.enc.module.params
        p.copy_(p.half().float())
    P = NetParams.from_reference_net(net)
    out = {f"net_{k}": v for k, v in P.to_npz_dict().items()}
    out["net_table"] = out["net_table"].astype(np.float16)
    assert np.array_equal(out["net_table"].astype(np.float32), P.table)
    H = net.num_hidden
    planes = [(l, h) for l in range(net.num_layers - 1) for h in range(H)] + [(net.num_layers - 2, H)]
    for tag, force in (("curve", False), ("planar", True)):
        t0 = time.time()
        with torch.no_grad():
            v, e = net.enc.skeleton(net)
            outputs, sizes = None, []
            for (l, h) in planes:
                v, e, outputs = sp.subpoly_(v, e, net, l, h, 1e-4, outputs, force=force)
                sizes.append((v.shape[0], e.shape[0]))
        out[f"{tag}_step_sizes"] = np.array(sizes, np.int64)
        out[f"{tag}_complex_edges"] = e.numpy().astype(np.int32)
        faces, vertices, tri = sp.subpoly(net, 3, 1.2, force=force)
        out[f"{tag}_surface_vertices"] = vertices.numpy()
        out[f"{tag}_triangles"] = np.asarray(tri).astype(np.int32)
        out[f"{tag}_ref_seconds_cpu"] = np.float64(time.time() - t0)
        print(f"\n{name} [{tag}]: {vertices.shape[0]} vertices, {len(tri)} triangles, "
              f"{time.time() - t0:.1f} s (two passes)", flush=True)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"{name}: fit loss {loss:.4f}; marks {P.marks.size}", flush=True)


if __name__ == "__main__":
    main()
