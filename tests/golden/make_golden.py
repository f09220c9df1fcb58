"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

For each case: build the reference's `Net`, fit it briefly to an analytic SDF, run the
reference's own `tropical.subpoly` functions on CPU (through `refenv`), and store
  * the network (hash table, MLP, marks) -- the input every arm gets,
  * the reference's skeleton, the complex after every hyperplane (sizes) and before
    face extraction (arrays), the surface skeleton, the polygon rows and triangles.
The reference's float results depend on MKL's summation order; the integer arrays
(edges, polygons) are what the parity tests hold bit-exact, positions are compared
within 1e-5.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refenv  # noqa: E402

CASES = {
    # name: (Net kwargs, sdf, fit steps, torch seed)
    "small_sphere": (dict(num_layers=3, num_hidden=16, levels=4, r_min=2, r_max=32, T=19),
                     "sphere", 200, 1),
    "small_torus": (dict(num_layers=3, num_hidden=16, levels=4, r_min=2, r_max=32, T=19),
                    "torus", 200, 3),
    "tiny_sphere_h8": (dict(num_layers=3, num_hidden=8, levels=2, r_min=2, r_max=8, T=19),
                       "sphere", 200, 5),
}


def fit_fast(net, sdf_fn, steps, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    for _ in range(steps):
        x = torch.rand(8192, 3, generator=g) * 2 - 1
        y = sdf_fn(x)
        opt.zero_grad()
        loss = (net.sdf(x)[:, 0] - y.clamp(-0.3, 0.3)).abs().mean()
        loss.backward()
        opt.step()
    return float(loss)


def run_case(name, kwargs, sdf_name, steps, seed):
    import torch
    tropical, sp, Net = refenv.import_reference()
    import tropical.geometry as gm
    from oracle.trinet import NetParams

    torch.manual_seed(seed)
    net = Net(**kwargs)
    sdf_fn = refenv.sphere_sdf if sdf_name == "sphere" else refenv.torus_sdf
    loss = fit_fast(net, sdf_fn, steps, seed)
    P = NetParams.from_reference_net(net)
    out = {f"net_{k}": v for k, v in P.to_npz_dict().items()}
    eps = 1e-4
    t0 = time.time()
    with torch.no_grad():
        v, e = net.enc.skeleton(net)
        out["skeleton_vertices"], out["skeleton_edges"] = v.numpy(), e.numpy().astype(np.int32)
        outputs = None
        sizes = []
        for l in range(net.num_layers - 1):
            for h in range(net.num_hidden):
                v, e, outputs = sp.subpoly_(v, e, net, l, h, eps, outputs, force=True)
                sizes.append((v.shape[0], e.shape[0]))
        v, e, outputs = sp.subpoly_(v, e, net, net.num_layers - 2, net.num_hidden, eps, outputs,
                                    force=True)
        sizes.append((v.shape[0], e.shape[0]))
        out["step_sizes"] = np.array(sizes, np.int64)
        out["complex_vertices"], out["complex_edges"] = v.numpy(), e.numpy().astype(np.int32)
        sv, se, vidx = sp.extract_skeleton(v, e, net, eps, outputs)
        out["surface_vertices"], out["surface_edges"] = sv.numpy(), se.numpy().astype(np.int32)
        so = outputs[vidx]
    # the reference's extract_faces steps (subpoly.py:606-648), keeping the polygon rows
    m_rgn, offset, _ = net.region(sv, so, eps)
    r_idx, aug = sp.regions_to_vertices(m_rgn[:, :-1], offset, return_inverse=True)
    v_indices = sp.r_idx_as_tensor(r_idx, aug, offset).unique(dim=0)
    mean_points, points, v_indices = sp.mean_points_with_valid(sv, v_indices, return_points=True)
    jac = net.normal(mean_points)
    _, indices = gm.sort_polygon_vertices_batch(points, jac, return_index=True)
    out["polygons"] = v_indices.gather(1, indices).numpy().astype(np.int32)
    faces, vertices, tri = sp.subpoly(net, 3, 1.2, force=True)
    out["triangles"] = np.asarray(tri).astype(np.int32)
    out["ref_seconds_cpu"] = np.float64(time.time() - t0)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(f"\n{name}: fit loss {loss:.4f}; marks {P.marks.size}; skeleton "
          f"{out['skeleton_vertices'].shape[0]}/{out['skeleton_edges'].shape[0]}; surface "
          f"{sv.shape[0]} vertices, {out['polygons'].shape[0]} polygons, "
          f"{out['triangles'].shape[0]} triangles", flush=True)


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, (kwargs, sdf_name, steps, seed) in CASES.items():
        if only and name not in only:
            continue
        run_case(name, kwargs, sdf_name, steps, seed)
