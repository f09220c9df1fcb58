"""Which root does the reference take, and where does a stock reference run stop being reproducible?

    python tests/golden/investigate_roots.py            # ~1 min; needs /root/reference (build container only)

Runs the UNMODIFIED reference (stock nn.Linear, stock stand-in encoding) on the medium-torus network with
the curve-approximation path and logs
  (1) every call of geometry._batched_polynomial_roots that has more than one admissible eigenvalue: the
      eigenvalues in LAPACK's order and the root the reference keeps (nonzero_last, geometry.py:296);
  (2) every call of geometry.intersection_of_two_planes: the corner values (p, q) and the result, compared
      with oracle/trinet_ref.c:curve_intersection on the SAME (p, q).
Findings on this box (torch 2.11 CPU, MKL), kept in DESIGN.md section 3:
  * 189 multi-root polynomials, 187 of them quadratics, all with ascending eigenvalue order: the reference
    keeps the LARGER root.  The two quartics (two tiny leading coefficients) went one each way.
  * on the same (p, q) the oracle agrees with the reference on which intersections are admissible; the
    coordinates differ by up to 1e-3 in y where AX - BX cancels (float32 eigenvalues carry ~1e-6 error);
  * candidates whose box is degenerate in one axis (corner pairs at the SAME point) are recognised by float
    equality of their network outputs (geometry.py:108-130); MKL rounds the same point differently in
    different rows of the batch, so 3 of 965 candidates at hyperplane 9 fail that test in the stock run and
    get split.  With row-position-independent arithmetic (make_golden_medium.py --detlin) the reference and
    the oracle agree bit for bit on every hyperplane of this network.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refenv  # noqa: E402


def main():
    import torch
    _, sp, Net = refenv.import_reference()
    import tropical.geometry as G
    from oracle.trinet import curve_intersections
    g = np.load(os.path.join(HERE, "medium_torus.npz"))
    net = Net(num_layers=3, num_hidden=16, levels=4, r_min=4, r_max=64, T=19)
    with torch.no_grad():
        net.enc.module.params.copy_(torch.from_numpy(g["net_table"].astype(np.float32)))
        for i in range(3):
            net.fc[i].weight.copy_(torch.from_numpy(g[f"net_w{i}"]))
            net.fc[i].bias.copy_(torch.from_numpy(g[f"net_b{i}"]))
    multi, calls = [], []
    roots_fn, ints_fn = G._batched_polynomial_roots, G.intersection_of_two_planes

    def roots_logged(coeffs, interval=[0, 1], eps=1e-9):
        out = roots_fn(coeffs, interval, eps)
        cf = torch.flip(coeffs, [1])
        N = cf.shape[1] - 1
        valid = cf.abs().mean(-1) > eps
        if int(valid.sum()) and N >= 2:
            C = cf.new_zeros(int(valid.sum()), N, N)
            for i in range(N - 1):
                C[:, i, i + 1] = 1
            lead = torch.ext.nonzero_last(cf[valid].abs() > eps)
            C[:, -1] = -cf[valid][:, :-1] / cf[valid].gather(1, lead[:, 1:])
            ev = torch.linalg.eigvals(C)
            ok = (ev.imag.abs() <= eps) & (ev.real >= 0) & (ev.real <= 1)
            for k in torch.nonzero(ok.sum(-1) > 1)[:, 0].tolist():
                multi.append((N, ev[k].real[ok[k]].numpy(), float(out[valid][k])))
        return out

    def ints_logged(p, q, *a, **k):
        out = ints_fn(p, q, *a, **k)
        calls.append((p.numpy().copy(), q.numpy().copy(), out.numpy().copy()))
        return out

    G._batched_polynomial_roots = roots_logged
    G.intersection_of_two_planes = ints_logged
    H = net.num_hidden
    planes = [(l, h) for l in range(net.num_layers - 1) for h in range(H)] + [(net.num_layers - 2, H)]
    with torch.no_grad():
        v, e = net.enc.skeleton(net)
        outputs = None
        for si, (l, h) in enumerate(planes):
            n0 = len(calls)
            v, e, outputs = sp.subpoly_(v, e, net, l, h, 1e-4, outputs, force=False)
            for p, q, out in calls[n0:]:
                mine = curve_intersections(p, q)
                adm_ref = ((out >= 0) & (out <= 1)).all(1)
                adm = ((mine >= 0) & (mine <= 1)).all(1)
                T_, U_ = [[0, 1, 4, 5], [0, 1, 2, 3], [0, 4, 2, 6]], [[2, 3, 6, 7], [4, 5, 6, 7], [1, 5, 3, 7]]
                near = np.zeros(p.shape[0], bool)   # degenerate up to rounding noise, but not exactly
                for t, u in zip(T_, U_):
                    d = np.maximum(np.abs(p[:, t] - p[:, u]).max(1), np.abs(q[:, t] - q[:, u]).max(1))
                    near |= (d > 0) & (d < 2e-6)
                print(f"hyperplane {si}: {p.shape[0]} curved candidates, admissible reference {int(adm_ref.sum())} / oracle "
                      f"{int(adm.sum())} (same set: {bool((adm == adm_ref).all())}), max |dx| {np.abs(mine - out)[adm & adm_ref][:, 0].max(initial=0):.1e}, "
                      f"max |dy| {np.abs(mine - out)[adm & adm_ref][:, 1].max(initial=0):.1e}; nearly degenerate boxes the reference "
                      f"solves anyway: {int((near & adm_ref).sum())}", flush=True)
    by = {}
    for N, roots, chosen in multi:
        which = "max" if chosen == roots.max() else ("min" if chosen == roots.min() else "mid")
        by[(N, which)] = by.get((N, which), 0) + 1
    print("multi-root polynomials (degree, root kept):", by)


if __name__ == "__main__":
    main()
