"""Golden fixture of the gradient-descent repair (subpoly_debug.deal_with_gradient_descent,
subpoly_debug.py:121-165), produced by the UNMODIFIED reference function.

    python tests/golden/make_golden_gd.py        # writes tests/golden/gd_stage.npz (a few minutes)

Two kinds of cases, all on the network of tests/golden/small_torus.npz:

* `nat_*`: the one call the reference itself makes when the small torus is extracted along the curve
  path with eps = 1e-5 (hyperplane 1/5, one edge): the 500 bodies do not bring the intersection back
  onto its planes, and the reference run ENDS there (subpoly.py:172-174; the stock code dies with a
  NameError on its way to the `exit()` of debug_test_idx).  The inputs of the call, its results and
  the step at which the run ended are stored.
* `syn<k>_*`: calls of the same function on crafted inputs (short edges next to the zero sets of the
  hyperplane and of an earlier one) for which the loop DOES end early, with several edges sharing the
  one iteration count; a row without an admissible intersection (gg) rides along untouched.
* `lad<k>_*`: the same edges, alone and together, with eps set so that the loop ends after a few steps
  and not near a tie (sharp comparison of the arithmetic).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refenv  # noqa: E402
from make_golden_curve import reference_net  # noqa: E402


def call_reference(dbg, net, e, ints, plane, idx, eps):
    """d_new as subpoly.py:158-161 forms it, then the reference's repair.  Returns (ints, d_new) before and after."""
    import torch
    with torch.no_grad():
        G = e.shape[0]
        c = torch.ones(G, dtype=torch.bool)
        inds = torch.stack([torch.arange(G), plane], -1)
        _, _, outs = net.region(e[:, 0] * (1 - ints) + e[:, 1] * ints)
        d_new = torch.stack([outs.gather(-1, inds[:, 1:]).squeeze(1), outs[:, idx]], dim=-1)
        gg = 0 < ((ints < 0) | (ints > 1)).sum(-1)
        ints0, d0 = ints.clone(), d_new.clone()
        ints1, d1 = dbg.deal_with_gradient_descent(c, d_new, e, eps, gg, idx, inds, ints, net)
    return ints0, d0, ints1.detach().clone(), d1.detach().clone(), gg


def natural_case(net, sp, dbg):
    """The curve-path extraction of the small torus with eps = 1e-5 up to the call that needs the repair."""
    import torch
    captured = {}
    stock = dbg.deal_with_gradient_descent

    def spy(c, d_new, e, eps, gg, idx, inds, ints, net_):
        gd = ~gg & (0 < (d_new.abs() > eps).sum(dim=-1))
        if 0 < gd.sum() and not captured:
            captured.update(e=e[c].clone(), ints=ints.clone(), d_new=d_new.clone(), plane=inds[:, 1].clone(), idx=idx, gg=gg.clone())
            r = stock(c, d_new, e, eps, gg, idx, inds, ints, net_)
            captured.update(ints_out=r[0].detach().clone(), d_out=r[1].detach().clone())
            return r
        return stock(c, d_new, e, eps, gg, idx, inds, ints, net_)

    dbg.deal_with_gradient_descent = spy
    ended = None
    sizes = []
    with torch.no_grad():
        v, e = net.enc.skeleton(net)
        outputs = None
        H = net.num_hidden
        for (l, h) in [(l, h) for l in range(net.num_layers - 1) for h in range(H)] + [(net.num_layers - 2, H)]:
            try:
                v, e, outputs = sp.subpoly_(v, e, net, l, h, 1e-5, outputs, force=False)
            except BaseException as ex:  # NameError on the way to exit()
                ended = (l, h, type(ex).__name__)
                break
            sizes.append((v.shape[0], e.shape[0]))
    dbg.deal_with_gradient_descent = stock
    assert captured and ended is not None, "the reference run was expected to need the repair and to end there"
    print("\nreference run ended at", ended)
    res = {"nat_eps": np.float32(1e-5), "nat_ended_at": np.array(ended[:2], np.int64), "nat_ended_with": np.array(ended[2]),
           "nat_idx": np.int64(captured["idx"]), "nat_step_sizes": np.array(sizes, np.int64)}
    for k in ("e", "ints", "d_new", "plane", "gg", "ints_out", "d_out"):
        res[f"nat_{k}"] = captured[k].numpy()
    return res


if __name__ == "__main__":
    import torch
    g = dict(np.load(os.path.join(HERE, "small_torus.npz")))
    net, sp = reference_net(g)
    import tropical.subpoly_debug as dbg
    out = {}

    # ---- the reference's own call (curve path, eps = 1e-5) --------------------------------------
    part = os.path.join(HERE, "gd_stage_nat.part.npz")   # scratch of an earlier run of this script (git-ignored)
    if os.path.exists(part):
        out.update(dict(np.load(part)))
    else:
        out.update(natural_case(net, sp, dbg))
        np.savez_compressed(part, **out)

    # ---- crafted calls that end early ---------------------------------------------------------
    # Points next to the zero sets of the hyperplane AND of an earlier one, short edges around them; the edges
    # whose own walk comes to rest within eps are put in one call (they reach it after different numbers of
    # steps, the call takes the steps of the slowest), next to one that is within eps from the start and one
    # without an admissible intersection.
    R = (net.num_layers - 1) * net.num_hidden + 1
    n_found = 0
    prev = os.path.join(HERE, "gd_stage.npz")
    if os.path.exists(prev) and "--keep-syn" in sys.argv:   # reuse the crafted calls of an earlier run
        old = dict(np.load(prev))
        n_found = int(old["n_syn"])
        out.update({k_: v_ for k_, v_ in old.items() if k_.startswith("syn")})
    for seed in ([] if n_found else range(1, 12)):
        if n_found == 3:
            break
        gen = torch.Generator().manual_seed(seed)
        idx = int(torch.randint(8, R, (1,), generator=gen))
        x = torch.rand(1000000, 3, generator=gen) * 1.6 - 0.8
        with torch.no_grad():
            o = net.region(x)[2]
        a = o[:, :idx].abs()
        plane_all = a.argmin(-1)
        score = torch.maximum(a.gather(-1, plane_all[:, None]).squeeze(1), o[:, idx].abs())
        near = score.argsort()[:16]
        eps = 1e-5
        if not (score[near[0]] < 3 * eps < score[near[-1]] * 3):
            print(f"\nseed {seed} idx {idx}: no points close enough ({score[near[0]]:.2e})")
            continue
        p, plane = x[near], plane_all[near]
        delta = (torch.rand(16, 3, generator=gen) - 0.5) * 0.01
        ints = torch.rand(16, 3, generator=gen) * 0.5 + 0.25
        e0 = p - ints * delta
        e = torch.stack([e0, e0 + delta], 1)
        keep = []
        for r in range(16):   # the reference's own verdict on every edge by itself
            i0, d0, i1, d1, gg = call_reference(dbg, net, e[r:r + 1], ints[r:r + 1].clone(), plane[r:r + 1], idx, eps)
            if bool((d1.abs() <= eps).all()):
                keep.append(r)
        print(f"\nseed {seed} idx {idx}: edges whose walk ends within eps (or starts there): {keep}")
        if len(keep) < 3:
            continue
        sel = torch.tensor(keep)
        e_s, ints_s, plane_s = e[sel], ints[sel].clone(), plane[sel]
        e_s = torch.cat([e_s, e_s[:1]], 0)
        ints_s = torch.cat([ints_s, torch.tensor([[0.5, 1.5, 0.5]])], 0)   # no admissible intersection: untouched
        plane_s = torch.cat([plane_s, plane_s[:1]], 0)
        i0, d0, i1, d1, gg = call_reference(dbg, net, e_s, ints_s.clone(), plane_s, idx, eps)
        gd = ~gg & (0 < (d0.abs() > eps).sum(-1))
        ok = bool((d1[~gg].abs() <= eps).all())
        print(f"\n  joint call: {int(gd.sum())} of {len(gd)} edges need the repair, ended early: {ok}")
        if ok and int(gd.sum()) >= 2:
            k = n_found
            out[f"syn{k}_eps"] = np.float32(eps)
            out[f"syn{k}_idx"] = np.int64(idx)
            out[f"syn{k}_e"] = e_s.numpy()
            out[f"syn{k}_ints"] = i0.numpy()
            out[f"syn{k}_d_new"] = d0.numpy()
            out[f"syn{k}_plane"] = plane_s.numpy()
            out[f"syn{k}_gg"] = gg.numpy()
            out[f"syn{k}_ints_out"] = i1.numpy()
            out[f"syn{k}_d_out"] = d1.numpy()
            n_found += 1
    out["n_syn"] = np.int64(n_found)

    # ---- calls that end after a FEW steps, away from any tie ---------------------------------------
    # The walks above zigzag across their planes for hundreds of steps before every one of them is within eps at
    # the same step: which step that is depends on the last bits of the gradient.  For a sharp comparison of the
    # arithmetic: the same edges with eps a fraction of their first distance, kept only if the reference's answer
    # does not change when eps moves by 3 % either way (the stopping step is then not a matter of rounding).
    k = 0
    if n_found:
        e_s, plane_s, idx = torch.from_numpy(out["syn0_e"]), torch.from_numpy(out["syn0_plane"]), int(out["syn0_idx"])
        ints_s, d_s, gg_s = torch.from_numpy(out["syn0_ints"]), torch.from_numpy(out["syn0_d_new"]), torch.from_numpy(out["syn0_gg"])
        rows = [r for r in range(e_s.shape[0]) if not gg_s[r]]
        groups = [[r] for r in rows[:4]] + [rows[:3], rows]
        for grp in groups:
            sel = torch.tensor(grp)
            d_first = d_s[sel].abs().max(-1).values.min().item()
            for frac in (0.85, 0.6, 0.4):
                eps = d_first * frac
                res = [call_reference(dbg, net, e_s[sel], ints_s[sel].clone(), plane_s[sel], idx, eps * f) for f in (1.0, 0.97, 1.03)]
                i0, d0, i1, d1, gg = res[0]
                clean = all(torch.equal(r[2], i1) for r in res[1:]) and bool((d1.abs() <= eps * 0.97).all())
                print(f"\n  ladder rows {grp} eps {eps:.3e}: clean {clean}")
                if not clean:
                    continue
                out[f"lad{k}_eps"] = np.float32(eps)
                out[f"lad{k}_idx"] = np.int64(idx)
                out[f"lad{k}_e"] = e_s[sel].numpy()
                out[f"lad{k}_ints"] = i0.numpy()
                out[f"lad{k}_d_new"] = d0.numpy()
                out[f"lad{k}_plane"] = plane_s[sel].numpy()
                out[f"lad{k}_gg"] = gg.numpy()
                out[f"lad{k}_ints_out"] = i1.numpy()
                out[f"lad{k}_d_out"] = d1.numpy()
                k += 1
    out["n_lad"] = np.int64(k)
    assert n_found > 0
    np.savez_compressed(os.path.join(HERE, "gd_stage.npz"), **out)
    print("wrote gd_stage.npz with", n_found, "crafted calls")
