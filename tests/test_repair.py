"""The gradient-descent repair of the curve path (subpoly_debug.py:121-165).

CPU: the oracle's restatement (oracle/trinet_ref.c curve_gradient_descent) against tests/golden/gd_stage.npz --
inputs and results of the UNMODIFIED reference function (tests/golden/make_golden_gd.py): the one call the
reference itself makes on the small torus with eps = 1e-5 (500 steps, still off the planes: the reference run
ends there, and so does the oracle's, at the same hyperplane), and crafted calls in which several walks share an
early end (`syn`), and calls that end after a few steps away from any tie (`lad`).  The walks are 1e-2-steps along
a normalised float32 gradient that autograd and the oracle round differently: where the number of steps is not a
matter of rounding (nat: all 500; lad: 4 and 9) positions agree to 1e-4 and distances to 1e-6; the `syn` walks
zigzag across their planes for hundreds of steps until all of them are within eps at the same step, which step
that is depends on the last bits, so there the positions agree to one step of the walk (1e-2) and the distances
to 2 eps.  Who is within eps at the end agrees exactly everywhere.

GPU: the device (csrc/repair.cuh through tnb_curve_gradient_descent, and fused in the step kernels) against the
oracle, bit for bit.
"""
import numpy as np
import pytest

from helpers import load_golden, native_net, oracle_net

TOL_INTS, TOL_D = 1e-4, 1e-6   # nat, lad: the number of steps is not a matter of rounding
STEP = 1.05e-2                 # syn: one step of the walk (subpoly_debug.py:149)


def _cases():
    g = load_golden("gd_stage")
    names = ["nat"] + [f"syn{k}" for k in range(int(g["n_syn"]))] + [f"lad{k}" for k in range(int(g["n_lad"]))]
    return g, names


def _oracle_call(P, g, name):
    from oracle.trinet import gradient_descent
    eps = np.float32(g[f"{name}_eps"])
    e, ints, d_new, plane, gg = (g[f"{name}_{k}"] for k in ("e", "ints", "d_new", "plane", "gg"))
    idx = int(g[f"{name}_idx"])
    gd = ~gg & ((np.abs(d_new) > eps).sum(-1) > 0)
    x, d, bodies = gradient_descent(P, e[gd, 0], e[gd, 1], ints[gd], plane[gd], idx, eps)
    ints_out, d_out = ints.copy(), d_new.copy()
    ints_out[gd], d_out[gd] = x, d
    return gd, ints_out, d_out, bodies


def test_oracle_repair_matches_reference_function():
    g, names = _cases()
    P = oracle_net(load_golden("small_torus"))
    assert len(names) >= 4
    for name in names:
        eps = np.float32(g[f"{name}_eps"])
        gd, ints_out, d_out, bodies = _oracle_call(P, g, name)
        gg = g[f"{name}_gg"]
        assert gd.sum() >= 1
        # rows the reference leaves alone are left alone
        assert np.array_equal(ints_out[~gd], g[f"{name}_ints_out"][~gd])
        assert np.array_equal(d_out[~gd], g[f"{name}_d_out"][~gd])
        ref_within = np.abs(g[f"{name}_d_out"][~gg]) <= eps
        assert np.array_equal(np.abs(d_out[~gg]) <= eps, ref_within), name
        if name == "nat":
            assert bodies == 500 and not ref_within.all()
        else:
            assert bodies < 500 and ref_within.all()
        if name.startswith("syn"):
            assert np.linalg.norm(ints_out - g[f"{name}_ints_out"], axis=-1).max() <= STEP, name
            assert np.abs(d_out - g[f"{name}_d_out"]).max() <= 2 * eps, name
        else:
            assert np.abs(ints_out - g[f"{name}_ints_out"]).max() <= TOL_INTS, name
            assert np.abs(d_out - g[f"{name}_d_out"]).max() <= TOL_D, name


def test_oracle_run_ends_where_the_reference_run_ends():
    """Small torus, curve path, eps = 1e-5: same complexes up to hyperplane 1/5, where the repair leaves an
    intersection off its planes and the reference ends the extraction (subpoly.py:172-174)."""
    from oracle import subpoly_ref as R
    g = load_golden("gd_stage")
    P = oracle_net(load_golden("small_torus"))
    v, e = R.skeleton(P)
    o = P.outputs(v)
    H = P.num_hidden
    steps = [(l, h) for l in range(P.num_layers - 1) for h in range(H)] + [(P.num_layers - 2, H)]
    ended = None
    for i, (l, h) in enumerate(steps):
        try:
            v, e, o = R.subpoly_step(P, v, e, o, l, h, 1e-5, force=False)
        except R.GradientDescentFailed:
            ended = (l, h)
            break
        # the stock run rounds through BLAS: a handful of |d| vs eps ties fall the other way at this eps
        # (tests/golden/investigate_roots.py has the same effect at 1e-4); the sizes agree to 1e-3
        assert np.allclose((v.shape[0], e.shape[0]), g["nat_step_sizes"][i], rtol=1e-3), (l, h)
    assert ended == tuple(g["nat_ended_at"])


@pytest.mark.gpu
def test_device_repair_bit_exact_vs_oracle():
    import torch
    g, names = _cases()
    P = oracle_net(load_golden("small_torus"))
    N = native_net(P)
    for name in names:
        eps = np.float32(g[f"{name}_eps"])
        gd, ints_out, d_out, bodies = _oracle_call(P, g, name)
        e, ints, plane = g[f"{name}_e"], g[f"{name}_ints"], g[f"{name}_plane"]
        x, d, n, ok = N.gradient_descent(torch.from_numpy(e[gd]).cuda(), torch.from_numpy(ints[gd]).cuda(),
                                         torch.from_numpy(plane[gd]).cuda(), int(g[f"{name}_idx"]), float(eps))
        assert n == bodies, name
        assert ok == bool((np.abs(d_out[gd]) <= eps).all())
        assert np.array_equal(x.cpu().numpy(), ints_out[gd]), name
        assert np.array_equal(d.cpu().numpy(), d_out[gd]), name


@pytest.mark.gpu
def test_reference_signature_wrapper():
    """tropical.subpoly_debug.deal_with_gradient_descent with the reference's own argument list."""
    import torch
    from tropical import subpoly_debug as dbg
    g, names = _cases()
    P = oracle_net(load_golden("small_torus"))
    N = native_net(P)
    name = "syn0"
    eps = float(g[f"{name}_eps"])
    _, ints_out, d_out, _ = _oracle_call(P, g, name)
    e = torch.from_numpy(g[f"{name}_e"]).cuda()
    G = e.shape[0]
    c = torch.ones(G, dtype=torch.bool, device="cuda")
    inds = torch.stack([torch.arange(G), torch.from_numpy(g[f"{name}_plane"])], -1).cuda()
    ints, d_new = dbg.deal_with_gradient_descent(c, torch.from_numpy(g[f"{name}_d_new"]).cuda(), e, eps,
                                                 torch.from_numpy(g[f"{name}_gg"]).cuda(), int(g[f"{name}_idx"]), inds,
                                                 torch.from_numpy(g[f"{name}_ints"]).cuda(), N)
    assert np.array_equal(ints.cpu().numpy(), ints_out)
    assert np.array_equal(d_new.cpu().numpy(), d_out)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_device_run_ends_where_the_reference_run_ends(fused):
    """The step kernels (persistent and multi-launch) take the repair at hyperplane 1/5 of the small torus with
    eps = 1e-5, find the intersection still off its planes after 500 steps and end the extraction there, like the
    reference; up to there the complexes have the sizes of the oracle's."""
    from tropical import _native
    g = load_golden("gd_stage")
    P = oracle_net(load_golden("small_torus"))
    N = native_net(P)
    H = P.num_hidden
    steps = [(l, h) for l in range(P.num_layers - 1) for h in range(H)] + [(P.num_layers - 2, H)]
    c = N.skeleton(128)
    from oracle import subpoly_ref as R
    v, e = R.skeleton(P)
    o = P.outputs(v)
    ended = None
    for i, (l, h) in enumerate(steps):
        try:
            if fused:
                c.steps([(l, h)], eps=1e-5, force=False)   # persistent step kernel
            else:
                c.step(l, h, eps=1e-5, force=False)        # one launch per phase
            sizes = (c.num_vertices, c.num_edges)
        except _native.NativeError as ex:
            assert "gradient-descent" in str(ex)
            ended = (l, h)
            break
        v, e, o = R.subpoly_step(P, v, e, o, l, h, 1e-5, force=False)   # the oracle's complex after the same hyperplane
        assert sizes == (v.shape[0], e.shape[0]), (l, h)
    assert ended == tuple(g["nat_ended_at"])
