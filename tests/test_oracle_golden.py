"""CPU: the oracle against the golden fixtures the UNMODIFIED reference produced
(tests/golden/make_golden.py).  Integer arrays bit-exact, positions within 1e-5."""
import numpy as np
import pytest

from helpers import CASES, canonical_polygons, load_golden, oracle_net


@pytest.fixture(scope="module", params=CASES)
def run(request):
    from oracle import subpoly_ref as R
    g = load_golden(request.param)
    P = oracle_net(g)
    faces, v, tri, inter = R.subpoly(P, return_intermediate=True)
    return g, P, faces, v, tri, inter


def test_skeleton_matches_reference(run):
    from oracle import subpoly_ref as R
    g, P = run[0], run[1]
    v, e = R.skeleton(P)
    assert np.array_equal(e, g["skeleton_edges"].astype(np.int64))
    assert np.abs(v - g["skeleton_vertices"]).max() <= 1e-6


def test_complex_before_extraction_matches_reference(run):
    g, _, _, _, _, inter = run
    assert np.array_equal(inter["edges"], g["complex_edges"].astype(np.int64))
    assert inter["vertices"].shape == g["complex_vertices"].shape
    assert np.abs(inter["vertices"] - g["complex_vertices"]).max() <= 1e-5


def test_every_step_size_matches_reference(run):
    from oracle import subpoly_ref as R
    g, P = run[0], run[1]
    v, e = R.skeleton(P)
    o = P.outputs(v)
    H = P.num_hidden
    steps = [(l, h) for l in range(P.num_layers - 1) for h in range(H)] + [(P.num_layers - 2, H)]
    for i, (l, h) in enumerate(steps):
        v, e, o = R.subpoly_step(P, v, e, o, l, h, 1e-4)
        assert (v.shape[0], e.shape[0]) == tuple(g["step_sizes"][i]), (l, h)


def test_surface_mesh_matches_reference(run):
    g, _, faces, v, tri, inter = run
    assert np.array_equal(inter["surface_edges"], g["surface_edges"].astype(np.int64))
    assert np.abs(v - g["surface_vertices"]).max() <= 1e-5
    ours, ref = canonical_polygons(inter["polygons"]), canonical_polygons(g["polygons"])
    assert set(ours) == set(ref)
    assert len(ours) == len(set(ours))
    if len(ref) == len(set(ref)):
        assert tri.shape[0] == g["triangles"].shape[0]
    assert np.array_equal(faces, v[tri])


def test_mesh_is_a_closed_surface(run):
    # size-independent property: every polygon edge is shared by exactly two polygons
    _, _, _, _, _, inter = run
    from collections import Counter
    c = Counter()
    for poly in set(canonical_polygons(inter["polygons"])):
        for a, b in zip(poly, poly[1:] + poly[:1]):
            c[(min(a, b), max(a, b))] += 1
    counts = Counter(c.values())
    assert counts[2] >= 0.95 * sum(counts.values())  # the reference mesh itself is ~98.8% manifold


def test_empty_and_degenerate_inputs():
    from helpers import random_net
    from oracle import subpoly_ref as R
    P = random_net(3)
    assert P.outputs(np.zeros((0, 3), np.float32)).shape == (0, P.n_outputs)
    m, off, _ = P.region(np.zeros((0, 3), np.float32))
    assert m.shape == (0, 3 + P.n_outputs) and off.shape == (0, 3)
    v, e = R.get_hypercube(1.2)
    assert v.shape == (8, 3) and e.shape == (12, 2)
    r_idx, org = R.regions_to_vertices(np.zeros((0, 5), np.int64), np.zeros((0, 3), np.int64))
    assert r_idx.size == 0 and org.size == 0


def test_deterministic_tanh_is_close_to_libm():
    from oracle.trinet import lib
    xs = np.concatenate([np.linspace(-12, 12, 4001), np.linspace(-0.3, 0.3, 2001)]).astype(np.float32)
    got = np.array([lib().det_tanhf(float(x)) for x in xs], np.float32)
    assert np.abs(got - np.tanh(xs.astype(np.float64))).max() < 2.5e-7


@pytest.mark.parametrize("case", CASES)
def test_curve_path_oracle_matches_reference(case):
    """force=False: the oracle's surface mesh against the reference's (fixture): same size,
    max nearest-vertex error 1e-5, Chamfer 1e-6 (BASELINE.json tolerances)."""
    import os
    from scipy.spatial import cKDTree
    from helpers import GOLDEN
    from oracle import subpoly_ref as R
    g = load_golden(case)
    gc = dict(np.load(os.path.join(GOLDEN, f"{case}_curve.npz")))
    faces, v, tri = R.subpoly(oracle_net(g), force=False)
    ref_v = gc["surface_vertices"]
    assert v.shape[0] == ref_v.shape[0]
    # the reference can emit a face twice (non-stable argsort before unique(dim=0)): a few extra triangles
    assert 0 <= gc["triangles"].shape[0] - tri.shape[0] <= 4
    d1, _ = cKDTree(ref_v).query(v)
    d2, _ = cKDTree(v).query(ref_v)
    assert d1.max() <= 1e-5 and d2.max() <= 1e-5
    assert (d1.mean() + d2.mean()) / 2 <= 1e-6


def test_curve_path_oracle_equals_reproducible_reference_on_the_medium_torus():
    """BASELINE configs[2] (medium model, torus, force=False) against `medium_torus_detlin`: the reference run
    with row-position-independent arithmetic (tests/golden/make_golden_medium.py --detlin).  Every hyperplane's
    (V, E) and the final edge array must be the reference's, bit for bit: this pins the whole curve path,
    including WHICH root of the intersection polynomial wins (geometry.py:259-300: the last admissible
    eigenvalue in LAPACK's order = the largest root in [0,1], oracle/trinet_ref.c last_root01)."""
    from oracle import subpoly_ref as R
    g = load_golden("medium_torus")
    g["net_table"] = g["net_table"].astype(np.float32)
    det = load_golden("medium_torus_detlin")
    P = oracle_net(g)
    H = P.num_hidden
    v, e = R.skeleton(P)
    out = P.outputs(v)
    steps = [(l, h) for l in range(P.num_layers - 1) for h in range(H)] + [(P.num_layers - 2, H)]
    for i, (l, h) in enumerate(steps):
        v, e, out = R.subpoly_step(P, v, e, out, l, h, 1e-4, force=False)
        assert (v.shape[0], e.shape[0]) == tuple(det["curve_step_sizes"][i]), (l, h)
    assert np.array_equal(e, det["curve_complex_edges"].astype(np.int64))
