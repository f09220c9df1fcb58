"""GPU: the stage-level API of the `tropical` mirror (tropical.geometry, the stage functions of tropical.subpoly,
Net.forward(group=8)) against the oracle, bit for bit, through the C ABI entry points they bind."""
import numpy as np
import pytest
import torch

from helpers import load_golden, native_net, oracle_net

pytestmark = pytest.mark.gpu


class _NetShim:
    """What tropical.subpoly's stage functions need of a Net: .native() and .eps."""

    def __init__(self, N, eps):
        self._n, self.eps = N, eps

    def native(self):
        return self._n


def _curved_inputs(P, n=4000, seed=3):
    """Edges with two or three moving axes inside one marks-grid cell, as the curve path meets them."""
    from oracle import subpoly_ref as R
    rng = np.random.default_rng(seed)
    m = P.marks
    i = rng.integers(0, len(m) - 1, size=(n, 3))
    lo, hi = P.preprocess_inverse(m[i]), P.preprocess_inverse(m[i + 1])
    t0, t1 = rng.random((n, 3), dtype=np.float32), rng.random((n, 3), dtype=np.float32)
    e0, e1 = lo + (hi - lo) * t0, lo + (hi - lo) * t1
    flat = rng.integers(0, 3, size=n)           # a third of the edges keep one coordinate: degenerate boxes
    keep = rng.random(n) < 0.35
    e1[keep, flat[keep]] = e0[keep, flat[keep]]
    e = np.stack([e0, e1], 1).astype(np.float32)
    return e, R.corner_points(e)


def test_corner_points_and_group8_rows_bit_exact():
    from tropical import geometry as gm
    g = load_golden("small_torus")
    P = oracle_net(g)
    N = native_net(P)
    e, corners = _curved_inputs(P)
    got = gm.corner_points(torch.from_numpy(e).cuda())
    assert np.array_equal(got.cpu().numpy(), corners)
    rows, raw = N.outputs_group8(got.reshape(-1, 3))
    want = P.outputs_group8(corners)
    assert np.array_equal(rows.cpu().numpy().reshape(want.shape), want)
    assert np.array_equal((raw[:, 1] - raw[:, 0]).cpu().numpy(), want.reshape(-1, want.shape[-1])[:, -1])
    with pytest.raises(Exception):
        N.outputs_group8(got.reshape(-1, 3)[:9])


def test_forward_returns_rows_and_raw_output():
    g = load_golden("small_sphere")
    P = oracle_net(g)
    N = native_net(P)
    x = torch.rand(5000, 3, device="cuda") * 2.4 - 1.2
    rows, raw = N.forward(x)
    assert np.array_equal(rows.cpu().numpy(), P.outputs(x.cpu().numpy()))
    assert np.array_equal((raw[:, 1] - raw[:, 0]).cpu().numpy(), rows[:, -1].cpu().numpy())
    none, raw2 = N.forward(x, rows=False)
    assert none is None and torch.equal(raw, raw2)


def test_intersection_of_two_planes_bit_exact():
    from oracle.trinet import curve_intersections
    from tropical import geometry as gm
    g = load_golden("small_torus")
    P = oracle_net(g)
    e, corners = _curved_inputs(P, n=6000, seed=5)
    dd = P.outputs_group8(corners)                          # [E,8,R]: real corner values, degenerate boxes among them
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, dd.shape[2], size=2 * dd.shape[0]).reshape(2, -1)
    rows = np.arange(dd.shape[0])
    p, q = dd[rows, :, a], dd[rows, :, b]
    want = curve_intersections(p, q)
    got = gm.intersection_of_two_planes(torch.from_numpy(p).cuda(), torch.from_numpy(q).cuda()).cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True)
    assert (want[:, 0] >= 0).sum() > 100 and (want[:, 0] == -1).sum() > 100   # roots found / none (degenerate boxes among them)


def test_sort_polygon_vertices_batch_matches_the_oracle_order():
    from oracle import subpoly_ref as R
    from tropical import geometry as gm
    rng = np.random.default_rng(1)
    B, M = 500, 9
    centre = rng.normal(size=(B, 1, 3)).astype(np.float32)
    v = (centre + 0.1 * rng.normal(size=(B, M, 3))).astype(np.float32)
    cnt = rng.integers(3, M + 1, size=B)
    v[np.arange(M)[None, :] >= cnt[:, None]] = 0            # right-padded rows, as r_idx_as_tensor builds them
    n = rng.normal(size=(B, 3)).astype(np.float32)
    order, valid = R.polygon_order(v, n)
    faces_want = R.fan_triangles(np.take_along_axis(v, order[..., None], 1), np.take_along_axis(valid, order, 1))
    faces, idx = gm.sort_polygon_vertices_batch(torch.from_numpy(v).cuda(), torch.from_numpy(n).cuda(), return_index=True)
    assert np.array_equal(idx.cpu().numpy(), order)
    assert np.array_equal(faces, faces_want)


def test_extract_skeleton_and_extract_faces_stage_functions():
    from oracle import subpoly_ref as R
    from tropical import subpoly as sp
    g = load_golden("small_sphere")
    P = oracle_net(g)
    N = native_net(P)
    net = _NetShim(N, 1e-4)
    H = P.num_hidden
    c = N.skeleton(128)
    c.steps([(l, h) for l in range(P.num_layers - 1) for h in range(H)] + [(P.num_layers - 2, H)])
    v, e, o = c.read()
    sv, se, v_idx = sp.extract_skeleton(v, e, net, 1e-4, outputs=o)
    vo, eo, idx_o = R.extract_skeleton(P, v.cpu().numpy(), e.cpu().numpy(), o.cpu().numpy(), 1e-4)
    assert np.array_equal(sv.cpu().numpy(), vo) and np.array_equal(se.cpu().numpy(), eo)
    assert np.array_equal(v_idx.cpu().numpy(), idx_o)
    faces, tri = sp.extract_faces(sv, se, net, outputs=o[v_idx], eps=1e-4)
    want = [a.cpu().numpy() for a in N.subpoly().read()]
    assert np.array_equal(tri, want[2]) and np.array_equal(faces, want[3])
