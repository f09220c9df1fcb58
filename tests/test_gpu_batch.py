"""GPU: tnb_subpoly_batch (many objects in one call, several in flight on worker streams of the library) returns,
object by object, exactly the arrays of the single-object call -- which the parity tests pin to the oracle -- and
one of them is checked against the oracle here as well."""
import numpy as np
import pytest
import torch

from helpers import load_golden, native_net, oracle_net

pytestmark = pytest.mark.gpu


def _arrays(mesh):
    return [a.cpu().numpy() for a in mesh.read()]


def test_batch_equals_single_calls_and_oracle():
    from oracle import subpoly_ref as R
    from tropical import _native
    names = ["tiny_sphere_h8", "small_sphere", "small_torus"]
    P = {n: oracle_net(load_golden(n)) for n in names}
    N = {n: native_net(P[n]) for n in names}
    order = ["small_sphere", "tiny_sphere_h8", "small_torus", "small_sphere", "small_torus", "tiny_sphere_h8",
             "small_sphere", "small_sphere", "small_torus", "small_sphere", "tiny_sphere_h8"]
    want = {n: _arrays(N[n].subpoly(size=1.2, eps=1e-4, force=True)) for n in names}
    for in_flight in (1, 3, 8):
        meshes = _native.subpoly_batch([N[n] for n in order], size=1.2, eps=1e-4, force=True, in_flight=in_flight)
        torch.cuda.synchronize()
        assert len(meshes) == len(order)
        for n, m in zip(order, meshes):
            for a, b in zip(_arrays(m), want[n]):
                assert np.array_equal(a, b), (n, in_flight)
    faces, vo, tri = R.subpoly(P["tiny_sphere_h8"])
    v, e, t, f, p = want["tiny_sphere_h8"]
    assert np.array_equal(v, vo) and np.array_equal(t, tri) and np.array_equal(f, faces)


def test_batch_curve_path_and_repeat():
    from tropical import _native
    N = native_net(oracle_net(load_golden("small_torus")))
    want = _arrays(N.subpoly(size=1.2, eps=1e-4, force=False))
    for _ in range(3):   # the workers and their cached blocks are reused from call to call
        meshes = _native.subpoly_batch([N] * 5, size=1.2, eps=1e-4, force=False, in_flight=4)
        for m in meshes:
            for a, b in zip(_arrays(m), want):
                assert np.array_equal(a, b)
    _native.lib().tnb_release_cached_blocks()
    meshes = _native.subpoly_batch([N] * 2, size=1.2, eps=1e-4, force=False)
    for a, b in zip(_arrays(meshes[1]), want):
        assert np.array_equal(a, b)


def test_batch_reports_the_failing_object():
    from tropical import _native
    P = oracle_net(load_golden("small_torus"))
    N = native_net(P)
    # eps = 1e-5 on the curve path: the gradient-descent repair leaves an intersection off its planes (test_repair.py)
    with pytest.raises(_native.NativeError) as ex:
        _native.subpoly_batch([N, N], size=1.2, eps=1e-5, force=False, in_flight=2)
    assert "object" in str(ex.value) and "gradient-descent" in str(ex.value)
    assert _native.subpoly_batch([], size=1.2) == []


def test_mirror_subpoly_batch():
    from tropical import subpoly as sp

    class Shim:
        def __init__(self, n):
            self._n = n

        def native(self):
            return self._n
    N = native_net(oracle_net(load_golden("tiny_sphere_h8")))
    f1, v1, t1 = sp.subpoly(Shim(N), 3, 1.2, force=True)
    res = sp.subpoly_batch([Shim(N), Shim(N)], 3, 1.2, force=True)
    assert len(res) == 2
    for f, v, t in res:
        assert np.array_equal(f, f1) and np.array_equal(t, t1) and torch.equal(v, v1)
