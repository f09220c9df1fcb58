"""CPU, build container only: the oracle against the reference's own Python, imported from
/root/reference through tests/golden/refenv.py.  Skipped where the reference tree is absent
(the GPU box)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import refenv  # noqa: E402

pytestmark = pytest.mark.skipif(not refenv.available(), reason="reference tree not present")


@pytest.fixture(scope="module", autouse=True)
def _leave_no_reference_modules_behind():
    yield
    refenv.release_reference()


@pytest.fixture(scope="module")
def ref():
    import torch
    tropical, sp, Net = refenv.import_reference()
    torch.manual_seed(11)
    net = Net(num_layers=3, num_hidden=16, levels=4, r_min=2, r_max=32, T=19)
    with torch.no_grad():
        net.enc.module.params.uniform_(-0.5, 0.5)
    from oracle.trinet import NetParams
    yield net, NetParams.from_reference_net(net), sp
    for k in [k for k in sys.modules if k == "tropical" or k.startswith("tropical.")]:
        del sys.modules[k]  # the product mirror is also called `tropical`


def test_network_rows_and_regions(ref):
    import torch
    net, P, _ = ref
    x = torch.rand(4000, 3) * 2.4 - 1.2
    with torch.no_grad():
        rows = torch.cat(net(x, gather=True)[1], -1).numpy()
        enc = net.enc(net.preprocess(x)).numpy()
    assert np.abs(P.encode(P.preprocess(x.numpy())) - enc).max() <= 5e-6
    assert np.abs(P.outputs(x.numpy()) - rows).max() <= 2e-5
    m_ref, off_ref, _ = net.region(x)
    m, off, _ = P.region(x.numpy(), rows)   # same rows in: the indicator logic itself
    assert np.array_equal(m, m_ref.numpy()) and np.array_equal(off, off_ref.numpy())


def test_sdf_and_gradient(ref):
    import torch
    net, P, _ = ref
    x = (torch.rand(2000, 3) * 2 - 1).requires_grad_(True)
    s = net.sdf(x)[:, 0]
    g = torch.autograd.grad(s.sum(), x)[0]
    so, go = P.sdf_grad(x.detach().numpy())
    assert np.abs(so - s.detach().numpy()).max() <= 1e-6
    assert np.abs(go - g.numpy()).max() <= 1e-4 * max(1.0, float(g.abs().max()))


def test_region_grouping_and_connecting_edges(ref):
    # regions_to_vertices / edge_vertices on the reference's own inputs
    import torch
    from oracle import subpoly_ref as R
    net, P, sp = ref
    rng = np.random.default_rng(0)
    m = rng.integers(-1, 2, size=(300, 9)).astype(np.int64)
    m[:, :3] = rng.integers(0, 2, size=(300, 3))
    off = rng.integers(0, 4, size=(300, 3)).astype(np.int64)
    m[np.arange(300), rng.integers(3, 9, size=300)] = 0  # the reference needs >= 1 zero per row (cartesian_prod)
    r_ref, o_ref = sp.regions_to_vertices(torch.from_numpy(m), torch.from_numpy(off), return_inverse=True)
    r, o = R.regions_to_vertices(m, off)
    assert np.array_equal(r, r_ref.numpy()) and np.array_equal(o, o_ref.numpy())
    pairs_ref = sp.edge_vertices(None, torch.from_numpy(m), net, torch.from_numpy(off)).numpy()
    pairs = R.edge_vertices(m, off)
    # subpoly_ sorts each pair and takes unique(dim=0) next (subpoly.py:243-244)
    canon = lambda p: sorted(set(map(tuple, np.sort(p, 1).tolist())))  # noqa: E731
    assert canon(pairs) == canon(pairs_ref)


def test_marks_mirror_equals_reference(ref):
    net, P, _ = ref
    for k in [k for k in sys.modules if k == "tropical" or k.startswith("tropical.")]:
        del sys.modules[k]
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tropical-nerf.pytorch_b200")
    saved = list(sys.path)
    sys.path.insert(0, pkg)
    try:
        from tropical.stanford.model import Net as MirrorNet
        for kw in (dict(r_min=2, r_max=32), dict(r_min=4, r_max=64), dict(r_min=8, r_max=128)):
            mirror = MirrorNet(num_layers=3, num_hidden=16, levels=4, T=19, **kw)
            refenv.import_reference()
            from tropical.stanford.model import Net as RefNet
            r = RefNet(num_layers=3, num_hidden=16, levels=4, T=19, **kw)
            assert np.array_equal(mirror.enc.marks.numpy(), r.enc.marks.numpy())
            assert mirror.enc.module.params.numel() == r.enc.module.params.numel()
            for k in [k for k in sys.modules if k == "tropical" or k.startswith("tropical.")]:
                del sys.modules[k]
    finally:
        sys.path[:] = saved
