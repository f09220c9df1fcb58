"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact for
every integer array and for every float the fixed operation order defines; against the
reference's golden fixtures within the tolerances BASELINE.json states (positions 1e-5)."""
import numpy as np
import pytest
import torch

from helpers import CASES, canonical_polygons, load_golden, native_net, oracle_net, random_net

pytestmark = pytest.mark.gpu


def _pts(n, seed, lo=-1.0, hi=1.0):
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


@pytest.mark.parametrize("case", CASES)
def test_network_evaluation_bit_exact(case):
    g = load_golden(case)
    P = oracle_net(g)
    N = native_net(P)
    x = _pts(20000, 1, -1.2, 1.2)  # includes points outside the grid (uint32 wrap path)
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(N.encode(torch.from_numpy(P.preprocess(x)).cuda()).cpu().numpy(),
                          P.encode(P.preprocess(x)))
    assert np.array_equal(N.outputs(xd).cpu().numpy(), P.outputs(x))
    sdf, grad = N.sdf_grad(xd)
    so, go = P.sdf_grad(x)
    assert np.array_equal(sdf.cpu().numpy(), so)
    assert np.array_equal(grad.cpu().numpy(), go)
    signs, off, _, packed = N.region(xd, packed=True)
    m, o, _ = P.region(x)
    assert np.array_equal(signs.cpu().numpy(), m)
    assert np.array_equal(off.cpu().numpy(), o)


def test_generic_shapes_bit_exact():
    # shapes outside the compiled-in (4,16,3) configuration go through the runtime-sized kernels
    for seed, kw in enumerate([dict(levels=2, num_hidden=8), dict(levels=6, num_hidden=12, num_layers=4),
                               dict(levels=4, num_hidden=16, log2_T=10, n_max=64)]):
        P = random_net(seed, **kw)
        N = native_net(P)
        x = _pts(5000, seed)
        xd = torch.from_numpy(x).cuda()
        assert np.array_equal(N.outputs(xd).cpu().numpy(), P.outputs(x)), kw
        sdf, grad = N.sdf_grad(xd)
        so, go = P.sdf_grad(x)
        assert np.array_equal(sdf.cpu().numpy(), so) and np.array_equal(grad.cpu().numpy(), go), kw


@pytest.mark.parametrize("case", CASES)
def test_skeleton_matches_oracle_and_reference(case):
    from oracle import subpoly_ref as R
    g = load_golden(case)
    P = oracle_net(g)
    N = native_net(P)
    c = N.skeleton(128)
    v, e, o = c.read()
    vo, eo = R.skeleton(P)
    assert np.array_equal(e.cpu().numpy(), eo)
    assert np.array_equal(v.cpu().numpy(), vo)
    assert np.array_equal(o.cpu().numpy(), P.outputs(vo))
    # the reference's own skeleton
    assert np.array_equal(e.cpu().numpy(), g["skeleton_edges"].astype(np.int64))
    assert np.abs(v.cpu().numpy() - g["skeleton_vertices"]).max() <= 1e-6


@pytest.mark.parametrize("case", CASES)
def test_every_hyperplane_step_bit_exact(case):
    from oracle import subpoly_ref as R
    g = load_golden(case)
    P = oracle_net(g)
    N = native_net(P)
    c = N.skeleton(128)
    vo, eo = R.skeleton(P)
    oo = P.outputs(vo)
    H = P.num_hidden
    steps = [(l, h) for l in range(P.num_layers - 1) for h in range(H)] + [(P.num_layers - 2, H)]
    for i, (l, h) in enumerate(steps):
        c.step(l, h)
        vo, eo, oo = R.subpoly_step(P, vo, eo, oo, l, h, 1e-4)
        assert (c.num_vertices, c.num_edges) == (vo.shape[0], eo.shape[0]), (l, h)
        assert (c.num_vertices, c.num_edges) == tuple(g["step_sizes"][i]), (l, h)
        v, e, o = c.read()
        assert np.array_equal(e.cpu().numpy(), eo), (l, h)
        assert np.array_equal(v.cpu().numpy(), vo), (l, h)
        assert np.array_equal(o.cpu().numpy(), oo), (l, h)
    assert np.array_equal(eo, g["complex_edges"].astype(np.int64))
    assert np.abs(vo - g["complex_vertices"]).max() <= 1e-5


@pytest.mark.parametrize("case", CASES)
def test_whole_path_mesh(case):
    from oracle import subpoly_ref as R
    g = load_golden(case)
    P = oracle_net(g)
    N = native_net(P)
    mesh = N.subpoly()
    v, e, t, f, p = [a.cpu().numpy() for a in mesh.read()]
    faces, vo, tri, inter = R.subpoly(P, return_intermediate=True)
    # against the oracle: identical arrays
    assert np.array_equal(v, vo)
    assert np.array_equal(e, inter["surface_edges"])
    assert np.array_equal(t, tri)
    assert np.array_equal(f, faces)
    # polygon rows: the oracle keeps torch's interleaved -1 padding, the device rows are
    # left-packed; the vertex sequence of every row must be identical
    op = inter["polygons"]
    assert p.shape[0] == op.shape[0]
    packed = np.full_like(p, -1)
    for r, row in enumerate(op):
        keep = row[row != -1]
        packed[r, :keep.size] = keep
    assert np.array_equal(p, packed)
    # against the reference: same surface skeleton, same polygons, positions within 1e-5
    assert np.array_equal(e, g["surface_edges"].astype(np.int64))
    assert np.abs(v - g["surface_vertices"]).max() <= 1e-5
    # The reference's rows come out of a non-stable argsort (subpoly.py:357), so a face whose
    # vertices all lie on one more common plane can survive its unique(dim=0) twice in two
    # orders; as a set of cyclic vertex sequences the faces must be identical.
    ours, ref = canonical_polygons(p), canonical_polygons(g["polygons"])
    assert set(ours) == set(ref)
    assert len(ours) == len(set(ours))  # the device path never emits a face twice
    if len(ref) == len(set(ref)):
        assert t.shape[0] == g["triangles"].shape[0]
    # host-buffer read (the e2e entry) agrees with the device read
    hv, ht, hf, hp = mesh.read_host()
    assert np.array_equal(hv, v) and np.array_equal(ht, t) and np.array_equal(hf, f) and np.array_equal(hp, p)


def test_no_surface_falls_back_to_hypercube():
    # a network whose SDF never crosses zero: the skeleton is empty and subpoly starts from
    # the hypercube (subpoly.py:51-52)
    from oracle import subpoly_ref as R
    P = random_net(7, table_amp=1e-4)
    P.biases[-1][:] = [0.0, 5.0]
    P = type(P)(P.levels, 2, P.log2_T, P.n_min, P.per_level_scale, P.num_layers, P.num_hidden,
                P.table, P.weights, P.biases, P.marks)
    N = native_net(P)
    vo, eo = R.skeleton(P)
    assert eo.shape[0] == 0
    c = N.skeleton(128, 1.2)
    v, e, _ = c.read()
    hv, he = R.get_hypercube(1.2)
    assert np.array_equal(v.cpu().numpy(), hv) and np.array_equal(e.cpu().numpy(), he)
    mesh = N.subpoly()
    assert mesh.sizes()["T"] == R.subpoly(P)[2].shape[0]


def test_sweep_signs_matches_region():
    g = load_golden("small_sphere")
    P = oracle_net(g)
    N = native_net(P)
    n = (17, 9, 33)
    lo, hi = (-1.0, -0.5, -1.0), (1.0, 0.5, 1.0)
    packed = N.sweep_signs(lo, hi, n).cpu().numpy().view(np.uint64)
    ax = [np.float32(l) + np.arange(k, dtype=np.float32) * (np.float32(h - l) / np.float32(k - 1))
          for l, h, k in zip(lo, hi, n)]
    # lattice points as the kernel forms them: fma(i, step, lo)
    # output order: x fastest, then y, then z
    zyx = np.stack(np.meshgrid(*[np.arange(k) for k in n[::-1]], indexing="ij"), -1).reshape(-1, 3)
    pts = zyx[:, ::-1]
    step = [np.float32(h - l) / np.float32(k - 1) for l, h, k in zip(lo, hi, n)]
    x = np.stack([(pts[:, d].astype(np.float64) * np.float64(step[d]) + np.float64(lo[d])).astype(np.float32)
                  for d in range(3)], -1)
    m, _, _ = P.region(x)
    s = m[:, 3:]
    pos = np.zeros(len(x), np.uint64)
    neg = np.zeros(len(x), np.uint64)
    for c in range(s.shape[1]):
        pos |= (s[:, c] == 1).astype(np.uint64) << np.uint64(c)
        neg |= (s[:, c] == -1).astype(np.uint64) << np.uint64(c)
    assert np.array_equal(packed[:, 0], pos) and np.array_equal(packed[:, 1], neg)


@pytest.mark.parametrize("case", CASES)
def test_curve_path_every_step_bit_exact(case):
    """force=False (curve approximation): device == oracle at every hyperplane."""
    from oracle import subpoly_ref as R
    g = load_golden(case)
    P = oracle_net(g)
    N = native_net(P)
    c = N.skeleton(128)
    vo, eo = R.skeleton(P)
    oo = P.outputs(vo)
    H = P.num_hidden
    for (l, h) in [(l, h) for l in range(P.num_layers - 1) for h in range(H)] + [(P.num_layers - 2, H)]:
        c.step(l, h, force=False)
        vo, eo, oo = R.subpoly_step(P, vo, eo, oo, l, h, 1e-4, force=False)
        assert (c.num_vertices, c.num_edges) == (vo.shape[0], eo.shape[0]), (l, h)
        v, e, o = c.read()
        assert np.array_equal(e.cpu().numpy(), eo), (l, h)
        assert np.array_equal(v.cpu().numpy(), vo), (l, h)
        assert np.array_equal(o.cpu().numpy(), oo), (l, h)


@pytest.mark.parametrize("case", CASES)
def test_curve_path_mesh_matches_reference(case):
    """force=False whole path: identical to the oracle; against the reference's mesh the
    tolerances of BASELINE.json (max nearest-vertex error 1e-5, Chamfer 1e-6).  The float32
    eigenvalue root finder of the reference decides near-degenerate in-plane edges by rounding
    noise, so the complex away from the surface may differ in a handful of vertices."""
    from scipy.spatial import cKDTree
    from oracle import subpoly_ref as R
    g = load_golden(case)
    gc = dict(np.load(__import__("os").path.join(__import__("helpers").GOLDEN, f"{case}_curve.npz")))
    P = oracle_net(g)
    N = native_net(P)
    mesh = N.subpoly(force=False)
    v, e, t, f, p = [a.cpu().numpy() for a in mesh.read()]
    faces, vo, tri = R.subpoly(P, force=False)
    assert np.array_equal(v, vo) and np.array_equal(t, tri) and np.array_equal(f, faces)
    ref_v = gc["surface_vertices"]
    d1, _ = cKDTree(ref_v).query(v)
    d2, _ = cKDTree(v).query(ref_v)
    assert d1.max() <= 1e-5 and d2.max() <= 1e-5
    assert (d1.mean() + d2.mean()) / 2 <= 1e-6
    assert v.shape[0] == ref_v.shape[0]
    assert 0 <= gc["triangles"].shape[0] - t.shape[0] <= 4  # duplicated reference faces, see test_whole_path_mesh


def test_train_entry_point_runs_end_to_end(tmp_path, monkeypatch):
    """python -m tropical.stanford.train -d sphere -e : fit, extract on the device, write the mesh."""
    from tropical.stanford import train
    monkeypatch.chdir(tmp_path)
    train.main(["-d", "sphere", "-s", "1", "-e"])   # the reference's 10 epochs x 50 batches; without -c the fitted network is saved
    ply = tmp_path / "meshes" / "sphere" / "our_mesh_small_1.ply"
    assert ply.exists()
    lines = ply.read_text().splitlines()
    n_v, n_f = int(lines[2].split()[-1]), int(lines[6].split()[-1])
    assert lines[0] == "ply" and n_v > 1000   # a real surface came out
    # ... and it is the mesh the oracle extracts from the network the run saved: same vertices, same triangles
    import os
    from oracle import subpoly_ref as R
    from oracle.trinet import NetParams
    from tropical.stanford.dataset import StanfordDataset
    from tropical.stanford.model import Net
    ckpt = os.path.join(os.path.dirname(train.__file__), "models", "sphere", "sphere_sdf_small_1.pth")
    try:
        net = Net(num_layers=3, num_hidden=16, levels=4, r_min=2, r_max=32, T=19)
        net.load_state_dict(torch.load(ckpt, map_location="cpu"))
        _, vo, to = R.subpoly(NetParams.from_reference_net(net), force=True)
    finally:
        if os.path.exists(ckpt):
            os.remove(ckpt)   # the next run trains again
    assert (n_v, n_f) == (vo.shape[0], to.shape[0])
    body = np.loadtxt(lines[9:9 + n_v], dtype=np.float64)
    assert np.abs(body * StanfordDataset("sphere").R - vo).max() <= 1e-6     # the .ply is text with 8 significant digits
    tri = np.loadtxt(lines[9 + n_v:9 + n_v + n_f], dtype=np.int64)[:, 1:]
    assert np.array_equal(tri, to)


def test_mirror_api_matches_native():
    """tropical.subpoly.subpoly / Net methods (the reference-facing API) on the device."""
    import tropical.subpoly as sp
    from tropical.stanford.model import Net
    g = load_golden("small_sphere")
    P = oracle_net(g)
    net = Net(num_layers=3, num_hidden=16, levels=4, r_min=2, r_max=32, T=19)
    sd = {"enc.module.params": torch.from_numpy(g["net_table"])}
    for i in range(3):
        sd[f"fc.{i}.weight"], sd[f"fc.{i}.bias"] = torch.from_numpy(g[f"net_w{i}"]), torch.from_numpy(g[f"net_b{i}"])
    net.load_state_dict(sd)
    net = net.cuda()
    with torch.no_grad():
        x = torch.rand(1000, 3, device="cuda") * 2 - 1
        out, inputs = net(x, gather=True)
        assert np.array_equal(torch.cat(inputs, -1).cpu().numpy(), P.outputs(x.cpu().numpy()))
        m, off, _ = net.region(x)
        mo, oo, _ = P.region(x.cpu().numpy())
        assert np.array_equal(m.cpu().numpy(), mo) and np.array_equal(off.cpu().numpy(), oo)
        assert np.array_equal(net.sdf(x)[:, 0].cpu().numpy(), P.sdf_grad(x.cpu().numpy(), False)[0])
        assert np.array_equal(net.normal(x).cpu().numpy(), P.sdf_grad(x.cpu().numpy())[1])
        v, e = net.enc.skeleton(net)
        assert np.array_equal(e.cpu().numpy(), g["skeleton_edges"].astype(np.int64))
        faces, vertices, tri = sp.subpoly(net, 3, 1.2, force=True)
    from oracle import subpoly_ref as R
    fo, vo, to = R.subpoly(P)
    assert np.array_equal(vertices.cpu().numpy(), vo) and np.array_equal(tri, to) and np.array_equal(faces, fo)


@pytest.mark.parametrize("unit", [9, 17, 25])
def test_chunked_skeleton_reproduces_the_overlap_duplicates(unit):
    """TropicalHashGrid.skeleton walks the marks grid in chunks of `unit` vertices that overlap by
    one plane (tropical.py:176-181): per-chunk thresholds, and grid edges on a chunk-boundary plane
    come out twice.  Small chunks exercise that path (the large model hits it with unit=128).
    unit=9 leaves clusters of 40 coincident vertices: partner lists longer than the cache, i.e. the
    warp-per-list write pass (pair_write_long) inside the persistent step kernel."""
    from oracle import subpoly_ref as R
    g = load_golden("small_sphere")
    P = oracle_net(g)
    N = native_net(P)
    c = N.skeleton(unit, 1.2)
    v, e, o = c.read()
    vo, eo = R.skeleton(P, unit)
    assert np.array_equal(e.cpu().numpy(), eo) and np.array_equal(v.cpu().numpy(), vo)
    e_np = e.cpu().numpy()
    assert len(np.unique(e_np, axis=0)) < len(e_np)   # the duplicated boundary edges are really there
    mesh = N.subpoly(unit=unit)
    vv, _, t, f, _ = [a.cpu().numpy() for a in mesh.read()]
    fo, vo2, to = R.subpoly(P, unit=unit)
    assert np.array_equal(vv, vo2) and np.array_equal(t, to) and np.array_equal(f, fo)


def test_work_buffers_grow_on_demand():
    """capacity factor 1.0: no head-room at all, every growing step takes the re-allocation path."""
    from oracle import subpoly_ref as R
    from tropical import _native
    g = load_golden("small_torus")
    P = oracle_net(g)
    N = native_net(P)
    _native.check(_native.lib().tnb_set_capacity_factor(1.0))
    try:
        mesh = N.subpoly()
        v, _, t, _, _ = [a.cpu().numpy() for a in mesh.read()]
    finally:
        _native.check(_native.lib().tnb_set_capacity_factor(4.0))
    _, vo, to = R.subpoly(P)
    assert np.array_equal(v, vo) and np.array_equal(t, to)


def test_empty_and_caller_supplied_inputs():
    from oracle import subpoly_ref as R
    g = load_golden("tiny_sphere_h8")
    P = oracle_net(g)
    N = native_net(P)
    empty = torch.empty((0, 3), device="cuda")
    assert N.outputs(empty).shape == (0, P.n_outputs)
    assert N.sdf_grad(empty)[0].shape == (0,)
    s, off, _ = N.region(empty)
    assert s.shape == (0, 3 + P.n_outputs) and off.shape == (0, 3)
    # a complex built from caller arrays (the reference passes its own vertices/edges to subpoly_)
    vo, eo = R.skeleton(P)
    c = N.complex_from_arrays(torch.from_numpy(vo).cuda(), torch.from_numpy(eo).cuda())
    oo = P.outputs(vo)
    for (l, h) in [(0, 0), (0, 1), (1, 3)]:
        c.step(l, h)
        vo, eo, oo = R.subpoly_step(P, vo, eo, oo, l, h, 1e-4)
        v, e, o = c.read()
        assert np.array_equal(e.cpu().numpy(), eo) and np.array_equal(v.cpu().numpy(), vo)
    # a complex without edges stays empty
    c0 = N.complex_from_arrays(torch.from_numpy(vo[:4]).cuda(), torch.empty((0, 2), dtype=torch.int64, device="cuda"))
    c0.step(0, 0)
    assert c0.num_edges == 0
    assert c0.extract_mesh().sizes()["T"] == 0


def _all_steps(P):
    H = P.num_hidden
    return [(l, h) for l in range(P.num_layers - 1) for h in range(H)] + [(P.num_layers - 2, H)]


@pytest.mark.parametrize("case", CASES)
def test_cluster_steps_equal_single_steps(case):
    """All hyperplanes in ONE persistent launch (tnb_subpoly_steps; here by a single thread-block
    cluster) leave exactly the complex that 33 separate tnb_subpoly_step calls (cooperative grid)
    leave, and the oracle's; prefixes of the step list too (the device-side step loop has no host
    in between)."""
    from oracle import subpoly_ref as R
    from tropical._native import lib
    g = load_golden(case)
    P = oracle_net(g)
    N = native_net(P)
    steps = _all_steps(P)
    before = lib().tnb_set_cluster_max_items(-1)
    try:
        for upto in (len(steps), 7):
            lib().tnb_set_cluster_max_items(0)          # one persistent cooperative-grid launch per step
            a = N.skeleton(128)
            for l, h in steps[:upto]:
                a.step(l, h)
            va, ea, oa = [t.cpu().numpy() for t in a.read()]
            lib().tnb_set_cluster_max_items(1 << 40)    # one cluster launch for the whole list
            b = N.skeleton(128)
            b.steps(steps[:upto])
            vb, eb, ob = [t.cpu().numpy() for t in b.read()]
            assert np.array_equal(ea, eb) and np.array_equal(va, vb) and np.array_equal(oa, ob), upto
    finally:
        lib().tnb_set_cluster_max_items(before)
    vo, eo = R.skeleton(P)
    oo = P.outputs(vo)
    for l, h in steps[:7]:
        vo, eo, oo = R.subpoly_step(P, vo, eo, oo, l, h, 1e-4)
    assert np.array_equal(eb, eo) and np.array_equal(vb, vo) and np.array_equal(ob, oo)


def test_whole_path_same_mesh_with_and_without_cluster_launch():
    from tropical._native import lib
    g = load_golden("small_sphere")
    N = native_net(oracle_net(g))
    before = lib().tnb_set_cluster_max_items(-1)
    try:
        lib().tnb_set_cluster_max_items(0)
        m0 = [a.cpu().numpy() for a in N.subpoly().read()]
        lib().tnb_set_cluster_max_items(1 << 40)
        m1 = [a.cpu().numpy() for a in N.subpoly().read()]
    finally:
        lib().tnb_set_cluster_max_items(before)
    for x, y in zip(m0, m1):
        assert np.array_equal(x, y)


def test_steps_with_an_eps_other_than_the_networks():
    """The no-op shortcut (crossing mask from the packed signs) only holds when the step's eps is the
    eps the signs were packed with; with another eps every step looks at the edges again."""
    from oracle import subpoly_ref as R
    g = load_golden("tiny_sphere_h8")
    P = oracle_net(g)
    N = native_net(P)
    steps = _all_steps(P)[:9]
    eps = 3e-4
    c = N.skeleton(128)
    c.steps(steps, eps=eps)
    v, e, o = [t.cpu().numpy() for t in c.read()]
    d = N.skeleton(128)
    for l, h in steps:
        d.step(l, h, eps=eps)
    v2, e2, o2 = [t.cpu().numpy() for t in d.read()]
    vo, eo = R.skeleton(P)
    oo = P.outputs(vo)
    for l, h in steps:
        vo, eo, oo = R.subpoly_step(P, vo, eo, oo, l, h, eps)
    assert np.array_equal(e, eo) and np.array_equal(v, vo) and np.array_equal(o, oo)
    assert np.array_equal(e2, eo) and np.array_equal(v2, vo) and np.array_equal(o2, oo)


@pytest.mark.parametrize("case", CASES)
def test_device_driven_step_stream_equals_single_steps(case):
    """The launch stream of LARGE complexes (tnb_subpoly_steps without the persistent kernel: the device picks
    the next crossing hyperplane, the host never synchronises), forced onto the small fixtures: the same
    complex as separate tnb_subpoly_step calls and as the oracle, for the whole list, for prefixes, and when
    the stream hands the rest of the list to the persistent kernel half-way."""
    from oracle import subpoly_ref as R
    from tropical._native import lib
    g = load_golden(case)
    P = oracle_net(g)
    N = native_net(P)
    steps = _all_steps(P)
    before = lib().tnb_set_fused_max_items(-1)
    try:
        a = N.skeleton(128)
        for l, h in steps:
            a.step(l, h)
        va, ea, oa = [t.cpu().numpy() for t in a.read()]
        lib().tnb_set_fused_max_items(0)            # never the persistent kernel
        b = N.skeleton(128)
        b.steps(steps)
        vb, eb, ob = [t.cpu().numpy() for t in b.read()]
        assert np.array_equal(ea, eb) and np.array_equal(va, vb) and np.array_equal(oa, ob)
        m_stream = [x.cpu().numpy() for x in N.subpoly().read()]
        p7 = N.skeleton(128)
        p7.steps(steps[:7])
        v7, e7, o7 = [t.cpu().numpy() for t in p7.read()]
        # hand-over: the stream starts (the skeleton is larger than the limit) and the persistent kernel finishes
        e0 = N.skeleton(128).num_edges
        lib().tnb_set_fused_max_items(int(e0))
        h = N.skeleton(128)
        h.steps(steps)
        vh, eh, oh = [t.cpu().numpy() for t in h.read()]
        assert np.array_equal(ea, eh) and np.array_equal(va, vh) and np.array_equal(oa, oh)
    finally:
        lib().tnb_set_fused_max_items(before)
    m_default = [x.cpu().numpy() for x in N.subpoly().read()]
    for x, y in zip(m_stream, m_default):
        assert np.array_equal(x, y)
    vo, eo = R.skeleton(P)
    oo = P.outputs(vo)
    for l, h in steps[:7]:
        vo, eo, oo = R.subpoly_step(P, vo, eo, oo, l, h, 1e-4)
    assert np.array_equal(e7, eo) and np.array_equal(v7, vo) and np.array_equal(o7, oo)


def test_device_driven_step_stream_reports_capacity():
    """Work arrays do not grow inside the stream: an overflow is latched and tnb_subpoly repeats the extraction
    with more head-room (same mesh in the end)."""
    from tropical import _native
    from tropical._native import lib
    g = load_golden("small_sphere")
    N = native_net(oracle_net(g))
    want = [x.cpu().numpy() for x in N.subpoly().read()]
    before = lib().tnb_set_fused_max_items(0)
    try:
        _native.check(lib().tnb_set_capacity_factor(1.0))
        try:
            c = N.skeleton(128)
            got = [x.cpu().numpy() for x in N.subpoly().read()]   # retries inside tnb_subpoly
        finally:
            _native.check(lib().tnb_set_capacity_factor(4.0))
        c.steps(_all_steps(oracle_net(g)))
        with pytest.raises(_native.NativeError, match="CAPACITY"):
            c.num_vertices
    finally:
        lib().tnb_set_fused_max_items(before)
    for x, y in zip(got, want):
        assert np.array_equal(x, y)
