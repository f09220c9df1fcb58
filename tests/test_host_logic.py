"""CPU: host-side mirror of the reference package (no kernels involved)."""
import numpy as np
import torch


def test_torch_ext_helpers_match_loops():
    import tropical  # noqa: F401  (installs torch.ext)
    t = (torch.rand(50, 9) > 0.7).long()
    t[3] = 0
    last = torch.ext.nonzero_last(t)
    first = torch.ext.nonzero_first(t)
    rows = [r for r in range(50) if t[r].any()]
    assert last[:, 0].tolist() == rows and first[:, 0].tolist() == rows
    for (r, c), (_, c0) in zip(last.tolist(), first.tolist()):
        nz = t[r].nonzero()[:, 0]
        assert c == int(nz[-1]) and c0 == int(nz[0])
    x = torch.arange(24).view(2, 4, 3)
    idx = torch.tensor([[3, 0], [1, 1]])
    assert torch.equal(torch.ext.batched_index_select(x, 1, idx)[0, 0], x[0, 3])
    u = torch.ext.batched_unique_consecutive(torch.tensor([[1, 1, 2], [3, 4, 5]]))
    assert u.tolist() == [[1, 2, -1], [3, 4, 5]]


def test_hashgrid_host_methods():
    from tropical import TropicalHashGrid
    g = TropicalHashGrid(1.0, 3, 4, 2, 19, 2, 32)
    M = len(g.marks)
    assert g.marks[0] == 0 and g.marks[-1] == 1 and bool((g.marks[1:] > g.marks[:-1]).all())
    idx = torch.tensor([[0, 0, 0], [1, 2, 3], [M - 1, M - 1, M - 1]])
    v = g.p2v(idx)
    assert v.tolist() == [0, M * M + 2 * M + 3, M ** 3 - 1]
    assert torch.equal(g.v2p(v), idx)
    x = torch.stack([g.marks[3], (g.marks[3] + g.marks[4]) / 2, torch.tensor(0.999999)]).view(1, 3)
    mask, off = g.region(x.repeat(1, 1))
    assert mask.tolist() == [[0, 1, 0]] and off.tolist()[0][:2] == [3, 3]


def test_net_layout_matches_reference_checkpoints():
    from tropical.stanford.model import Net
    net = Net(num_layers=3, num_hidden=16, levels=4, r_min=8, r_max=128, T=19)
    keys = list(net.state_dict().keys())
    assert keys == ["enc.module.params", "fc.0.weight", "fc.0.bias", "fc.1.weight", "fc.1.bias",
                    "fc.2.weight", "fc.2.bias"]
    assert net.enc.module.params.numel() == 2 * (512 + 9264 + 132656 + 524288)
    assert [tuple(fc.weight.shape) for fc in net.fc] == [(16, 8), (16, 16), (2, 16)]
    x = torch.rand(5, 3)
    assert torch.allclose(net.preprocess_inverse(net.preprocess(x)), x, atol=1e-6)


def test_hypercube():
    from tropical.subpoly import get_hypercube
    v, e, f = get_hypercube(3, 1.2)
    assert v.shape == (8, 3) and e.shape == (12, 2) and len(f) == 6
    from oracle import subpoly_ref as R
    vo, eo = R.get_hypercube(1.2)
    assert np.array_equal(v.numpy(), vo) and np.array_equal(e.numpy(), eo)


def test_training_route_restatement_matches_the_oracle_encoding_and_has_no_cpu_route():
    """`forward_autograd` (the plain-torch restatement the training kernels are tested against on
    the GPU, tests/test_gpu_train.py) is the oracle's interpolation; the product route itself
    refuses CPU tensors."""
    import pytest
    from helpers import load_golden, oracle_net
    from tropical import _native
    from tropical.stanford.model import Net
    g = load_golden("small_sphere")
    P = oracle_net(g)
    net = Net()
    net.enc.module.params.data = torch.from_numpy(g["net_table"])
    xp = torch.rand(500, 3)
    enc = net.enc.module.forward_autograd(xp.clone().requires_grad_(True))
    assert np.abs(enc.detach().numpy() - P.encode(xp.numpy())).max() <= 1e-6
    x = (torch.rand(64, 3) * 2 - 1).requires_grad_(True)
    with pytest.raises(_native.NativeError):
        net.sdf(x)


def test_train_entry_point_cli_matches_the_reference():
    from tropical.stanford import train
    a = train.parse(["-d", "bunny", "-s", "1", "-m", "large", "-e"])
    assert (a.dataset, a.seed, a.model_size, a.eval, a.force, a.cache) == ("bunny", 1, "large", True, True, True)
    a = train.parse(["-f", "-c"])   # both switches are store_false in the reference (train.py:44-51)
    assert a.force is False and a.cache is False and a.dataset == "dragon" and a.seed == 45


def test_training_grid_description_matches_the_parameter_layout():
    """tnb_grid_desc (include/tropical_b200.h) as HashEncoding fills it: the library derives the same
    table length as the tiny-cuda-nn layout the parameters were allocated with (no device needed)."""
    import ctypes
    from tropical import _native
    from tropical.stanford.model import Net
    for r_min, r_max, T in ((2, 32, 19), (4, 64, 19), (8, 128, 19), (8, 128, 21)):
        net = Net(num_layers=3, num_hidden=16, levels=4, r_min=r_min, r_max=r_max, T=T)
        d = net.enc.module.grid_desc()
        assert _native.lib().tnb_grid_train_table_len(ctypes.byref(d)) == net.enc.module.params.numel()
    bad = _native.GridDesc(0, 19, 2, 2.0)
    assert _native.lib().tnb_grid_train_table_len(ctypes.byref(bad)) == -1
