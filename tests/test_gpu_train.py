"""GPU: the hash-grid training kernels (csrc/grid_train.cu: forward, backward, backward of the
backward) against the plain-torch restatement of the same interpolation
(`HashEncoding.forward_autograd`, itself checked against the oracle's encoding on the CPU in
tests/test_host_logic.py), and the training loss of stanford/train.py:186-201 through them.
fp32 with atomics in a different summation order: tolerance 2e-5 relative to the largest entry."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(a, b, what, tol=2e-5):
    scale = max(float(b.abs().max()), 1e-12)
    err = float((a - b).abs().max()) / scale
    assert err <= tol, f"{what}: relative error {err:.3e} > {tol}"


def _module(size, seed=0):
    from tropical.stanford.model import Net
    r_min, r_max = {"small": (2, 32), "medium": (4, 64), "large": (8, 128)}[size]
    torch.manual_seed(seed)
    net = Net(num_layers=3, num_hidden=16, levels=4, r_min=r_min, r_max=r_max, T=19).cuda()
    with torch.no_grad():
        net.enc.module.params.uniform_(-1.0, 1.0)  # amplitudes that make every term matter
    return net


@pytest.mark.parametrize("size", ["small", "large"])
def test_encoding_forward_backward_double_backward_match_torch(size):
    net = _module(size)
    enc = net.enc.module
    n = 4096
    torch.manual_seed(1)
    x0 = torch.rand(n, 3, device="cuda") * 1.1 - 0.05   # a few points outside [0,1]: index wrap paths
    w1 = torch.randn(n, 8, device="cuda")
    w2 = torch.randn(n, 3, device="cuda")

    def run(fn):
        enc.params.grad = None
        x = x0.clone().requires_grad_(True)
        y = fn(x)
        (dx,) = torch.autograd.grad((y * w1).sum(), x, create_graph=True)
        # a loss that uses the value, and the input gradient non-linearly (like the eikonal term)
        loss = (y ** 2).sum() + ((dx * w2).sum(-1) ** 2).sum() + (dx ** 2).sum()
        loss.backward()
        return y.detach(), dx.detach(), x.grad.detach(), enc.params.grad.detach().clone()

    ya, dxa, gxa, gpa = run(enc.forward_train)
    yb, dxb, gxb, gpb = run(enc.forward_autograd)
    _close(ya, yb, "encoding")
    _close(dxa, dxb, "d enc / d x")
    _close(gxa, gxb, "d loss / d x through the double backward")
    _close(gpa, gpb, "d loss / d table through the double backward")
    assert float(gpb.abs().max()) > 0 and float(gxb.abs().max()) > 0


def test_first_order_table_gradient_matches_torch():
    net = _module("medium")
    enc = net.enc.module
    x = torch.rand(3000, 3, device="cuda")
    w = torch.randn(3000, 8, device="cuda")
    out = []
    for fn in (enc.forward_train, enc.forward_autograd):
        enc.params.grad = None
        (fn(x) * w).sum().backward()
        out.append(enc.params.grad.detach().clone())
    _close(out[0], out[1], "d loss / d table")


def test_training_loss_of_the_reference_runs_through_the_kernels():
    """stanford/train.py:186-201: L1 on the clamped SDF + eikonal term + weight-norm term; the
    gradients of every parameter match the plain-torch route."""
    import torch.nn.functional as F
    net = _module("small", seed=3)
    pts0 = torch.rand(1000, 3, device="cuda") * 2 - 1
    labels = torch.randn(1000, device="cuda") * 0.1

    def loss_of(sdf_fn):
        net.zero_grad(set_to_none=True)
        pts = pts0.clone().requires_grad_(True)
        sdf = sdf_fn(pts)
        l1 = F.l1_loss(torch.clamp(sdf[:, 0], -0.2, 0.2), torch.clamp(labels, -0.2, 0.2))
        J = torch.autograd.grad(sdf.sum(), pts, create_graph=True)[0]
        loss = l1 + 1e-2 * (J.norm(p=2) - 1).pow(2) / 1000
        loss = loss + 1e-1 * sum((1 - fc.weight.norm(p=2, dim=1)).pow(2).mean() for fc in net.fc) / len(net.fc)
        loss.backward()
        return float(loss.detach()), [p.grad.detach().clone() for p in net.parameters()]

    def sdf_torch(pts):  # same network with the plain-torch encoding
        h = net.enc.module.forward_autograd(net.preprocess(pts)).float()
        for i, fc in enumerate(net.fc):
            h = fc(h)
            if i != len(net.fc) - 1:
                h = F.relu(h)
        return torch.tanh(h[:, 1:] - h[:, :1])

    la, ga = loss_of(net.sdf)
    lb, gb = loss_of(sdf_torch)
    assert abs(la - lb) <= 1e-6 * max(1.0, abs(lb))
    for (name, _), a, b in zip(net.named_parameters(), ga, gb):
        _close(a, b, name, tol=1e-4)


def test_adam_steps_follow_the_plain_torch_trajectory():
    """Thirty optimizer steps on the reference's loss from the same initial weights and batches:
    the kernels' route and the plain-torch route end at the same loss and the same parameters."""
    import torch.nn.functional as F
    from tropical.stanford.dataset import analytic_sdf
    from tropical.stanford.model import Net

    def sdf_torch(net, pts):
        h = net.enc.module.forward_autograd(net.preprocess(pts)).float()
        for i, fc in enumerate(net.fc):
            h = fc(h)
            if i != len(net.fc) - 1:
                h = F.relu(h)
        return torch.tanh(h[:, 1:] - h[:, :1])

    ends = []
    for route in ("kernels", "torch"):
        torch.manual_seed(0)
        net = Net().cuda()
        with torch.no_grad():
            net.enc.module.params.uniform_(-0.1, 0.1)
        opt = torch.optim.Adam(net.parameters(), lr=1e-3)
        gen = torch.Generator(device="cuda").manual_seed(7)
        for it in range(30):
            pts = (torch.rand(1000, 3, device="cuda", generator=gen) * 2 - 1).requires_grad_(True)
            target = analytic_sdf("sphere", pts.detach())
            opt.zero_grad()
            sdf = net.sdf(pts) if route == "kernels" else sdf_torch(net, pts)
            J = torch.autograd.grad(sdf.sum(), pts, create_graph=True)[0]
            loss = F.l1_loss(torch.clamp(sdf[:, 0], -0.2, 0.2), torch.clamp(target, -0.2, 0.2)) + 1e-2 * (J.norm() - 1) ** 2 / 1000
            loss.backward()
            opt.step()
        ends.append((float(loss.detach()), [p.detach().clone() for p in net.parameters()]))
    (la, pa), (lb, pb) = ends
    assert abs(la - lb) <= 1e-3 * max(abs(lb), 1e-6)
    for a, b in zip(pa, pb):
        _close(a, b, "parameters after 30 Adam steps", tol=1e-3)
