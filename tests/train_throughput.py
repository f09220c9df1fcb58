"""Manual: milliseconds per training iteration (stanford/train.py:180-201, batch 1000) with the
encoding in the sm_100a training kernels against the plain-torch restatement of the same loss."""
import os, sys, time
import torch
import torch.nn.functional as F
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tropical-nerf.pytorch_b200"), HERE):
    sys.path.insert(0, p)
from tropical.stanford.model import Net
from tropical.stanford.dataset import analytic_sdf

size = sys.argv[1] if len(sys.argv) > 1 else "small"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
r_min, r_max = {"small": (2, 32), "medium": (4, 64), "large": (8, 128)}[size]


def sdf_torch(net, pts):
    h = net.enc.module.forward_autograd(net.preprocess(pts)).float()
    for i, fc in enumerate(net.fc):
        h = fc(h)
        if i != len(net.fc) - 1:
            h = F.relu(h)
    return torch.tanh(h[:, 1:] - h[:, :1])


for route in ("kernels", "torch"):
    torch.manual_seed(0)
    net = Net(num_layers=3, num_hidden=16, levels=4, r_min=r_min, r_max=r_max, T=19).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    fn = net.sdf if route == "kernels" else (lambda p: sdf_torch(net, p))
    def it():
        pts = (torch.rand(batch, 3, device="cuda") * 2 - 1).requires_grad_(True)
        target = analytic_sdf("sphere", pts.detach())
        opt.zero_grad()
        sdf = fn(pts)
        l1 = F.l1_loss(torch.clamp(sdf[:, 0], -0.2, 0.2), torch.clamp(target, -0.2, 0.2))
        J = torch.autograd.grad(sdf.sum(), pts, create_graph=True)[0]
        loss = l1 + 1e-2 * (J.norm(p=2) - 1).pow(2) / batch
        loss = loss + 1e-1 * sum((1 - fc.weight.norm(p=2, dim=1)).pow(2).mean() for fc in net.fc) / len(net.fc)
        loss.backward()
        opt.step()
        return loss
    for _ in range(10):
        it()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 100
    for _ in range(n):
        loss = it()
    torch.cuda.synchronize()
    print(f"{size} batch {batch} {route}: {(time.perf_counter() - t0) / n * 1e3:.3f} ms / iteration, loss {float(loss):.5f}")
