"""Import the UNMODIFIED reference (`/root/reference/tropical`) in this container.

Only usable where `/root/reference` exists (the build container, never the GPU box):
used by `make_golden.py` to generate the fixtures in this directory and by the
`-m "not gpu"` oracle-vs-reference tests (skipped when the reference is absent).

The reference needs CUDA-only / absent packages at import time; `_stubs/` provides
import-only stand-ins plus a torch restatement of the tiny-cuda-nn hash encoding
(see `_stubs/tinycudann.py`).  Nothing here is product code.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("TROPICAL_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tropical"))


def import_reference():
    """Returns (tropical, subpoly_module, Net) from the reference tree."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    # the product mirror is also called `tropical`; make sure the reference wins here
    for name in [k for k in sys.modules if k == "tropical" or k.startswith("tropical.")]:
        mod = sys.modules[name]
        if not getattr(mod, "__file__", "") or REFERENCE_ROOT not in (mod.__file__ or ""):
            del sys.modules[name]
    for p in (_STUBS, REFERENCE_ROOT):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import tropical
    import tropical.subpoly as sp
    from tropical.stanford.model import Net
    assert REFERENCE_ROOT in tropical.__file__
    return tropical, sp, Net


def release_reference():
    """Undo import_reference(): the product mirror is also called `tropical`."""
    for name in [k for k in sys.modules if k == "tropical" or k.startswith("tropical.")]:
        del sys.modules[name]
    for p in (_STUBS, REFERENCE_ROOT):
        while p in sys.path:
            sys.path.remove(p)
    for name in ("tinycudann", "matplotlib", "matplotlib.pyplot", "deprecation", "trimesh", "cubvh", "mcubes"):
        mod = sys.modules.get(name)
        if mod is not None and _STUBS in (getattr(mod, "__file__", "") or ""):
            del sys.modules[name]


def sphere_sdf(x, r=0.6):
    """Analytic SDF, inside positive (the reference's convention, dataset.py:94)."""
    return r - x.norm(dim=-1)


def torus_sdf(x, R=0.55, r=0.22):
    q = ((x[:, 0] ** 2 + x[:, 1] ** 2).sqrt() - R)
    return r - (q ** 2 + x[:, 2] ** 2).sqrt()


def fit(net, sdf_fn, steps=300, batch=4096, lr=1e-2, seed=0, log=None):
    """Brief fit with the reference's loss terms (train.py:181-201): clamped L1,
    eikonal and weight-norm."""
    import torch
    g = torch.Generator().manual_seed(seed)
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, steps)
    for it in range(steps):
        x = torch.rand(batch, 3, generator=g) * 2 - 1
        y = sdf_fn(x)
        opt.zero_grad()
        pred = net.sdf(x)[:, 0]
        loss = (pred.clamp(-0.2, 0.2) - y.clamp(-0.2, 0.2)).abs().mean()
        pts = x.clone().requires_grad_(True)
        J = torch.autograd.grad(net.sdf(pts).sum(), pts, create_graph=True)[0]
        loss = loss + 1e-2 * (J.norm(p=2) - 1).pow(2) / batch
        loss = loss + 1e-1 * sum((1 - fc.weight.norm(p=2, dim=1)).pow(2).mean()
                                 for fc in net.fc) / len(net.fc)
        loss.backward()
        opt.step()
        sched.step()
        if log and (it % 50 == 0 or it == steps - 1):
            log(f"fit step {it}: loss {loss.item():.5f}")
    return net
