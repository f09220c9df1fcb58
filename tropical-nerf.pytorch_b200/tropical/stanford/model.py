"""Net (reference: tropical/stanford/model.py), CUDA-backed for the extraction path."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from tropical import TropicalHashGrid
from tropical import _native


class Net(nn.Module):
    """HashGrid + ReLU MLP SDF network (model.py:18-50).  Same constructor, parameter names
    and methods as the reference.  Under torch.no_grad() (the extraction path, which is
    @torch.no_grad in the reference too) forward/sdf/region/normal run in the fused sm_100a
    kernels and input gradients are computed analytically on the device; with autograd enabled
    (the training loop) the encoding runs in the sm_100a training kernels (twice differentiable) and
    the three nn.Linear layers in torch."""

    def __init__(self, num_layers: int = 3, num_hidden: int = 16, levels: int = 4,
                 r_min: int = 2, r_max: int = 32, T: int = 19, eps: float = 1e-4):
        super().__init__()
        DIM = 3
        self.num_layers, self.num_hidden, self.eps = num_layers, num_hidden, eps
        L, Fe = levels, 2
        self.scale = 1
        self.enc = TropicalHashGrid(1.0, DIM, L, Fe, T, r_min, r_max, eps)
        num_nodes = [L * Fe] + [num_hidden] * (num_layers - 1) + [2]
        self.num_nodes = num_nodes
        self.fc = nn.ModuleList([nn.Linear(num_nodes[i], num_nodes[i + 1])
                                 for i in range(len(num_nodes) - 1)])
        self._native_cache = None

    # ---- device copy of the network ------------------------------------------------------
    def native(self) -> "_native.NativeNet":
        """The tnb_net for the current parameters (rebuilt when any parameter changed)."""
        key = tuple((p._version, p.data_ptr()) for p in self.parameters()) + (float(self.eps), float(self.scale),
                                                                                  self.enc.marks.data_ptr(), self.enc.marks._version)
        if self._native_cache is None or self._native_cache[0] != key:
            mlp = np.concatenate([np.concatenate([fc.weight.detach().cpu().numpy().reshape(-1),
                                                  fc.bias.detach().cpu().numpy().reshape(-1)])
                                  for fc in self.fc]).astype(np.float32)
            nat = _native.NativeNet(self.enc.L, self.enc.F, self.enc.T, self.enc.N_min, self.enc.b,
                                    self.num_layers, self.num_hidden,
                                    self.enc.module.params.detach().cpu().numpy(), mlp,
                                    self.enc.marks.cpu().numpy(), self.eps, self.scale)
            self._native_cache = (key, nat)
        return self._native_cache[1]

    def forward(self, x, gather: bool = False, group: int = 1):
        """model.py:52-76.  gather=True returns (output, [hidden pre-activations..., o1-o0])."""
        if torch.is_grad_enabled() and (x.requires_grad or self.fc[0].weight.requires_grad):
            return self._forward_autograd(x, gather, group)
        nat = self.native()
        H = self.num_hidden
        if group == 1:
            rows, out = nat.forward(x, rows=gather)   # both return values from one fused pass (no torch layer)
            if not gather:
                return out
        elif group == 8:  # "infer within a common linear space" (model.py:65-70): the corners of an edge's box
            rows, out = nat.outputs_group8(x)
        else:
            raise _native.NativeError("Net.forward: group must be 1 or 8 on the device (the path uses 8: subpoly.py:125)")
        inputs = [rows[:, i * H:(i + 1) * H] for i in range(self.num_layers - 1)] + [rows[:, -1:]]
        if gather:
            return out, inputs
        return out

    def _forward_autograd(self, x, gather=False, group=1):
        """model.py:52-76 under autograd (training): device encoding kernels + nn.Linear."""
        inputs = []
        h = self.enc.module.forward_train(self.preprocess(x))
        last = len(self.fc) - 1
        for i, fc in enumerate(self.fc):
            h = fc(h)
            if i != last:
                inputs.append(h)
                if group == 1:
                    h = F.relu(h)
                else:  # infer within a common linear space (model.py:65-70)
                    m = (h[::group] > self.eps) | (h[group - 1::group] > self.eps)
                    h = h * m.repeat(1, group).view(*h.shape)
            else:
                inputs.append(h[:, 1:] - h[:, :1])
        return (h, inputs) if gather else h

    def preprocess(self, x):
        return (x + self.scale) / (self.scale * 2)

    def preprocess_inverse(self, x):
        return x * (self.scale * 2) - self.scale

    def sdf(self, x):
        """tanh(o1 - o0) (model.py:84-88), [n,1]."""
        if torch.is_grad_enabled() and (x.requires_grad or self.fc[0].weight.requires_grad):
            out = self._forward_autograd(x)
            return torch.tanh(out[:, 1:] - out[:, :1])
        sdf, _ = self.native().sdf_grad(x, want_grad=False)
        return sdf.unsqueeze(-1)

    def region(self, vertices: Tensor, output: Tensor = None, eps=None):
        """Sign vectors + grid offsets (model.py:90-103): (m [n,3+R] int64, offset [n,3]
        int64, output [n,R])."""
        signs, offset, output = self.native().region(vertices, output, eps)
        return signs.long(), offset.long(), output

    def normal(self, vertices: Tensor, l: int = None, h: int = None, create_graph=False,
               return_y=False) -> Tensor:
        """d sdf / d x (model.py:105-123), analytic on the device."""
        if not (l is None or h is None or h == self.num_hidden) or create_graph:
            raise _native.NativeError("only the SDF normal (l=h=None) is available on the device")
        y, J = self.native().sdf_grad(vertices.detach())
        if return_y:
            return J, y.unsqueeze(-1)
        return J

    def device(self):
        return next(self.parameters()).device
