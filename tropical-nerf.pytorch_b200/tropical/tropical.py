"""TropicalHashGrid / Tropical (reference: tropical/tropical.py), CUDA-backed.

`TropicalHashGrid.module` is a `HashEncoding` -- the tiny-cuda-nn `Encoding` stand-in whose
forward runs the fused sm_100a gather kernel (`tnb_grid_encode`).  Parameter names match
the reference (`enc.module.params`), so its released state_dicts load unchanged.
"""
import ctypes
from typing import Any, Tuple

import numpy as np
import torch
from torch import LongTensor, Tensor
from torch.nn import Module

from . import _native

__all__ = ["HashEncoding", "TropicalHashGrid", "Tropical", "low_precision"]


class _GridEncode(torch.autograd.Function):
    """enc = encoding(x; table) on the device (`tnb_grid_train_forward`), differentiable twice:
    its backward is itself an autograd Function, because the reference's training loss
    differentiates d sdf / d x once more (eikonal term, stanford/train.py:194-197)."""

    @staticmethod
    def forward(ctx, x, params, desc):
        xc = x.detach().contiguous().float()
        n = xc.shape[0]
        enc = torch.empty(n, 2 * desc.n_levels, device=xc.device, dtype=torch.float32)
        with torch.cuda.device(xc.device):
            _native.check(_native.lib().tnb_grid_train_forward(
                ctypes.byref(desc), _native._ptr(params.detach()), _native._ptr(xc), n, _native._ptr(enc), _native._stream()))
        ctx.save_for_backward(x, params)  # the inputs themselves: the backward is differentiable w.r.t. both
        ctx.desc = desc
        return enc

    @staticmethod
    def backward(ctx, denc):
        x, params = ctx.saved_tensors
        dx, dparams = _GridEncodeBackward.apply(denc, x, params, ctx.desc, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dx, dparams, None


class _GridEncodeBackward(torch.autograd.Function):
    """(denc, x, table) -> (dLoss/dx, dLoss/dtable) (`tnb_grid_train_backward`); its own backward
    (`tnb_grid_train_backward_backward`) differentiates the dx output w.r.t. denc, x and the table.
    The table-gradient output is first order only."""

    @staticmethod
    def forward(ctx, denc, x, params, desc, want_dx, want_dparams):
        dc = denc.detach().contiguous().float()
        xc = x.detach().contiguous().float()
        n = xc.shape[0]
        dx = torch.empty(n, 3, device=xc.device, dtype=torch.float32) if want_dx else None
        dparams = torch.zeros_like(params) if want_dparams else None
        with torch.cuda.device(xc.device):
            _native.check(_native.lib().tnb_grid_train_backward(
                ctypes.byref(desc), _native._ptr(params.detach()), _native._ptr(xc), n, _native._ptr(dc),
                _native._ptr(dparams), _native._ptr(dx), _native._stream()))
        ctx.save_for_backward(denc, x, params)
        ctx.desc = desc
        ctx.set_materialize_grads(False)
        return dx, dparams

    @staticmethod
    def backward(ctx, ddx, ddparams):
        if ddparams is not None:
            raise _native.NativeError("the table gradient of the hash-grid encoding is first order only")
        if ddx is None:
            return None, None, None, None, None, None
        denc, x, params = ctx.saved_tensors
        denc, x = denc.detach().contiguous().float(), x.detach().contiguous().float()
        n = x.shape[0]
        need = ctx.needs_input_grad
        ddx = ddx.detach().contiguous().float()
        g_denc = torch.empty_like(denc) if need[0] else None
        g_x = torch.empty(n, 3, device=x.device, dtype=torch.float32) if need[1] else None
        g_params = torch.zeros_like(params) if need[2] else None
        with torch.cuda.device(x.device):
            _native.check(_native.lib().tnb_grid_train_backward_backward(
                ctypes.byref(ctx.desc), _native._ptr(params.detach()), _native._ptr(x), n, _native._ptr(denc), _native._ptr(ddx),
                _native._ptr(g_params), _native._ptr(g_denc), _native._ptr(g_x), _native._stream()))
        return g_denc, g_x, g_params, None, None, None


class HashEncoding(Module):
    """Multiresolution hash encoding with tiny-cuda-nn's parameter layout
    (`tcnn.Encoding(D, {"otype": "Grid", "type": "Hash", ...}, dtype=torch.float)`,
    tropical.py:32-40).

    Under `torch.no_grad()` (the whole extraction path) the fused sm_100a kernel runs
    (`tnb_grid_encode`).  When autograd needs the result (the training loop of stanford/train.py,
    including its double-backward eikonal term) the sm_100a training kernels run
    (`tnb_grid_train_forward/_backward/_backward_backward`, csrc/grid_train.cu) straight on the
    parameter storage; there is no CPU route: parameters that are not on a CUDA device raise.
    `forward_autograd` is the same interpolation in plain torch ops: the restatement the kernels
    are tested against (and what bench.py's workload generator mirrors), not a product path."""

    def __init__(self, n_input_dims, n_levels, n_features_per_level, log2_hashmap_size,
                 base_resolution, per_level_scale, seed=1337):
        super().__init__()
        assert n_input_dims == 3 and n_features_per_level == 2
        self.n_levels, self.n_features = n_levels, n_features_per_level
        self.log2_hashmap_size, self.base_resolution = log2_hashmap_size, base_resolution
        self.per_level_scale = float(per_level_scale)
        self.n_output_dims = n_levels * n_features_per_level
        self.level_sizes, self.level_scales, self.level_res = self._layout()
        g = torch.Generator().manual_seed(seed)
        init = (torch.rand(int(sum(self.level_sizes)) * self.n_features, generator=g) * 2 - 1) * 1e-4
        self.params = torch.nn.Parameter(init)  # U(-1e-4, 1e-4) like tiny-cuda-nn
        # set by TropicalHashGrid (kept out of nn.Module's registry: no module cycle)
        self.__dict__["_owner"] = None

    def _layout(self):
        import ctypes
        libm = ctypes.CDLL("libm.so.6")
        for fn in (libm.log2f, libm.exp2f):
            fn.restype, fn.argtypes = ctypes.c_float, [ctypes.c_float]
        log2_pls = np.float32(libm.log2f(np.float32(self.per_level_scale)))
        sizes, scales, ress = [], [], []
        for l in range(self.n_levels):
            s = np.float32(np.float32(libm.exp2f(np.float32(l) * log2_pls))
                           * np.float32(self.base_resolution) - np.float32(1.0))
            res = int(np.ceil(s)) + 1
            n = min((res ** 3 + 7) // 8 * 8, 1 << self.log2_hashmap_size)
            sizes.append(n)
            scales.append(float(s))
            ress.append(res)
        return sizes, scales, ress

    def forward_autograd(self, x: Tensor) -> Tensor:
        """Differentiable (any order) torch expression of the encoding, for training."""
        U32 = 0xFFFFFFFF
        table = self.params.view(-1, self.n_features)
        outs, off = [], 0
        for l in range(self.n_levels):
            res, size = self.level_res[l], self.level_sizes[l]
            pos = x * self.level_scales[l] + 0.5
            cell_f = torch.floor(pos)
            frac = pos - cell_f
            cell = cell_f.detach().long() & U32
            acc = 0
            for corner in range(8):
                w, cc = 1, []
                for d in range(3):
                    bit = (corner >> d) & 1
                    w = w * (frac[:, d] if bit else 1 - frac[:, d])
                    cc.append((cell[:, d] + bit) & U32)
                if res ** 3 <= size:
                    idx = (cc[0] + cc[1] * res + cc[2] * res * res) & U32
                else:
                    idx = cc[0] ^ ((cc[1] * 2654435761) & U32) ^ ((cc[2] * 805459861) & U32)
                acc = acc + w.unsqueeze(-1) * table[off + idx % size]
            outs.append(acc)
            off += size
        return torch.cat(outs, dim=-1)

    def grid_desc(self):
        return _native.GridDesc(self.n_levels, self.log2_hashmap_size, self.base_resolution, self.per_level_scale)

    def forward_train(self, x: Tensor) -> Tensor:
        """The encoding under autograd, on the device (twice differentiable in x and the table)."""
        if not (self.params.is_cuda and x.is_cuda):
            raise _native.NativeError("the hash-grid encoding trains on a CUDA device only (no CPU route); "
                                      "move the network and the batch to cuda")
        return _GridEncode.apply(x, self.params, self.grid_desc())

    def forward(self, x: Tensor) -> Tensor:
        if torch.is_grad_enabled() and (x.requires_grad or self.params.requires_grad):
            return self.forward_train(x)
        return self._owner._native_for_encoding().encode(x)


class TropicalHashGrid(Module):
    """tropical.py:20-239."""

    def __init__(self, scale: float = 1.0, D: int = 3, L: int = 16, F: int = 2,
                 T: int = 19, N_min: int = 16, N_max: int = 2048, eps: float = 1e-4):
        super().__init__()
        self.scale, self.D, self.L, self.F, self.T = scale, D, L, F, T
        self.N_min, self.N_max = N_min, N_max
        self.b = np.exp2(np.log2(N_max * scale / N_min) / (L - 1))
        self.module = HashEncoding(D, L, F, T, N_min, self.b)
        self.module.__dict__["_owner"] = self
        self.eps = eps
        self.marks = self._marks()
        self._enc_native = None

    def forward(self, x):
        return self.module(x)

    def _marks(self):
        """Aggregated, sorted, eps-merged grid marks of all levels (tropical.py:49-79).
        Same torch operations as the reference, evaluated on the host."""
        vertices = []
        for l in range(self.L):
            grid_scale = np.exp2(l * np.log2(self.b)) * self.N_min - 1.0
            unit = 1 / grid_scale
            vertices += [torch.arange(0, 1.5, unit) - 0.5 * unit]
        vertices += [torch.Tensor([0, self.scale])]
        marks, _ = torch.cat(vertices).unique().sort()
        m = marks.new_zeros(len(marks)).bool().fill_(True)
        for i in range(len(marks) - 1):
            if self.eps > (marks[i] - marks[i + 1]).abs():
                marks[i + 1] = (marks[i] + marks[i + 1]) / 2
                m[i] = False
        marks = marks[m]
        marks = marks[marks >= 0]
        marks = marks[marks <= self.scale]
        return marks

    # encoding-only native net (no MLP owner): a 2-layer dummy MLP keeps the ABI uniform
    def _native_for_encoding(self):
        key = (self.module.params._version, self.module.params.data_ptr())
        if self._enc_native is None or self._enc_native[0] != key:
            mlp = np.zeros(2 * self.L * self.F + 2 + 2 * 2 + 2, np.float32)  # [L*F -> 2 -> 2]
            nn_ = _native.NativeNet(self.L, self.F, self.T, self.N_min, self.b, 2, 2,
                                    self.module.params.detach().cpu().numpy(), mlp,
                                    self.marks.cpu().numpy(), self.eps, self.scale)
            self._enc_native = (key, nn_)
        return self._enc_native[1]

    def p2v(self, indices: LongTensor) -> LongTensor:
        """Serialized vertex index from marks-grid indices (tropical.py:141-146)."""
        L = len(self.marks)
        weights = indices.new_tensor([L ** (self.D - 1 - i) for i in range(self.D)])
        return (indices * weights).sum(dim=-1).long()

    def v2p(self, v_idx: LongTensor) -> LongTensor:
        """tropical.py:149-156."""
        L = len(self.marks)
        p, rest = [], v_idx.clone()
        for i in range(self.D - 1, -1, -1):
            q = torch.div(rest, L ** i, rounding_mode="floor")
            p.append(q.long())
            rest = rest - q * L ** i
        return torch.stack(p, dim=-1)

    def skeleton(self, net: Module, unit: int = 128) -> Tuple[Tensor, Tensor]:
        """Starting skeleton: the marks-grid edges near the surface (tropical.py:158-225,
        distance pruning).  Runs `tnb_skeleton`; returns (vertices [V,3], edges [V,2]) on the
        device, or two empty tensors when nothing survives (tropical.py:208-209)."""
        cx = net.native().skeleton(unit, 0.0)  # size 0: no hypercube fallback here
        if cx.num_edges == 0:
            e = torch.empty(0, dtype=torch.int64, device="cuda")
            return e, e
        v, e, _ = cx.read(outputs=False)
        return v, e

    def region(self, x: Tensor, eps: float = None) -> Tuple[Tensor, Tensor]:
        """Epsilon-tolerant grid offsets and on-plane masks (tropical.py:227-236)."""
        eps = eps if eps is not None else self.eps
        marks = self.marks.to(x.device)
        offset = torch.searchsorted(marks, x + eps) - 1
        mask = ((marks[offset] - x).abs() > eps).long()
        return mask, offset

    def device(self):
        return next(self.parameters()).device


class Tropical(Module):
    """tropical.py:242-281."""

    def __init__(self, module: Module, dim: int = 3, scale: float = 1.0):
        super().__init__()
        self.module, self.dim, self.scale = module, dim, scale

    def region(self, x: Tensor) -> Any:
        return NotImplementedError

    def grid(self) -> Tuple[Tensor, Tensor]:
        for m in self.module.modules():
            if isinstance(m, TropicalHashGrid):
                return m.skeleton(self.module)
        vertices, edges, _ = self.get_hypercube(self.dim, self.scale / 2)
        return vertices, edges

    def get_hypercube(self, d, size):
        from .subpoly import get_hypercube
        return get_hypercube(d, size)


def low_precision(x):
    """tropical.py:284-288."""
    x *= 100000
    x = x.floor()
    x /= 100000
    return x
