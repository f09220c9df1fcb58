"""CPU stand-in for `tinycudann` used ONLY to import the reference in this container.

The reference (`/root/reference/tropical/tropical.py:17,32-40`) builds its HashGrid
with `tcnn.Encoding(D, {"otype": "Grid", "type": "Hash", ...}, dtype=torch.float)`.
tiny-cuda-nn is not vendored in the reference tree, is not listed in its
requirements.txt (so it is unpinned) and cannot be built here (no GPU, no network).
This module restates the published multiresolution hash encoding (Mueller et al.
2022, Sec. 3 and Appendix A; tiny-cuda-nn `GridEncoding`) in differentiable torch
so that the reference's own Python (`tropical.subpoly`, `tropical.tropical`,
`tropical.stanford.model`) runs unmodified on CPU and produces golden vectors.

Algorithm restated (per level l):
    scale_l  = exp2f(l * log2f(per_level_scale)) * base_resolution - 1
    res_l    = ceil(scale_l) + 1
    size_l   = min(next_multiple(res_l^3, 8), 2^log2_hashmap_size)
    pos      = x * scale_l + 0.5 ; cell = floor(pos) ; frac = pos - cell
    index(c) = (cx + cy*res + cz*res^2)            if res^3 <= size_l  (dense)
             = cx ^ cy*2654435761 ^ cz*805459861   otherwise           (hashed)
               both taken mod size_l in uint32 arithmetic
    out_f    = sum_{corner} prod_d (frac_d | 1-frac_d) * table[offset_l + index][f]

It is test infrastructure (golden-vector generation), never imported by the product.
"""
import os

import numpy as np
import torch
import torch.nn as nn

_PRIMES = (1, 2654435761, 805459861)
_U32 = 0xFFFFFFFF



def _libm():
    """glibc's log2f/exp2f: what tiny-cuda-nn's host code calls for the level scales
    (numpy's float32 log2/exp2 differ from libm in the last ulp for some levels)."""
    import ctypes
    global _LIBM
    try:
        return _LIBM
    except NameError:
        _LIBM = ctypes.CDLL("libm.so.6")
        for fn in (_LIBM.log2f, _LIBM.exp2f):
            fn.restype = ctypes.c_float
            fn.argtypes = [ctypes.c_float]
        return _LIBM


def grid_layout(n_levels, log2_hashmap_size, base_resolution, per_level_scale, n_dims=3):
    """Level scales / resolutions / table sizes exactly as tiny-cuda-nn derives them
    (float32 arithmetic for the scale, uint32 for sizes)."""
    log2_pls = np.float32(_libm().log2f(np.float32(per_level_scale)))
    scales, ress, sizes, offsets = [], [], [], []
    offset = 0
    for l in range(n_levels):
        s = np.float32(np.float32(_libm().exp2f(np.float32(l) * log2_pls))
                       * np.float32(base_resolution) - np.float32(1.0))
        res = int(np.ceil(s)) + 1
        max_params = (2 ** 32 - 1) // 2
        dense = res ** n_dims
        n = max_params if float(res) ** n_dims > float(max_params) else dense
        n = (n + 7) // 8 * 8
        n = min(n, 1 << log2_hashmap_size)
        scales.append(float(s)); ress.append(res); sizes.append(n); offsets.append(offset)
        offset += n
    return scales, ress, sizes, offsets, offset


class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config, dtype=torch.float, seed=1337):
        super().__init__()
        cfg = encoding_config
        assert cfg["otype"] == "Grid" and cfg["type"] == "Hash" and n_input_dims == 3
        self.n_input_dims = n_input_dims
        self.L = int(cfg["n_levels"])
        self.F = int(cfg["n_features_per_level"])
        self.T = int(cfg["log2_hashmap_size"])
        self.N_min = int(cfg["base_resolution"])
        self.b = float(cfg["per_level_scale"])
        (self.scales, self.ress, self.sizes, self.offsets, total) = grid_layout(
            self.L, self.T, self.N_min, self.b)
        self.n_output_dims = self.L * self.F
        g = torch.Generator().manual_seed(seed)
        # tiny-cuda-nn initialises grid parameters U(-1e-4, 1e-4)
        init = (torch.rand(total * self.F, generator=g) * 2 - 1) * 1e-4
        self.params = nn.Parameter(init.to(dtype))

    def _index(self, c, res, size):
        # c: (N,3) int64 holding uint32 values
        stride, index, dim = 1, torch.zeros_like(c[:, 0]), 0
        while dim < 3 and stride <= size:
            index = (index + c[:, dim] * stride) & _U32
            stride = (stride * res) & _U32
            dim += 1
        if size < stride:
            index = torch.zeros_like(c[:, 0])
            for d in range(3):
                index = index ^ ((c[:, d] * _PRIMES[d]) & _U32)
        return index % size

    def forward(self, x):
        if os.environ.get("TNB_STUB_FMA") == "1" and not torch.is_grad_enabled():
            return self._forward_fma(x)
        table = self.params.view(-1, self.F)
        outs = []
        for l in range(self.L):
            scale = torch.tensor(self.scales[l], dtype=x.dtype)
            pos = x * scale + 0.5
            cell_f = torch.floor(pos)
            frac = pos - cell_f
            cell = cell_f.detach().to(torch.int64) & _U32  # (uint32)(int) wrap
            acc = 0
            for corner in range(8):
                w = 1
                cc = []
                for d in range(3):
                    if corner & (1 << d):
                        w = w * frac[:, d]
                        cc.append((cell[:, d] + 1) & _U32)
                    else:
                        w = w * (1 - frac[:, d])
                        cc.append(cell[:, d])
                idx = self._index(torch.stack(cc, -1), self.ress[l], self.sizes[l])
                acc = acc + w.unsqueeze(-1) * table[self.offsets[l] + idx]
            outs.append(acc)
        return torch.cat(outs, dim=-1)

    def _forward_fma(self, x):
        """The same interpolation with the two fused multiply-adds tiny-cuda-nn's CUDA code compiles to
        (`pos = fma(scale, x, 0.5)`, `acc = fma(w, value, acc)`), emulated in float64 (exact product, one
        rounding of the sum, then the rounding to float32): the operation order oracle/trinet_ref.c and the
        device define.  Opt-in (TNB_STUB_FMA=1, inference only) for tests/golden/make_golden_medium.py --detlin,
        whose point is a reference run that is reproducible bit for bit."""
        table = self.params.detach().view(-1, self.F)
        outs = []
        for l in range(self.L):
            scale = torch.tensor(self.scales[l], dtype=torch.float64)
            pos = (x.double() * scale + 0.5).float()
            cell_f = torch.floor(pos)
            frac = pos - cell_f
            cell = cell_f.to(torch.int64) & _U32
            acc = torch.zeros(x.shape[0], self.F, dtype=torch.float32)
            for corner in range(8):
                w = torch.ones(x.shape[0], dtype=torch.float32)
                cc = []
                for d in range(3):
                    if corner & (1 << d):
                        w = w * frac[:, d]
                        cc.append((cell[:, d] + 1) & _U32)
                    else:
                        w = w * (1 - frac[:, d])
                        cc.append(cell[:, d])
                idx = self._index(torch.stack(cc, -1), self.ress[l], self.sizes[l])
                acc = (acc.double() + w.double().unsqueeze(-1) * table[self.offsets[l] + idx].double()).float()
            outs.append(acc)
        return torch.cat(outs, dim=-1)
