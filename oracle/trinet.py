"""oracle/trinet.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes binding of `oracle/trinet_ref.c` plus the plain-data description of a
trilinear network (`NetParams`).  Only tests/, `__graft_entry__.smoke()` and
bench.py's cpu_baseline / `--impl reference` legs may import this module.

`NetParams` carries exactly what the reference's `Net` (tropical/stanford/model.py:18-50)
and `TropicalHashGrid` (tropical/tropical.py:20-44) hold: the hash-grid table, the
per-level layout tiny-cuda-nn derives from (L, F, T, N_min, per_level_scale), the
nn.Linear weights, the marks and eps.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libtrinet_ref.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc via oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "trinet_ref.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


class _CNet(ctypes.Structure):
    _fields_ = [("n_levels", ctypes.c_int32), ("n_feat", ctypes.c_int32),
                ("n_linear", ctypes.c_int32), ("n_hidden", ctypes.c_int32),
                ("pre_scale", ctypes.c_float),
                ("lvl_scale", ctypes.c_void_p), ("lvl_res", ctypes.c_void_p),
                ("lvl_size", ctypes.c_void_p), ("lvl_off", ctypes.c_void_p),
                ("table", ctypes.c_void_p), ("mlp", ctypes.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.det_tanhf.restype = ctypes.c_float
        _lib.det_tanhf.argtypes = [ctypes.c_float]
    return _lib



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
    """Per-level (scale, resolution, table size, table offset) as tiny-cuda-nn's
    GridEncoding constructor derives them: float32 scale, uint32 sizes rounded up to a
    multiple of 8 and capped at 2^T."""
    log2_pls = np.float32(_libm().log2f(np.float32(per_level_scale)))
    scale = np.zeros(n_levels, np.float32)
    res = np.zeros(n_levels, np.uint32)
    size = np.zeros(n_levels, np.uint32)
    off = np.zeros(n_levels, np.uint32)
    total = 0
    for l in range(n_levels):
        s = np.float32(np.float32(_libm().exp2f(np.float32(l) * log2_pls))
                       * np.float32(base_resolution) - np.float32(1.0))
        r = int(np.ceil(s)) + 1
        max_params = (2 ** 32 - 1) // 2
        n = max_params if float(r) ** n_dims > float(max_params) else r ** n_dims
        n = (n + 7) // 8 * 8
        n = min(n, 1 << log2_hashmap_size)
        scale[l], res[l], size[l], off[l] = s, r, n, total
        total += n
    return scale, res, size, off, total


class NetParams:
    """Plain-data trilinear network: hash grid + ReLU MLP + marks."""

    def __init__(self, levels, n_feat, log2_T, n_min, per_level_scale, num_layers, num_hidden,
                 table, weights, biases, marks, eps=1e-4, scale=1.0):
        self.levels, self.n_feat, self.log2_T = int(levels), int(n_feat), int(log2_T)
        self.n_min, self.per_level_scale = int(n_min), float(per_level_scale)
        self.num_layers, self.num_hidden = int(num_layers), int(num_hidden)
        self.eps, self.scale = float(eps), float(scale)
        (self.lvl_scale, self.lvl_res, self.lvl_size, self.lvl_off, total) = grid_layout(
            self.levels, self.log2_T, self.n_min, self.per_level_scale)
        self.table = np.ascontiguousarray(table, np.float32).reshape(-1)
        assert self.table.size == total * self.n_feat, (self.table.size, total, self.n_feat)
        self.weights = [np.ascontiguousarray(w, np.float32) for w in weights]
        self.biases = [np.ascontiguousarray(b, np.float32) for b in biases]
        assert len(self.weights) == self.num_layers
        self.mlp = np.concatenate([np.concatenate([w.reshape(-1), b.reshape(-1)])
                                   for w, b in zip(self.weights, self.biases)]).astype(np.float32)
        self.marks = np.ascontiguousarray(marks, np.float32)
        self.n_outputs = (self.num_layers - 1) * self.num_hidden + 1
        self._c = _CNet(self.levels, self.n_feat, self.num_layers, self.num_hidden,
                        np.float32(self.scale),
                        self.lvl_scale.ctypes.data, self.lvl_res.ctypes.data,
                        self.lvl_size.ctypes.data, self.lvl_off.ctypes.data,
                        self.table.ctypes.data, self.mlp.ctypes.data)

    # ---- (de)serialisation used by the golden fixtures -------------------------------
    def to_npz_dict(self):
        d = dict(levels=self.levels, n_feat=self.n_feat, log2_T=self.log2_T, n_min=self.n_min,
                 per_level_scale=np.float64(self.per_level_scale), num_layers=self.num_layers,
                 num_hidden=self.num_hidden, eps=np.float64(self.eps),
                 scale=np.float64(self.scale), table=self.table, marks=self.marks)
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            d[f"w{i}"], d[f"b{i}"] = w, b
        return d

    @classmethod
    def from_npz_dict(cls, d):
        nl = int(d["num_layers"])
        return cls(int(d["levels"]), int(d["n_feat"]), int(d["log2_T"]), int(d["n_min"]),
                   float(d["per_level_scale"]), nl, int(d["num_hidden"]), d["table"],
                   [d[f"w{i}"] for i in range(nl)], [d[f"b{i}"] for i in range(nl)],
                   d["marks"], float(d["eps"]), float(d["scale"]))

    @classmethod
    def from_reference_net(cls, net):
        """From an instance of the reference's `Net` (golden generation only)."""
        enc = net.enc
        return cls(enc.L, enc.F, enc.T, enc.N_min, enc.b, net.num_layers, net.num_hidden,
                   enc.module.params.detach().cpu().numpy(),
                   [fc.weight.detach().cpu().numpy() for fc in net.fc],
                   [fc.bias.detach().cpu().numpy() for fc in net.fc],
                   enc.marks.detach().cpu().numpy(), net.eps, net.scale)

    # ---- evaluation through the C restatement ---------------------------------------
    def preprocess(self, x):
        x = np.asarray(x, np.float32)
        return (x + np.float32(self.scale)) / np.float32(self.scale * 2)

    def preprocess_inverse(self, x):
        x = np.asarray(x, np.float32)
        return x * np.float32(self.scale * 2) - np.float32(self.scale)

    def encode(self, xp):
        xp = np.ascontiguousarray(xp, np.float32).reshape(-1, 3)
        out = np.empty((xp.shape[0], self.levels * self.n_feat), np.float32)
        lib().trinet_encode(ctypes.byref(self._c), ctypes.c_void_p(xp.ctypes.data),
                            ctypes.c_int64(xp.shape[0]), ctypes.c_void_p(out.ctypes.data))
        return out

    def outputs(self, x):
        x = np.ascontiguousarray(x, np.float32).reshape(-1, 3)
        out = np.empty((x.shape[0], self.n_outputs), np.float32)
        lib().trinet_outputs(ctypes.byref(self._c), ctypes.c_void_p(x.ctypes.data),
                             ctypes.c_int64(x.shape[0]), ctypes.c_void_p(out.ctypes.data))
        return out

    def sdf_grad(self, x, want_grad=True):
        x = np.ascontiguousarray(x, np.float32).reshape(-1, 3)
        sdf = np.empty(x.shape[0], np.float32)
        grad = np.empty((x.shape[0], 3), np.float32) if want_grad else None
        lib().trinet_sdf_grad(ctypes.byref(self._c), ctypes.c_void_p(x.ctypes.data),
                              ctypes.c_int64(x.shape[0]), ctypes.c_void_p(sdf.ctypes.data),
                              ctypes.c_void_p(grad.ctypes.data if want_grad else None))
        return sdf, grad

    @staticmethod
    def grad_norm(grad):
        grad = np.ascontiguousarray(grad, np.float32).reshape(-1, 3)
        out = np.empty(grad.shape[0], np.float32)
        lib().trinet_grad_norm(ctypes.c_void_p(grad.ctypes.data), ctypes.c_int64(grad.shape[0]),
                               ctypes.c_void_p(out.ctypes.data))
        return out

    def region(self, x, outputs=None, eps=None):
        """Net.region (model.py:90-103): returns (m, offset, outputs); m is int8
        [n, 3+R] = grid masks then neuron signs, offset int32 [n, 3]."""
        x = np.ascontiguousarray(x, np.float32).reshape(-1, 3)
        if outputs is None:
            outputs = self.outputs(x)
        outputs = np.ascontiguousarray(outputs, np.float32)
        eps = np.float32(self.eps if eps is None else eps)
        m = np.empty((x.shape[0], 3 + self.n_outputs), np.int8)
        off = np.empty((x.shape[0], 3), np.int32)
        lib().trinet_region(ctypes.byref(self._c), ctypes.c_void_p(self.marks.ctypes.data),
                            ctypes.c_int32(self.marks.size), ctypes.c_float(eps),
                            ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(outputs.ctypes.data),
                            ctypes.c_int64(x.shape[0]), ctypes.c_void_p(m.ctypes.data),
                            ctypes.c_void_p(off.ctypes.data))
        return m, off, outputs

    # ---- curve-approximation path ------------------------------------------------------
    def outputs_group8(self, corners, eps=None):
        """Net.forward(gather=True, group=8) rows for [G,8,3] corner points -> [G,8,R]."""
        x = np.ascontiguousarray(corners, np.float32).reshape(-1, 8, 3)
        out = np.empty((x.shape[0], 8, self.n_outputs), np.float32)
        lib().trinet_outputs_group8(ctypes.byref(self._c), ctypes.c_void_p(x.ctypes.data),
                                    ctypes.c_int64(x.shape[0]),
                                    ctypes.c_float(np.float32(self.eps if eps is None else eps)),
                                    ctypes.c_void_p(out.ctypes.data))
        return out


def curve_intersections(p, q):
    """geometry.intersection_of_two_planes for [E,8] corner values -> [E,3]."""
    p = np.ascontiguousarray(p, np.float32).reshape(-1, 8)
    q = np.ascontiguousarray(q, np.float32).reshape(-1, 8)
    out = np.empty((p.shape[0], 3), np.float32)
    lib().curve_intersections(ctypes.c_void_p(p.ctypes.data), ctypes.c_void_p(q.ctypes.data),
                              ctypes.c_int64(p.shape[0]), ctypes.c_void_p(out.ctypes.data))
    return out


def gradient_descent(P, e0, e1, ints, plane, idx, eps):
    """subpoly_debug.deal_with_gradient_descent (:121-165) for the edges that need it.
    e0, e1, ints: [G,3]; plane: [G] column of the earlier plane each edge lies in.
    Returns (ints, d_new [G,2], bodies executed)."""
    e0 = np.ascontiguousarray(e0, np.float32).reshape(-1, 3)
    e1 = np.ascontiguousarray(e1, np.float32).reshape(-1, 3)
    x = np.array(ints, np.float32, copy=True).reshape(-1, 3)
    pl = np.ascontiguousarray(plane, np.int32).reshape(-1)
    d = np.zeros((x.shape[0], 2), np.float32)
    fn = lib().curve_gradient_descent
    fn.restype = ctypes.c_int
    n = fn(ctypes.byref(P._c), ctypes.c_void_p(e0.ctypes.data), ctypes.c_void_p(e1.ctypes.data),
           ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(pl.ctypes.data), ctypes.c_int32(int(idx)),
           ctypes.c_float(np.float32(eps)), ctypes.c_int64(x.shape[0]), ctypes.c_void_p(d.ctypes.data))
    return x, d, int(n)
