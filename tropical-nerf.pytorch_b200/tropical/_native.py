"""ctypes binding of libtropical_b200.so (include/tropical_b200.h).

PyTorch is used for device memory and streams only: every function here takes torch
CUDA tensors, hands their raw pointers to the C ABI, and returns torch CUDA tensors.
There is no CPU implementation: importing works without a GPU (so the ABI can be
checked), every compute call raises `NativeError` when the library or a CUDA device is
missing.
"""
import ctypes
import os

import numpy as np
import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_ROOT, "lib", "libtropical_b200.so")

TNB_OK = 0
ERRORS = {-1: "TNB_ERR_INVALID", -2: "TNB_ERR_CUDA", -3: "TNB_ERR_CAPACITY", -4: "TNB_ERR_UNSUPPORTED"}


class NativeError(RuntimeError):
    pass


class NetDesc(ctypes.Structure):
    _fields_ = [("n_levels", ctypes.c_int32), ("n_features", ctypes.c_int32),
                ("log2_hashmap", ctypes.c_int32), ("base_resolution", ctypes.c_int32),
                ("per_level_scale", ctypes.c_double),
                ("num_layers", ctypes.c_int32), ("num_hidden", ctypes.c_int32),
                ("scale", ctypes.c_float), ("eps", ctypes.c_float),
                ("table", ctypes.c_void_p), ("table_len", ctypes.c_int64),
                ("mlp", ctypes.c_void_p), ("mlp_len", ctypes.c_int64),
                ("marks", ctypes.c_void_p), ("n_marks", ctypes.c_int32)]


class GridDesc(ctypes.Structure):
    _fields_ = [("n_levels", ctypes.c_int32), ("log2_hashmap", ctypes.c_int32),
                ("base_resolution", ctypes.c_int32), ("per_level_scale", ctypes.c_double)]


_lib = None

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I32 = ctypes.c_int32
_F = ctypes.c_float
# name -> (restype, argtypes); every symbol include/tropical_b200.h declares
SIGNATURES = {
    "tnb_last_error": (ctypes.c_char_p, []),
    "tnb_version": (ctypes.c_int, []),
    "tnb_device_count": (ctypes.c_int, []),
    "tnb_net_create": (ctypes.c_int, [ctypes.POINTER(NetDesc), ctypes.POINTER(_P)]),
    "tnb_net_destroy": (None, [_P]),
    "tnb_net_num_outputs": (ctypes.c_int, [_P]),
    "tnb_net_level_layout": (ctypes.c_int, [_P, _P, _P, _P, _P]),
    "tnb_grid_encode": (ctypes.c_int, [_P, _P, _I64, _P, _P]),
    "tnb_net_outputs": (ctypes.c_int, [_P, _P, _I64, _P, _P]),
    "tnb_net_sdf_grad": (ctypes.c_int, [_P, _P, _I64, _P, _P, _P]),
    "tnb_net_region": (ctypes.c_int, [_P, _P, _P, _I64, _F, _P, _P, _P, _P]),
    "tnb_sweep_signs": (ctypes.c_int, [_P, _P, _P, _P, _F, _P, _P]),
    "tnb_skeleton": (ctypes.c_int, [_P, _I32, _F, ctypes.POINTER(_P), _P]),
    "tnb_complex_from_arrays": (ctypes.c_int, [_P, _P, _I64, _P, _I64, ctypes.POINTER(_P), _P]),
    "tnb_complex_destroy": (None, [_P]),
    "tnb_complex_num_vertices": (_I64, [_P]),
    "tnb_complex_num_edges": (_I64, [_P]),
    "tnb_complex_read": (ctypes.c_int, [_P, _P, _P, _P, _P]),
    "tnb_subpoly_step": (ctypes.c_int, [_P, _P, _I32, _I32, _F, _I32, _P]),
    "tnb_subpoly_steps": (ctypes.c_int, [_P, _P, _P, _I32, _F, _I32, _P]),
    "tnb_set_cluster_max_items": (_I64, [_I64]),
    "tnb_set_fused_max_items": (_I64, [_I64]),
    "tnb_extract_mesh": (ctypes.c_int, [_P, _P, _F, ctypes.POINTER(_P), _P]),
    "tnb_mesh_destroy": (None, [_P]),
    "tnb_mesh_num_vertices": (_I64, [_P]),
    "tnb_mesh_num_edges": (_I64, [_P]),
    "tnb_mesh_num_triangles": (_I64, [_P]),
    "tnb_mesh_num_polygons": (_I64, [_P]),
    "tnb_mesh_polygon_width": (_I64, [_P]),
    "tnb_mesh_read": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "tnb_subpoly": (ctypes.c_int, [_P, _F, _F, _I32, _I32, ctypes.POINTER(_P), _P]),
    "tnb_mesh_read_host": (ctypes.c_int, [_P, _P, _P, _P, _P]),
    "tnb_set_capacity_factor": (ctypes.c_int, [ctypes.c_double]),
    "tnb_skeleton_sweep": (ctypes.c_int, [_P, _I32, _I32, _I32, _I32, _I32, ctypes.POINTER(_P), _P]),
    "tnb_sweep_destroy": (None, [_P]),
    "tnb_skeleton_sweep_alloc": (ctypes.c_int, [_P, _I32, ctypes.POINTER(_P), _P]),
    "tnb_sweep_num_planes": (_I64, [_P]),
    "tnb_sweep_read_dist": (ctypes.c_int, [_P, _P, _P]),
    "tnb_sweep_write_dist": (ctypes.c_int, [_P, _P, _I32, _I32, _P]),
    "tnb_sweep_num_chunks": (_I32, [_P]),
    "tnb_sweep_read_max_grad": (ctypes.c_int, [_P, _P, _P]),
    "tnb_sweep_write_max_grad": (ctypes.c_int, [_P, _P, _P]),
    "tnb_skeleton_finish": (ctypes.c_int, [_P, _P, ctypes.POINTER(_P), _P]),
    "tnb_mailbox_bytes": (_I64, [_I64]),
    "tnb_mailbox_create": (ctypes.c_int, [_I64, ctypes.POINTER(_P)]),
    "tnb_mailbox_destroy": (ctypes.c_int, [_P]),
    "tnb_mailbox_export": (ctypes.c_int, [_P, _P]),
    "tnb_mailbox_import": (ctypes.c_int, [_P, ctypes.POINTER(_P)]),
    "tnb_mailbox_release": (ctypes.c_int, [_P]),
    "tnb_complex_set_halo": (ctypes.c_int, [_P, _I32, _I32, ctypes.POINTER(_P), _I64, _I32, ctypes.c_uint32]),
    "tnb_subpoly_step_part": (ctypes.c_int, [_P, _P, _I32, _I32, _F, _I32, _I32, _P]),
    "tnb_extract_mesh_begin": (ctypes.c_int, [_P, _P, _F, ctypes.POINTER(_P), _P]),
    "tnb_extract_mesh_finish": (ctypes.c_int, [_P, _P, _P, _P]),
    "tnb_mesh_read_tags": (ctypes.c_int, [_P, _P, _P]),
    "tnb_mesh_read_vertex_index": (ctypes.c_int, [_P, _P, _P]),
    "tnb_complex_write_outputs": (ctypes.c_int, [_P, _P, _P, _P]),
    "tnb_net_outputs_group8": (ctypes.c_int, [_P, _P, _I64, _F, _P, _P, _P]),
    "tnb_curve_intersections": (ctypes.c_int, [_P, _P, _I64, _P, _P]),
    "tnb_net_forward": (ctypes.c_int, [_P, _P, _I64, _P, _P, _P]),
    "tnb_curve_gradient_descent": (ctypes.c_int, [_P, _P, _P, _P, _I32, _F, _I64, _P, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _P]),
    "tnb_polygon_order": (ctypes.c_int, [_P, _P, _I64, ctypes.c_int32, ctypes.c_int32, _P, _P, _P]),
    "tnb_mesh_near_plane": (_I64, [_P]),
    "tnb_grid_train_table_len": (_I64, [_P]),
    "tnb_grid_train_forward": (ctypes.c_int, [_P, _P, _P, _I64, _P, _P]),
    "tnb_grid_train_backward": (ctypes.c_int, [_P, _P, _P, _I64, _P, _P, _P, _P]),
    "tnb_grid_train_backward_backward": (ctypes.c_int, [_P, _P, _P, _I64, _P, _P, _P, _P, _P, _P]),
    "tnb_launch_count": (_I64, []),
    "tnb_launch_count_reset": (None, []),
    "tnb_release_cached_blocks": (None, []),
    "tnb_subpoly_batch": (ctypes.c_int, [ctypes.POINTER(_P), _I32, _F, _F, _I32, _I32, _I32, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_int32), _P]),
    "tnb_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "tnb_profile_read": (ctypes.c_int, [ctypes.c_int, _P, _P, _P, _P]),
    "tnb_profile_reset": (None, []),
}


def lib():
    """Load the CUDA library; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                "(nvcc, sm_100a). There is no CPU fallback for the mesh-extraction path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc != TNB_OK:
        msg = lib().tnb_last_error().decode("utf-8", "replace")
        raise NativeError(f"{ERRORS.get(rc, rc)}: {msg}")


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


PROFILE_CLASSES = ["sweep", "vertex_rows", "step_front", "step_back", "face_rows", "sign_sweep", "steps_persistent"]


def profile_enable(on=True):
    check(lib().tnb_profile_enable(int(bool(on))))


def profile_reset():
    lib().tnb_profile_reset()


def profile_read():
    """{class: (milliseconds, launches, units, algorithmic bytes)} of the library's per-kernel
    event timers."""
    out = {}
    for i, name in enumerate(PROFILE_CLASSES):
        ms, n, u, b = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(lib().tnb_profile_read(i, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(u), ctypes.byref(b)))
        out[name] = (ms.value, n.value, u.value, b.value)
    return out


def subpoly_batch(nets, size=1.2, eps=1e-4, force=True, unit=128, in_flight=8):
    """tnb_subpoly_batch: the whole path for a list of NativeNet in one call -> list of NativeMesh."""
    require_cuda()
    n = len(nets)
    handles = (_P * n)(*[x.handle for x in nets])
    out = (_P * n)()
    rcs = (ctypes.c_int32 * n)()
    rc = lib().tnb_subpoly_batch(handles, n, float(size), float(eps), int(bool(force)), int(unit), int(in_flight), out, rcs, _stream())
    meshes = [NativeMesh(nets[i], ctypes.c_void_p(out[i])) if out[i] else None for i in range(n)]
    check(rc)
    return meshes


def curve_intersections(p, q):
    """geometry.intersection_of_two_planes on [E, 8] corner values -> [E, 3] (device tensors)."""
    require_cuda()
    p, q = p.contiguous().float(), q.contiguous().float()
    out = torch.empty((p.shape[0], 3), dtype=torch.float32, device=p.device)
    check(lib().tnb_curve_intersections(_ptr(p), _ptr(q), p.shape[0], _ptr(out), _stream()))
    return out


def polygon_order(v, normals, base=0):
    """Ordering of geometry.sort_polygon_vertices_batch: v [B, M, 3], normals [B, 3] -> (order [B, M] int64,
    valid mask in that order [B, M] bool)."""
    require_cuda()
    v, normals = v.contiguous().float(), normals.contiguous().float()
    B, M = v.shape[0], v.shape[1]
    order = torch.empty((B, M), dtype=torch.int64, device=v.device)
    valid = torch.empty((B, M), dtype=torch.uint8, device=v.device)
    check(lib().tnb_polygon_order(_ptr(v), _ptr(normals), B, M, int(base), _ptr(order), _ptr(valid), _stream()))
    return order, valid.bool()


def require_cuda():
    if not torch.cuda.is_available() or lib().tnb_device_count() == 0:
        raise NativeError("no CUDA device: the mesh-extraction path has no CPU fallback")


class NativeNet:
    """Device-resident copy of one trilinear network (tnb_net)."""

    def __init__(self, levels, n_feat, log2_T, n_min, per_level_scale, num_layers, num_hidden,
                 table, mlp, marks, eps=1e-4, scale=1.0):
        require_cuda()
        table = np.ascontiguousarray(table, np.float32).reshape(-1)
        mlp = np.ascontiguousarray(mlp, np.float32).reshape(-1)
        marks = np.ascontiguousarray(marks, np.float32).reshape(-1)
        d = NetDesc(int(levels), int(n_feat), int(log2_T), int(n_min), float(per_level_scale),
                    int(num_layers), int(num_hidden), float(scale), float(eps),
                    table.ctypes.data, table.size, mlp.ctypes.data, mlp.size,
                    marks.ctypes.data, marks.size)
        h = ctypes.c_void_p()
        check(lib().tnb_net_create(ctypes.byref(d), ctypes.byref(h)))
        self.handle = h
        self.levels, self.n_feat = int(levels), int(n_feat)
        self.num_layers, self.num_hidden = int(num_layers), int(num_hidden)
        self.n_outputs = lib().tnb_net_num_outputs(h)
        self.eps, self.scale = float(eps), float(scale)
        self.n_marks = marks.size

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:
            _lib.tnb_net_destroy(h)

    def level_layout(self):
        L = self.levels
        scale = np.zeros(L, np.float32)
        res, size, off = (np.zeros(L, np.uint32) for _ in range(3))
        check(lib().tnb_net_level_layout(self.handle, scale.ctypes.data, res.ctypes.data,
                                         size.ctypes.data, off.ctypes.data))
        return scale, res, size, off

    # ---- network evaluation ---------------------------------------------------------
    def encode(self, xp):
        xp = xp.contiguous().float()
        n = xp.shape[0]
        enc = torch.empty((n, self.levels * self.n_feat), dtype=torch.float32, device=xp.device)
        check(lib().tnb_grid_encode(self.handle, _ptr(xp), n, _ptr(enc), _stream()))
        return enc

    def outputs(self, x):
        x = x.contiguous().float()
        n = x.shape[0]
        out = torch.empty((n, self.n_outputs), dtype=torch.float32, device=x.device)
        check(lib().tnb_net_outputs(self.handle, _ptr(x), n, _ptr(out), _stream()))
        return out

    def forward(self, x, rows=True):
        """Net.forward(x, gather=True): (rows [n, R] or None, raw last-layer output [n, 2])."""
        x = x.contiguous().float()
        n = x.shape[0]
        out = torch.empty((n, self.n_outputs), dtype=torch.float32, device=x.device) if rows else None
        raw = torch.empty((n, 2), dtype=torch.float32, device=x.device)
        check(lib().tnb_net_forward(self.handle, _ptr(x), n, _ptr(out), _ptr(raw), _stream()))
        return out, raw

    def outputs_group8(self, x, eps=None):
        """Net.forward(x, gather=True, group=8) (model.py:52-76): x [8 G, 3] -> (rows [8 G, R], raw [8 G, 2])."""
        x = x.contiguous().float()
        n = x.shape[0]
        if n % 8:
            raise NativeError("group=8 needs a multiple of 8 points")
        out = torch.empty((n, self.n_outputs), dtype=torch.float32, device=x.device)
        raw = torch.empty((n, 2), dtype=torch.float32, device=x.device)
        check(lib().tnb_net_outputs_group8(self.handle, _ptr(x), n // 8, float(self.eps if eps is None else eps),
                                           _ptr(out), _ptr(raw), _stream()))
        return out, raw

    def gradient_descent(self, edges, ints, plane, idx, eps):
        """subpoly_debug.deal_with_gradient_descent (:121-165) for the edges it selects: edges [G, 2, 3], ints [G, 3],
        plane [G] -> (ints after the loop, d_new [G, 2], steps taken, True if every edge ended within eps)."""
        edges = edges.contiguous().float()
        x = ints.contiguous().float().clone()
        plane = plane.contiguous().to(torch.int32)
        G = x.shape[0]
        d = torch.zeros((G, 2), dtype=torch.float32, device=x.device)
        bodies, ok = ctypes.c_int32(0), ctypes.c_int32(1)
        check(lib().tnb_curve_gradient_descent(self.handle, _ptr(edges), _ptr(x), _ptr(plane), int(idx), float(eps), G, _ptr(d),
                                               ctypes.byref(bodies), ctypes.byref(ok), _stream()))
        return x, d, bodies.value, bool(ok.value)

    def sdf_grad(self, x, want_grad=True):
        x = x.contiguous().float()
        n = x.shape[0]
        sdf = torch.empty(n, dtype=torch.float32, device=x.device)
        grad = torch.empty((n, 3), dtype=torch.float32, device=x.device) if want_grad else None
        check(lib().tnb_net_sdf_grad(self.handle, _ptr(x), n, _ptr(sdf), _ptr(grad), _stream()))
        return sdf, grad

    def region(self, x, outputs=None, eps=None, packed=False):
        x = x.contiguous().float()
        n = x.shape[0]
        eps = self.eps if eps is None else float(eps)
        if outputs is None:
            outputs = self.outputs(x)
        outputs = outputs.contiguous().float()
        signs = torch.empty((n, 3 + self.n_outputs), dtype=torch.int8, device=x.device)
        offset = torch.empty((n, 3), dtype=torch.int32, device=x.device)
        pk = torch.empty((n, 3), dtype=torch.int64, device=x.device) if packed else None
        check(lib().tnb_net_region(self.handle, _ptr(x), _ptr(outputs), n, eps, _ptr(signs),
                                   _ptr(offset), _ptr(pk), _stream()))
        if packed:
            return signs, offset, outputs, pk
        return signs, offset, outputs

    def sweep_signs(self, lo, hi, n, eps=None, out=None):
        lo = (ctypes.c_float * 3)(*lo)
        hi = (ctypes.c_float * 3)(*hi)
        nn = (ctypes.c_int32 * 3)(*n)
        count = int(n[0]) * int(n[1]) * int(n[2])
        if out is None:
            out = torch.empty((count, 2), dtype=torch.int64, device="cuda")
        eps = self.eps if eps is None else float(eps)
        check(lib().tnb_sweep_signs(self.handle, lo, hi, nn, eps, _ptr(out), _stream()))
        return out

    # ---- polyhedral complex -----------------------------------------------------------
    def skeleton(self, unit=128, size=1.2):
        h = ctypes.c_void_p()
        check(lib().tnb_skeleton(self.handle, int(unit), float(size), ctypes.byref(h), _stream()))
        return NativeComplex(self, h)

    def complex_from_arrays(self, vertices, edges, outputs=None):
        """Complex from caller arrays; `outputs` [V, R] = cached network rows that replace the evaluated ones."""
        vertices = vertices.contiguous().float()
        edges = edges.contiguous().long()
        h = ctypes.c_void_p()
        check(lib().tnb_complex_from_arrays(self.handle, _ptr(vertices), vertices.shape[0], _ptr(edges),
                                            edges.shape[0], ctypes.byref(h), _stream()))
        c = NativeComplex(self, h)
        if outputs is not None and vertices.shape[0]:
            outputs = outputs.contiguous().float()
            assert outputs.shape == (vertices.shape[0], self.n_outputs)
            check(lib().tnb_complex_write_outputs(self.handle, c.handle, _ptr(outputs), _stream()))
        return c

    def skeleton_sweep_alloc(self, unit=128):
        """The whole grid's sweep with nothing evaluated (filled by write_dist / set_max_grad)."""
        h = ctypes.c_void_p()
        check(lib().tnb_skeleton_sweep_alloc(self.handle, int(unit), ctypes.byref(h), _stream()))
        return NativeSweep(self, h)

    def skeleton_sweep(self, x_lo, x_hi, shared_lower, shared_upper, unit=128):
        """First half of the skeleton of the slab of marks-grid planes [x_lo, x_hi]."""
        h = ctypes.c_void_p()
        check(lib().tnb_skeleton_sweep(self.handle, int(unit), int(x_lo), int(x_hi), int(bool(shared_lower)),
                                       int(bool(shared_upper)), ctypes.byref(h), _stream()))
        return NativeSweep(self, h)

    def subpoly(self, size=1.2, eps=1e-4, force=True, unit=128):
        """The whole path (tnb_subpoly); returns a NativeMesh."""
        h = ctypes.c_void_p()
        check(lib().tnb_subpoly(self.handle, float(size), float(eps), int(bool(force)), int(unit),
                                ctypes.byref(h), _stream()))
        return NativeMesh(self, h)


class NativeSweep:
    """|sdf| and per-chunk max |grad| of one marks-grid slab (tnb_sweep)."""

    def __init__(self, net, handle):
        self.net, self.handle = net, handle
        self.n_chunks = int(lib().tnb_sweep_num_chunks(handle))

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:
            _lib.tnb_sweep_destroy(h)

    def max_grad(self):
        out = torch.empty(self.n_chunks, dtype=torch.float32, device="cuda")
        check(lib().tnb_sweep_read_max_grad(self.handle, _ptr(out), _stream()))
        return out

    def set_max_grad(self, t):
        t = t.contiguous().float()
        assert t.numel() == self.n_chunks
        check(lib().tnb_sweep_write_max_grad(self.handle, _ptr(t), _stream()))

    def read_dist(self, out=None):
        """|sdf| of the sweep's planes: [planes * M * M] floats, first plane first."""
        n = int(lib().tnb_sweep_num_planes(self.handle)) * self.net.n_marks ** 2
        if out is None:
            out = torch.empty(n, dtype=torch.float32, device="cuda")
        assert out.numel() >= n and out.is_contiguous()
        check(lib().tnb_sweep_read_dist(self.handle, _ptr(out), _stream()))
        return out

    def write_dist(self, t, x_lo, x_hi):
        assert t.is_contiguous() and t.numel() >= (x_hi - x_lo + 1) * self.net.n_marks ** 2
        check(lib().tnb_sweep_write_dist(self.handle, _ptr(t), int(x_lo), int(x_hi), _stream()))

    def finish(self):
        h = ctypes.c_void_p()
        check(lib().tnb_skeleton_finish(self.net.handle, self.handle, ctypes.byref(h), _stream()))
        return NativeComplex(self.net, h)


class Mailbox:
    """Device memory the slab exchange writes into (tnb_mailbox_*): created here, or mapped
    from another process through its 64-byte CUDA IPC handle."""

    def __init__(self, payload=None, handle=None):
        self.owned = handle is None
        p = ctypes.c_void_p()
        if self.owned:
            check(lib().tnb_mailbox_create(int(payload), ctypes.byref(p)))
        else:
            buf = ctypes.create_string_buffer(bytes(handle), 64)
            check(lib().tnb_mailbox_import(buf, ctypes.byref(p)))
        self.ptr = p

    def export(self):
        buf = ctypes.create_string_buffer(64)
        check(lib().tnb_mailbox_export(self.ptr, buf))
        return bytes(buf.raw)

    def close(self):
        p, self.ptr = getattr(self, "ptr", None), None
        if p and _lib is not None:
            (_lib.tnb_mailbox_destroy if self.owned else _lib.tnb_mailbox_release)(p)

    def __del__(self):
        self.close()


class NativeComplex:
    """Device-resident vertices / edges / cached outputs (tnb_complex)."""

    def __init__(self, net, handle):
        self.net, self.handle = net, handle

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:
            _lib.tnb_complex_destroy(h)

    @property
    def num_vertices(self):
        return self._size(lib().tnb_complex_num_vertices(self.handle))

    @property
    def num_edges(self):
        return self._size(lib().tnb_complex_num_edges(self.handle))

    @staticmethod
    def _size(n):
        if n < 0:   # the complex carries a latched device-side error
            msg = lib().tnb_last_error().decode("utf-8", "replace")
            code = "TNB_ERR_CAPACITY" if "too small" in msg or "payload" in msg else "TNB_ERR"
            raise NativeError(f"{code}: {msg}")
        return int(n)

    def read(self, vertices=True, edges=True, outputs=True):
        V, E, R = self.num_vertices, self.num_edges, self.net.n_outputs
        v = torch.empty((V, 3), dtype=torch.float32, device="cuda") if vertices else None
        e = torch.empty((E, 2), dtype=torch.int64, device="cuda") if edges else None
        o = torch.empty((V, R), dtype=torch.float32, device="cuda") if outputs else None
        check(lib().tnb_complex_read(self.handle, _ptr(v), _ptr(e), _ptr(o), _stream()))
        return v, e, o

    def step(self, l, h, eps=1e-4, force=True):
        check(lib().tnb_subpoly_step(self.net.handle, self.handle, int(l), int(h), float(eps),
                                     int(bool(force)), _stream()))
        return self

    def steps(self, lh, eps=1e-4, force=True):
        """All hyperplanes `lh` = [(l, h), ...] in one call (tnb_subpoly_steps)."""
        arr = (ctypes.c_int32 * (2 * len(lh)))(*[int(x) for pair in lh for x in pair])
        check(lib().tnb_subpoly_steps(self.net.handle, self.handle, arr, len(lh), float(eps),
                                      int(bool(force)), _stream()))
        return self

    def extract_mesh(self, eps=1e-4):
        h = ctypes.c_void_p()
        check(lib().tnb_extract_mesh(self.net.handle, self.handle, float(eps), ctypes.byref(h), _stream()))
        return NativeMesh(self.net, h)

    # ---- slab sharding ----------------------------------------------------------------
    def set_halo(self, rank, world, boxes, payload, timeout_ms=0, seq0=0):
        arr = (ctypes.c_void_p * world)(*[b.ptr for b in boxes])
        self._boxes = list(boxes)  # keep the mailboxes alive as long as the complex
        check(lib().tnb_complex_set_halo(self.handle, int(rank), int(world), arr, int(payload), int(timeout_ms), int(seq0)))

    def step_part(self, l, h, part, eps=1e-4, force=True):
        check(lib().tnb_subpoly_step_part(self.net.handle, self.handle, int(l), int(h), float(eps),
                                          int(bool(force)), int(part), _stream()))

    def extract_mesh_begin(self, eps=1e-4):
        h = ctypes.c_void_p()
        check(lib().tnb_extract_mesh_begin(self.net.handle, self.handle, float(eps), ctypes.byref(h), _stream()))
        return NativeMesh(self.net, h)

    def extract_mesh_finish(self, mesh):
        check(lib().tnb_extract_mesh_finish(self.net.handle, self.handle, mesh.handle, _stream()))
        return mesh


class NativeMesh:
    """Extracted surface mesh (tnb_mesh)."""

    def __init__(self, net, handle):
        self.net, self.handle = net, handle

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and _lib is not None:
            _lib.tnb_mesh_destroy(h)

    def sizes(self):
        L = lib()
        return dict(V=int(L.tnb_mesh_num_vertices(self.handle)), E=int(L.tnb_mesh_num_edges(self.handle)),
                    T=int(L.tnb_mesh_num_triangles(self.handle)), P=int(L.tnb_mesh_num_polygons(self.handle)),
                    W=int(L.tnb_mesh_polygon_width(self.handle)))

    def read(self):
        """Device tensors: vertices [V,3], edges [E,2], triangles [T,3], faces [T,3,3],
        polygons [P,W]."""
        s = self.sizes()
        dev = "cuda"
        v = torch.empty((s["V"], 3), dtype=torch.float32, device=dev)
        e = torch.empty((s["E"], 2), dtype=torch.int64, device=dev)
        t = torch.empty((s["T"], 3), dtype=torch.int64, device=dev)
        f = torch.empty((s["T"], 3, 3), dtype=torch.float32, device=dev)
        p = torch.empty((s["P"], s["W"]), dtype=torch.int64, device=dev)
        check(lib().tnb_mesh_read(self.handle, _ptr(v), _ptr(e), _ptr(t), _ptr(f), _ptr(p), _stream()))
        return v, e, t, f, p

    def read_vertex_index(self):
        """int64 per mesh vertex: its row in the complex the mesh came from (extract_skeleton's v_idx)."""
        t = torch.empty(self.sizes()["V"], dtype=torch.int64, device="cuda")
        if t.numel():
            check(lib().tnb_mesh_read_vertex_index(self.handle, _ptr(t), _stream()))
        return t

    @property
    def near_plane(self):
        return int(lib().tnb_mesh_near_plane(self.handle))

    def read_tags(self):
        """uint8 per vertex: bit0 / bit1 = on the slab plane shared with the lower / upper neighbour."""
        t = torch.empty(self.sizes()["V"], dtype=torch.uint8, device="cuda")
        if t.numel():
            check(lib().tnb_mesh_read_tags(self.handle, _ptr(t), _stream()))
        return t

    def read_host(self, polygons=True, out=None):
        """numpy arrays through tnb_mesh_read_host (host buffers): vertices, triangles,
        faces (what the reference's subpoly() returns) and, optionally, the polygon rows.
        `out` = {"v": float32 buffer, "t": int64 buffer, "f": float32 buffer} of flat host arrays at least as
        large (e.g. views of pinned memory, reused from call to call): the results are views into them."""
        s = self.sizes()
        if out is not None:
            v = out["v"][:s["V"] * 3].reshape(s["V"], 3)
            t = out["t"][:s["T"] * 3].reshape(s["T"], 3)
            f = out["f"][:s["T"] * 9].reshape(s["T"], 3, 3)
            assert v.dtype == np.float32 and t.dtype == np.int64 and f.dtype == np.float32
        else:
            v = np.empty((s["V"], 3), np.float32)
            t = np.empty((s["T"], 3), np.int64)
            f = np.empty((s["T"], 3, 3), np.float32)
        p = np.empty((s["P"], s["W"]), np.int64) if polygons else None
        check(lib().tnb_mesh_read_host(self.handle, v.ctypes.data, t.ctypes.data, f.ctypes.data,
                                       p.ctypes.data if polygons else None))
        return v, t, f, p
