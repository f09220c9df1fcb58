"""Subdivision-of-polygons mesh extraction (reference: tropical/subpoly.py), CUDA-backed.

`subpoly(net, d, size, eps, force)` keeps the reference's signature and return values
(faces, vertices, faces_with_indices); the work is one `tnb_subpoly` call.  For callers that
drive the stages themselves: `subpoly_` (one hyperplane), `extract_skeleton`, `extract_faces`,
`get_hypercube` map onto the stage-level C ABI.  The reference's internal grouping helpers
(`regions_to_vertices`, `r_idx_as_tensor`, `edge_vertices`, `mean_points_with_valid`,
`tensor_to_triangle_faces`: subpoly.py:281-535, :669-728) have no device counterpart of their
own: the kernels group by marks-grid cell instead of expanding and sorting region rows
(csrc/complex.cu, csrc/faces.cu), so those names are not provided.

Difference to the reference's return value, on purpose: `subpoly()`'s first value is the triangle
positions [T, 3, 3] of the de-duplicated faces (== vertices[faces_with_indices]); the reference
returns the fan of sort_polygon_vertices_batch over rows that may repeat a face
(its own TODO, subpoly.py:645).  `mesh.read()[4]` holds the sorted polygon rows.
"""
from typing import List, Tuple

import numpy as np
import torch
from torch import Tensor
from torch.nn import Module

from tropical import TropicalHashGrid, _native


@torch.no_grad()
def subpoly(net: Module, d: int, size: float, eps: float = 1e-4, force: bool = False,
            return_mesh: bool = False):
    """Subdivision polygons algorithm (subpoly.py:23-86).

    Returns (faces, vertices, faces_with_indices): faces [T,3,3] numpy triangle positions,
    vertices [V,3] device tensor, faces_with_indices [T,3] numpy vertex indices.
    force=True is the planar path the reference's entry point uses (train.py:50-51,127);
    force=False is the curve-approximation path (subpoly.py:120-183 + strict_check) with its
    gradient-descent repair (subpoly_debug.py:121-165); where the repair leaves an intersection off
    its planes the reference ends the process (subpoly.py:172-174) and this call raises."""
    mesh = net.native().subpoly(size=size, eps=eps, force=force)
    s = mesh.sizes()
    print()
    print(f"# of vertices and edges => {s['V']}/{s['E']}, {s['P']} faces", end=", ")
    vertices, _, tri, faces, _ = mesh.read()
    out = (faces.cpu().numpy(), vertices, tri.cpu().numpy())
    return out + (mesh,) if return_mesh else out


def subpoly_batch(nets, d: int, size: float, eps: float = 1e-4, force: bool = False, in_flight: int = 8):
    """subpoly() for a list of networks in one call (tnb_subpoly_batch): up to `in_flight` objects are extracted
    side by side on one GPU.  Not in the reference (it extracts one object per run, train.py:127); every entry of
    the returned list is what subpoly(net, d, size, eps, force) returns for that network."""
    meshes = _native.subpoly_batch([n.native() for n in nets], size=size, eps=eps, force=force, in_flight=in_flight)
    out = []
    for mesh in meshes:
        vertices, _, tri, faces, _ = mesh.read()
        out.append((faces.cpu().numpy(), vertices, tri.cpu().numpy()))
    return out


class _ComplexState:
    """Keeps the device-resident complex between subpoly_ calls so that the cached
    `outputs_` tensor the reference threads through (subpoly.py:92-95) is a handle."""

    def __init__(self, cx):
        self.cx = cx


def subpoly_(vertices, edges, net, l, h, eps, outputs_=None, pruning=True, strict=True,
             force=False):
    """One hyperplane (subpoly.py:90-279).  `outputs_` is either None (first call: the
    complex is built from `vertices`/`edges`) or the state object returned by the previous
    call.  Returns (vertices, edges, state)."""
    if not pruning:
        raise _native.NativeError("pruning=False is not supported on the device path")
    if isinstance(outputs_, _ComplexState):
        state = outputs_
    else:
        state = _ComplexState(net.native().complex_from_arrays(vertices, edges))
    state.cx.step(l, h, eps, force)
    v, e, _ = state.cx.read(outputs=False)
    return v, e, state


def extract_skeleton(vertices, edges, net, eps, outputs=None):
    """Vertices and edges of the complex that lie on the zero level set (subpoly.py:556-581):
    (vertices, edges, v_idx) with v_idx = the numbers of the kept vertices in the input.  `outputs` =
    the cached network rows [V, R] of the vertices (they carry the zeros of the failover override);
    evaluated on the device when None."""
    cx = net.native().complex_from_arrays(vertices, edges, outputs)
    mesh = cx.extract_mesh(eps)
    if mesh.sizes()["V"] == 0:
        return torch.Tensor([]).to(edges), torch.Tensor([]).to(edges), None
    v, e, _, _, _ = mesh.read()
    return v, e, mesh.read_vertex_index()


def extract_faces(vertices: Tensor, edges: Tensor, net: Module, outputs: Tensor = None, eps: float = None):
    """Faces of a surface skeleton (subpoly.py:584-652): (faces [T, 3, 3] numpy positions,
    faces_with_indices [T, 3] numpy vertex numbers).  `vertices` / `edges` are what extract_skeleton
    returned (every vertex on the surface and on an edge), so the numbering is kept."""
    if 0 == vertices.shape[0]:
        return [], []
    mesh = net.native().complex_from_arrays(vertices, edges, outputs).extract_mesh(net.eps if eps is None else eps)
    if mesh.sizes()["V"] != vertices.shape[0]:
        raise _native.NativeError("extract_faces expects the output of extract_skeleton (vertices off the surface or on no edge)")
    _, _, tri, faces, _ = mesh.read()
    return faces.cpu().numpy(), tri.cpu().numpy()


def extract_mesh(state: _ComplexState, net, eps=1e-4):
    """extract_skeleton + extract_faces on the device (subpoly.py:556-652)."""
    return state.cx.extract_mesh(eps)


def get_hypercube(d, size):
    """The 8 corners, 12 edges and 6 faces of (-size, size)^3 (subpoly.py:731-750)."""
    x = torch.Tensor([-size, size])
    vertices = torch.stack(torch.meshgrid(x, x, x, indexing="ij"), dim=-1).view(-1, 3)
    edges = [[i, j] for i in range(8) for j in range(i + 1, 8)
             if 1 == int((vertices[i] * vertices[j] < 0).sum())]
    faces = [[0, 3, 5, 1], [0, 2, 8, 4], [3, 4, 10, 7], [1, 2, 9, 6], [8, 9, 11, 10], [7, 11, 6, 5]]
    return vertices, torch.LongTensor(edges), faces
