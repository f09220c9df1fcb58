"""Subdivision-of-polygons mesh extraction (reference: tropical/subpoly.py), CUDA-backed.

`subpoly(net, d, size, eps, force)` keeps the reference's signature and return values
(faces, vertices, faces_with_indices); the work is one `tnb_subpoly` call.  The stage-level
functions (`subpoly_`, `extract_skeleton`, `extract_faces`, `edge_vertices`, ...) are kept
for callers that drive the stages themselves and map onto the stage-level C ABI.
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
    force=False is the curve-approximation path (subpoly.py:120-183 + strict_check); its
    gradient-descent repair (subpoly_debug.py:121-165) is not built and raises if needed."""
    mesh = net.native().subpoly(size=size, eps=eps, force=force)
    s = mesh.sizes()
    print()
    print(f"# of vertices and edges => {s['V']}/{s['E']}, {s['P']} faces", end=", ")
    vertices, _, tri, faces, _ = mesh.read()
    out = (faces.cpu().numpy(), vertices, tri.cpu().numpy())
    return out + (mesh,) if return_mesh else out


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
