"""`tropical.geometry` (reference: tropical/geometry.py) -- the functions the mesh-extraction path uses,
as stage-level calls on CUDA tensors.  `tropical.subpoly.subpoly` does not go through here: its step
kernels run the same device functions fused (csrc/curve.cuh, csrc/faces.cu).  The deprecated helpers of
the reference module (find_polygon*, cubic_bezier_roots, quad_roots*) are not part of the path and are
not provided.  No CPU route: tensors must live on the GPU.
"""
from typing import List

import torch
from torch import Tensor

from tropical import _native


def corner_points(expanded_edges: Tensor) -> Tensor:
    """Eight corner points of the box an edge spans (geometry.py:350-372).  expanded_edges [B, 2, 3] ->
    [B, 8, 3]; corner 4*i + 2*j + k takes x from endpoint k, y from endpoint j, z from endpoint i."""
    e = expanded_edges
    out = []
    for i in range(2):
        for j in range(2):
            for k in range(2):
                out.append(torch.stack([e[:, k, 0], e[:, j, 1], e[:, i, 2]], dim=-1))
    return torch.stack(out, dim=1)


def intersection_of_two_planes(p: Tensor, q: Tensor, plane="xz", eps=1e-6) -> Tensor:
    """Intersection of two (curved) planes with the edge's plane inside a trilinear cube (geometry.py:24-138).
    p, q [B, 8] corner distances -> [B, 3] trilinear coordinates, -1 where there is none."""
    assert "xz" == plane
    return _native.curve_intersections(p, q)


def batched_polynomial_roots(coeffs: Tensor, interval: List = [0, 1], eps=1e-9) -> Tensor:
    """The root geometry.intersection_of_two_planes keeps (geometry.py:259-300): for the quartic
    c0 x^4 + ... + c4 (coeffs [B, 5]) the largest real root in [0, 1] -- the last admissible eigenvalue of the
    companion matrix in LAPACK's order --, -1 if there is none.  Evaluated by solving the equivalent pair of
    planes p = (1, x), q from the coefficients is not possible in general, so this entry only accepts what the
    path feeds it: use intersection_of_two_planes."""
    raise _native.NativeError("batched_polynomial_roots has no stand-alone device entry: call intersection_of_two_planes")


def extract_triangles_from_sorted_vertices_and_mask(vertices: Tensor, mask: Tensor):
    """Fan triangles of angle-sorted padded face rows (geometry.py:536-556): vertices [B, M, 3], mask [B, M] ->
    numpy [T, 3, 3], ordered by fan step, then by row."""
    counts = mask.sum(-1)
    cumsum = counts.cumsum(0)
    first = torch.cat([counts.new_zeros(1), cumsum[:-1]], dim=0).long()
    flat = vertices[mask].view(-1, 3)
    faces = []
    for i in range(int(counts.max()) - 2 if counts.numel() else 0):
        ok = counts >= i + 3
        s = first[ok]
        faces.append(torch.stack([flat[s], flat[s + i + 1], flat[s + i + 2]], dim=1))
    if not faces:
        return flat.new_zeros((0, 3, 3)).cpu().numpy()
    return torch.cat(faces, dim=0).cpu().numpy()


def sort_polygon_vertices_batch(v: Tensor, n: Tensor, idx: int = 0, return_index: bool = False):
    """Sort the vertices of every padded face row by angle around the row's centre, seen against the normal n
    (geometry.py:483-525), and fan-triangulate.  v [B, M, 3] (norm 0 = padding), n [B, 3]."""
    if n is None:
        raise _native.NativeError("sort_polygon_vertices_batch: the path always passes the normals (subpoly.py:642)")
    order, valid = _native.polygon_order(v, n, idx)
    p = torch.gather(v, 1, order.unsqueeze(-1).expand(-1, -1, 3))
    faces = extract_triangles_from_sorted_vertices_and_mask(p, valid)
    if return_index:
        return faces, order
    return faces
