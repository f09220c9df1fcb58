"""`tropical.subpoly_debug` (reference: tropical/subpoly_debug.py) -- the one routine of that module with
work of its own on the mesh-extraction path, as a stage-level call on CUDA tensors.  `tropical.subpoly.subpoly`
does not go through here: its step kernels run the same device functions fused (csrc/repair.cuh).  The checks
of the reference module that only print (check_new_vertices*, debug_test_idx) are not provided;
check_edges_with_new_vertices and strict_check are part of the step kernels (csrc/complex.cu)."""
import torch

from tropical import _native


def deal_with_gradient_descent(c, d_new, e, eps, gg, idx, inds, ints, net):
    """subpoly_debug.py:121-165.  Intersections that are admissible (~gg) but farther than eps from one of
    their two planes walk down the gradient of d0^2 + d1^2, all of them while any of them is still off, at
    most 500 steps.  Same arguments and return value as the reference: c mask of the curved edges in e
    [E, 2, 3], d_new [Ec, 2], gg [Ec], inds [Ec, 2] (column 1 = the earlier plane's output column), ints
    [Ec, 3]; returns (ints, d_new), updated in place like the reference."""
    gd = ~gg & (0 < (d_new.abs() > eps).sum(dim=-1))
    if 0 < gd.sum():
        native = net.native() if hasattr(net, "native") else net
        x, d, _, _ = native.gradient_descent(e[c][gd], ints[gd], inds[gd, 1], idx, eps)
        ints[gd] = x
        d_new[gd, 0] = d[:, 0]
        d_new[gd, 1] = d[:, 1]
    return ints, d_new
