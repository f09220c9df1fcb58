"""Host logic of slab sharding (CPU): the slab plan and the merge of slab meshes."""
import numpy as np
import pytest
import torch

from helpers import canonical_triangles, canonical_vertices
from tropical import parallel


def test_slab_planes_cover_all_cells_once():
    for m, w in [(49, 2), (49, 3), (98, 4), (201, 8), (9, 8)]:
        planes = parallel.slab_planes(m, w)
        assert planes[0][0] == 0 and planes[-1][1] == m - 1
        for (a, b), (c, d) in zip(planes, planes[1:]):
            assert b == c and b > a and d > c   # neighbours share exactly one plane, every slab has a cell
    with pytest.raises(ValueError):
        parallel.slab_planes(5, 8)


def _grid_mesh(nx, ny, rng):
    """A triangulated height field over an nx x ny lattice, x = lattice column."""
    xs, ys = np.meshgrid(np.arange(nx, dtype=np.float32), np.arange(ny, dtype=np.float32), indexing="ij")
    v = np.stack([xs, ys, rng.random((nx, ny), dtype=np.float32)], -1).reshape(-1, 3)
    idx = np.arange(nx * ny).reshape(nx, ny)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    t = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
    return v, t


def _cut(v, t, lo, hi, last):
    """The part of the mesh with lo <= x <= hi, renumbered; tags as the slabs produce them."""
    keep_t = np.all((v[t][:, :, 0] >= lo) & (v[t][:, :, 0] <= hi), axis=1)
    used = np.zeros(len(v), bool)
    used[t[keep_t].reshape(-1)] = True
    remap = np.cumsum(used) - 1
    vv = v[used]
    tag = ((vv[:, 0] == lo) & (lo > 0)).astype(np.uint8) | (((vv[:, 0] == hi) & (not last)).astype(np.uint8) << 1)
    return torch.from_numpy(vv), torch.from_numpy(remap[t[keep_t]]), torch.from_numpy(tag)


@pytest.mark.parametrize("cuts", [[0, 3, 7], [0, 2, 4, 7], [0, 1, 2, 3, 4, 5, 6, 7]])
def test_merge_restores_the_mesh(cuts):
    rng = np.random.default_rng(0)
    v, t = _grid_mesh(8, 6, rng)
    parts = [_cut(v, t, cuts[i], cuts[i + 1], i == len(cuts) - 2) for i in range(len(cuts) - 1)]
    mv, mt, stats = parallel.merge_slab_meshes(parts)
    assert mv.shape[0] == v.shape[0]
    assert stats["shared_vertices"] == 6 * (len(cuts) - 2)
    assert np.array_equal(canonical_vertices(mv.numpy()), canonical_vertices(v))
    assert np.array_equal(canonical_triangles(mv.numpy(), mt.numpy()), canonical_triangles(v, t))


def test_merge_pairs_coincident_twins_in_order():
    # two coincident vertices on the shared plane (the reference's chunk-overlap duplicates)
    lower_v = torch.tensor([[0., 0, 0], [1, 0, 0], [1, 0, 0], [1, 1, 0]])
    lower_tag = torch.tensor([0, 2, 2, 2], dtype=torch.uint8)
    lower_t = torch.tensor([[0, 1, 3], [0, 2, 3]])
    upper_v = torch.tensor([[1., 0, 0], [1, 0, 0], [1, 1, 0], [2, 0, 0]])
    upper_tag = torch.tensor([1, 1, 1, 0], dtype=torch.uint8)
    upper_t = torch.tensor([[0, 3, 2], [1, 3, 2], [0, 1, 2]])   # the last one lies in the shared plane
    mv, mt, stats = parallel.merge_slab_meshes([(lower_v, lower_t, lower_tag), (upper_v, upper_t, upper_tag)])
    assert mv.shape[0] == 5 and stats["shared_vertices"] == 3
    assert mt.tolist() == [[0, 1, 3], [0, 2, 3], [1, 4, 3], [2, 4, 3]]


def test_merge_keeps_unmatched_plane_vertices():
    lower = (torch.tensor([[0., 0, 0], [1, 0, 0]]), torch.zeros((0, 3), dtype=torch.long), torch.tensor([0, 2], dtype=torch.uint8))
    upper = (torch.tensor([[1., 0, 0], [1, 5, 0]]), torch.zeros((0, 3), dtype=torch.long), torch.tensor([1, 1], dtype=torch.uint8))
    mv, mt, stats = parallel.merge_slab_meshes([lower, upper])
    assert mv.shape[0] == 3 and stats["shared_vertices"] == 1 and mt.shape[0] == 0
