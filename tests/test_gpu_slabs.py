"""Slab sharding of ONE object (csrc/halo.cuh, tropical/parallel.py): the mesh assembled from
slabs equals the single-GPU mesh -- same vertex positions bit for bit, same triangles -- for
every slab count.  `subpoly_slabs_local` runs all slabs on one device through the same kernels
and mailbox messages the multi-GPU run uses; `test_two_ranks_*` runs the real thing over
torch.distributed (NCCL + CUDA IPC) when two GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import canonical_triangles, canonical_vertices, load_golden, native_net, oracle_net

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sym_diff(a, b):
    sa, sb = {tuple(r) for r in a.tolist()}, {tuple(r) for r in b.tolist()}
    return len(sa ^ sb)


def _single(N):
    v, _, t, _, _ = N.subpoly(size=1.2, eps=1e-4, force=True).read()
    return v.cpu().numpy(), t.cpu().numpy()


@pytest.mark.parametrize("case", ["tiny_sphere_h8", "small_sphere", "small_torus"])
@pytest.mark.parametrize("slabs", [2, 3, 4, 8])
def test_slabs_on_one_device_equal_single(case, slabs):
    from tropical import parallel
    N = native_net(oracle_net(load_golden(case)))
    if slabs > N.n_marks - 1:
        pytest.skip("more slabs than cells")
    v1, t1 = _single(N)
    v, t, stats = parallel.subpoly_slabs_local(N, slabs)
    v, t = v.cpu().numpy(), t.cpu().numpy()
    assert stats["shared_vertices"] > 0
    near = stats["near_plane"]
    if near == 0:   # the exact case: no vertex within eps of a shared plane that only one slab holds
        assert v.shape == v1.shape and t.shape == t1.shape, (v.shape, v1.shape, t.shape, t1.shape, stats)
        assert np.array_equal(canonical_vertices(v), canonical_vertices(v1))
        assert np.array_equal(canonical_triangles(v, t), canonical_triangles(v1, t1))
    else:           # faces touching such a vertex from the other slab may differ: a handful, next to it
        assert _sym_diff(canonical_vertices(v), canonical_vertices(v1)) <= 4 * near
        assert _sym_diff(canonical_triangles(v, t), canonical_triangles(v1, t1)) <= 16 * near


def test_near_plane_indicator_explains_every_difference():
    # over all fixtures and slab counts: a difference from the single-GPU mesh never comes without the indicator
    from tropical import parallel
    exact = 0
    for case in ["tiny_sphere_h8", "small_sphere", "small_torus"]:
        N = native_net(oracle_net(load_golden(case)))
        v1, t1 = _single(N)
        for slabs in (2, 3, 5, 6):
            if slabs > N.n_marks - 1:
                continue
            v, t, stats = parallel.subpoly_slabs_local(N, slabs)
            same = np.array_equal(canonical_triangles(v.cpu().numpy(), t.cpu().numpy()), canonical_triangles(v1, t1))
            assert same or stats["near_plane"] > 0, (case, slabs, stats)
            exact += int(same)
    assert exact >= 6


def test_slabs_repeat_and_reuse_mailboxes():
    from tropical import parallel
    N = native_net(oracle_net(load_golden("small_sphere")))
    v1, t1 = _single(N)
    for _ in range(3):   # the cached mailboxes still hold the previous run's messages
        v, t, _ = parallel.subpoly_slabs_local(N, 3)
        assert np.array_equal(canonical_triangles(v.cpu().numpy(), t.cpu().numpy()), canonical_triangles(v1, t1))


def test_slab_of_empty_space():
    # 8 slabs of the small sphere: the outer slabs hold no surface at all and still take part
    from tropical import parallel
    N = native_net(oracle_net(load_golden("small_sphere")))
    v, t, stats = parallel.subpoly_slabs_local(N, 5)
    assert min(stats["slab_vertices"]) == 0 and stats["near_plane"] == 0
    v1, t1 = _single(N)
    assert np.array_equal(canonical_triangles(v.cpu().numpy(), t.cpu().numpy()), canonical_triangles(v1, t1))


def test_step_part_needs_halo():
    from tropical._native import NativeError
    N = native_net(oracle_net(load_golden("tiny_sphere_h8")))
    c = N.skeleton(128)
    with pytest.raises(NativeError):
        c.step_part(0, 0, 1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_equal_single():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "slab_dist_check.py"), "small_sphere"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "slab_dist_check ok" in out.stdout


@pytest.mark.parametrize("case,parts", [("small_sphere", 3), ("small_torus", 5)])
def test_sweep_assembled_from_plane_ranges_equals_the_whole_sweep(case, parts):
    """The plane-sharded sweep of parallel.subpoly_sweep_sharded on ONE device: the planes of the first axis are
    evaluated range by range (as the ranks would), |sdf| planes and per-chunk gradient maxima assembled into an
    empty whole-grid sweep: the same complex, and the same mesh, as tnb_skeleton + steps + extract."""
    from tropical.parallel import _hyperplanes, plane_ranges
    N = native_net(oracle_net(load_golden(case)))
    want_c = [t.cpu().numpy() for t in N.skeleton(128).read()]
    whole = N.skeleton_sweep_alloc(128)
    mg = None
    for a, b in plane_ranges(N.n_marks, parts):
        sw = N.skeleton_sweep(a, b, False, False, 128)
        whole.write_dist(sw.read_dist(), a, b)
        mg = sw.max_grad() if mg is None else torch.maximum(mg, sw.max_grad())
    whole.set_max_grad(mg)
    c = whole.finish()
    for x, y in zip([t.cpu().numpy() for t in c.read()], want_c):
        assert np.array_equal(x, y)
    c.steps(_hyperplanes(N))
    got = [t.cpu().numpy() for t in c.extract_mesh().read()]
    want = [t.cpu().numpy() for t in N.subpoly().read()]
    for x, y in zip(got, want):
        assert np.array_equal(x, y)
