"""CPU, world_size 2 over gloo: the host logic of the N>1 path (object sharding, max-over-ranks
timing, catalogue gather).  The device work of each rank is the single-GPU path tested in
test_gpu_parity.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tropical-nerf.pytorch_b200")]
    from tropical import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nets = [f"net{i}" for i in range(5)]
        catalogue = parallel.extract_many(nets, lambda n: {"name": n, "rank": rank, "V": 10 * len(n) + rank})
        slowest = parallel.max_over_ranks(1.0 + rank)
        out[rank] = (parallel.shard(5), sorted(catalogue), [catalogue[i]["rank"] for i in sorted(catalogue)], slowest)
        with pytest.raises(RuntimeError):
            parallel.gather_catalogue({0: {"rank": rank}})  # both ranks claim object 0
    finally:
        dist.destroy_process_group()


def test_two_ranks_share_the_objects():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] == [0, 1, 2] and res[1][0] == [3, 4]
    for r in range(world):
        assert res[r][1] == [0, 1, 2, 3, 4]          # every rank sees the whole catalogue
        assert res[r][2] == [0, 0, 0, 1, 1]          # and who extracted what
        assert res[r][3] == 2.0                      # time = the slowest rank


def test_shard_is_a_partition():
    from tropical.parallel import shard
    for n in (0, 1, 7, 8, 33):
        for w in (1, 2, 4, 8):
            parts = [shard(n, w, r) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_plane_ranges_deal_every_plane_once():
    from tropical.parallel import plane_ranges, slab_planes
    for m in (9, 49, 98, 201):
        for w in (1, 2, 3, 4, 8):
            r = plane_ranges(m, w)
            planes = [x for a, b in r for x in range(a, b + 1)]
            assert planes == list(range(m))                      # a partition, in rank order
            sizes = [b - a + 1 for a, b in r]
            assert max(sizes) - min(sizes) <= 1
            s = slab_planes(m, w)                                # cell slabs: neighbours share one plane
            assert s[0][0] == 0 and s[-1][1] == m - 1 and all(s[i][1] == s[i + 1][0] for i in range(w - 1))
    with pytest.raises(ValueError):
        plane_ranges(4, 5)
