"""torchrun entry: one object sharded over the ranks (one slab per GPU), compared with the
single-GPU mesh on every rank.  Usage: torchrun --nproc-per-node N tests/slab_dist_check.py CASE"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tropical-nerf.pytorch_b200"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers import canonical_triangles, canonical_vertices, load_golden, native_net, oracle_net  # noqa: E402


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "small_sphere"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tropical import parallel
    N = native_net(oracle_net(load_golden(case)))
    v1, _, t1, _, _ = N.subpoly(size=1.2, eps=1e-4, force=True).read()
    for rep in range(2):
        v, t, stats = parallel.subpoly_sharded(N)
        assert np.array_equal(canonical_vertices(v.cpu().numpy()), canonical_vertices(v1.cpu().numpy())), "vertices differ"
        assert np.array_equal(canonical_triangles(v.cpu().numpy(), t.cpu().numpy()),
                              canonical_triangles(v1.cpu().numpy(), t1.cpu().numpy())), "triangles differ"
    # plane-sharded sweep (exact by construction): every array of the mesh equals the single-GPU one, on every rank
    want = [a.cpu().numpy() for a in N.subpoly(size=1.2, eps=1e-4, force=True).read()]
    for rep in range(2):
        got = [a.cpu().numpy() for a in parallel.subpoly_sweep_sharded(N).read()]
        for x, y in zip(got, want):
            assert np.array_equal(x, y), "plane-sharded sweep: mesh differs from the single-GPU mesh"
    dist.barrier()
    if dist.get_rank() == 0:
        print(f"slab_dist_check ok: {case} world {dist.get_world_size()} {stats}; sweep-sharded mesh identical")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
