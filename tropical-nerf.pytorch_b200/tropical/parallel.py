"""Multi-GPU plumbing for the extraction path: one process per GPU over torch.distributed.

The path shards by OBJECT: every rank extracts the networks assigned to it and no
collective touches the data path.  What needs agreement across ranks is only
  * which objects a rank owns            -> `shard`
  * the time of the slowest rank         -> `max_over_ranks`
  * the catalogue of the extracted meshes -> `gather_catalogue`
All functions work with any backend (NCCL on the B200 box, gloo in the CPU tests).
"""
from typing import Any, Callable, Dict, List, Sequence

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(n_objects: int, world_size: int = None, r: int = None) -> List[int]:
    """Object numbers owned by rank r: contiguous blocks whose sizes differ by at most one."""
    world_size = world() if world_size is None else world_size
    r = rank() if r is None else r
    base, extra = divmod(n_objects, world_size)
    start = r * base + min(r, extra)
    return list(range(start, start + base + (1 if r < extra else 0)))


def max_over_ranks(value: float, device: str = None) -> float:
    """Largest value over all ranks (device times are reported as the max over ranks)."""
    if world() == 1:
        return float(value)
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_catalogue(local: Dict[int, Dict[str, Any]]) -> Dict[int, Dict[str, Any]]:
    """Every rank contributes {object number: summary}; every rank gets the union.  Raises if
    two ranks claim the same object."""
    if world() == 1:
        return dict(local)
    parts: List[Dict[int, Dict[str, Any]]] = [None] * world()
    dist.all_gather_object(parts, local)
    merged: Dict[int, Dict[str, Any]] = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise RuntimeError(f"object {k} extracted by two ranks")
            merged[k] = v
    return merged


def extract_many(nets: Sequence[Any], extract: Callable[[Any], Dict[str, Any]]) -> Dict[int, Dict[str, Any]]:
    """Extract every network of `nets` once across the ranks; `extract(net)` returns the
    summary to publish (e.g. mesh sizes).  Returns the full catalogue on every rank."""
    mine = {i: extract(nets[i]) for i in shard(len(nets))}
    return gather_catalogue(mine)


# ==================================================================================================
# Slab sharding of ONE object (BASELINE.json north_star: "each GPU owns a cell slab")
# ==================================================================================================
# The marks grid is cut along its first axis into `world` slabs of cells; neighbouring slabs
# share one plane of grid vertices.  Each rank runs skeleton -> 33 hyperplane steps -> faces on
# its slab; what crosses a shared plane travels once per step, written by the step's own kernels
# into the peers' mailboxes (csrc/halo.cuh).  torch.distributed is used for set-up only: the
# per-chunk max |grad| of the skeleton (one MAX all-reduce), the mailbox handles, and the
# all-gather of the slab meshes at merge time.
_RUNS = 0
_LOCAL_BOXES = {}
_DIST_BOXES = {}
DEFAULT_PAYLOAD = 1 << 22


def slab_planes(n_marks: int, world_size: int) -> List[tuple]:
    """[(x_lo, x_hi)] per rank: near-equal contiguous cell ranges; rank r and r+1 share plane x_hi(r)."""
    cells = n_marks - 1
    if world_size < 1 or world_size > cells:
        raise ValueError(f"cannot cut {cells} cells into {world_size} slabs")
    bounds = [(r * cells) // world_size for r in range(world_size + 1)]
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def _hyperplanes(net):
    nl, h = net.num_layers, net.num_hidden
    return [(l, k) for l in range(nl - 1) for k in range(h)] + [(nl - 2, h)]


def _occurrence(g: torch.Tensor) -> torch.Tensor:
    """For group numbers g [N]: the running index of every element inside its group (input order)."""
    n = g.numel()
    if n == 0:
        return g.clone()
    order = torch.argsort(g, stable=True)
    sg = g[order]
    pos = torch.arange(n, device=g.device)
    first = torch.cat([torch.ones(1, dtype=torch.bool, device=g.device), sg[1:] != sg[:-1]])
    start = torch.cummax(torch.where(first, pos, torch.zeros_like(pos)), 0)[0]
    occ = torch.empty_like(pos)
    occ[order] = pos - start
    return occ


def merge_slab_meshes(parts: Sequence[tuple]):
    """parts[r] = (vertices [V,3] f32, triangles [T,3] i64, tags [V] u8) of slab r, in slab order.

    Vertices on a shared plane exist on both neighbours with bit-identical positions (same
    arithmetic on the same in-plane edges): the upper slab's copy is dropped and its triangles
    are rewired to the lower slab's vertex.  Coincident twins (the reference's chunk-overlap
    duplicates) are paired in order.  Returns (vertices, triangles, stats); vertex numbering is
    slab-major, not the single-GPU numbering."""
    maps, verts, tris = [], [], []
    offset = 0
    shared = 0
    for r, (v, t, tag) in enumerate(parts):
        n = v.shape[0]
        matched = torch.zeros(n, dtype=torch.bool, device=v.device)
        target = torch.zeros(n, dtype=torch.long, device=v.device)
        if r > 0 and n > 0:
            lv, _, ltag = parts[r - 1]
            ui = torch.nonzero(ltag & 2).flatten()       # lower slab: vertices on the shared plane
            gi = torch.nonzero(tag & 1).flatten()        # this slab: its copies
            if ui.numel() and gi.numel():
                ub = lv[ui].contiguous().view(torch.int32)
                gb = v[gi].contiguous().view(torch.int32)
                _, inv = torch.unique(torch.cat([ub, gb]), dim=0, return_inverse=True)
                gu, gg = inv[:ui.numel()], inv[ui.numel():]
                ou, og = _occurrence(gu), _occurrence(gg)
                width = int(max(ou.max().item(), og.max().item())) + 1
                ku, kg = gu * width + ou, gg * width + og
                ks, perm = torch.sort(ku)
                at = torch.searchsorted(ks, kg).clamp(max=ks.numel() - 1)
                hit = ks[at] == kg
                matched[gi[hit]] = True
                target[gi[hit]] = maps[r - 1][ui[perm[at[hit]]]]
        own = ~matched
        ids = offset + torch.cumsum(own.long(), 0) - 1
        m = torch.where(own, ids, target)
        maps.append(m)
        verts.append(v[own])
        if t.shape[0]:
            dup = matched[t].all(dim=1)                  # a face inside the shared plane: the lower slab emits it
            tris.append(m[t[~dup]])
        n_own = int(own.sum().item())
        shared += n - n_own
        offset += n_own
    vertices = torch.cat(verts) if verts else torch.zeros((0, 3))
    triangles = torch.cat(tris) if tris else torch.zeros((0, 3), dtype=torch.long, device=vertices.device)
    return vertices, triangles, {"shared_vertices": shared, "slabs": len(parts)}


def _n_chunks(n_marks: int, unit: int) -> int:
    """Chunks of the skeleton sweep, as range(0, L, unit - 1) enumerates them per axis (tropical.py:176-181)."""
    return len(range(0, n_marks, unit - 1)) ** 3


def _run_slabs(net, mine, world_size, boxes_of, reduce_max, sum_int, eps, unit, payload, seq0, timeout_ms, marks=None):
    """Skeleton, hyperplane steps and face extraction of the slabs `mine` (rank numbers) on the
    current device and stream.  Several slabs per process run in lock step: every slab posts its
    messages before any slab waits for them.

    The two collectives of the set-up (`reduce_max`, `sum_int`) are reached by EVERY rank exactly once
    and in this order, whatever fails locally: a rank whose sweep or edge selection raised contributes
    neutral values and a failure flag (folded into `sum_int` as a large negative number), so that no rank is
    left alone in a collective.  From the steps on nothing but the mailboxes connects the ranks, and
    their receive has its own timeout and status word (csrc/halo.cuh).  Returns (meshes, error)."""
    from . import _native

    def mark(name):   # phase boundaries on the current stream (subpoly_sharded(..., phase_ms=True))
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    planes = slab_planes(net.n_marks, world_size)
    err, sweeps, cs = None, [], []
    mark("start")
    try:
        sweeps = [net.skeleton_sweep(planes[r][0], planes[r][1], r > 0, r < world_size - 1, unit) for r in mine]
        mg = sweeps[0].max_grad()
        for sw in sweeps[1:]:
            mg = torch.maximum(mg, sw.max_grad())
    except _native.NativeError as e:
        err, mg = e, torch.zeros(_n_chunks(net.n_marks, unit), dtype=torch.float32, device="cuda")
    mark("sweep")
    mg = reduce_max(mg)
    mark("max_grad_all_reduce")
    edges = 0
    if err is None:
        try:
            for sw in sweeps:
                sw.set_max_grad(mg)
                cs.append(sw.finish())
            edges = sum(c.num_edges for c in cs)
        except _native.NativeError as e:
            err = e
    del sweeps
    mark("skeleton_finish")
    FAIL = -(1 << 40)
    total = sum_int(FAIL if err is not None else edges)
    mark("edge_count_all_reduce")
    if total < 0:      # somebody failed during the set-up: every rank leaves here, together
        return None, (err or _native.NativeError("another rank failed during the set-up of the sharded extraction"))
    if total == 0:
        return None, None  # empty skeleton everywhere: the caller takes the hypercube route (subpoly.py:51-52)
    try:
        for r, c in zip(mine, cs):
            c.set_halo(r, world_size, boxes_of(r), payload, timeout_ms, seq0)
        for l, h in _hyperplanes(net):
            if len(cs) == 1:
                cs[0].step_part(l, h, 0, eps, True)
            else:
                for c in cs:
                    c.step_part(l, h, 1, eps, True)
                for c in cs:
                    c.step_part(l, h, 2, eps, True)
        mark("hyperplanes_with_exchanges")
        meshes = [c.extract_mesh_begin(eps) for c in cs]
        for c, m in zip(cs, meshes):
            c.extract_mesh_finish(m)
        mark("faces")
    except _native.NativeError as e:
        return None, e
    return meshes, None


def _read_part(mesh):
    v, _, t, _, _ = mesh.read()
    return v, t, mesh.read_tags()


def subpoly_slabs_local(net, n_slabs: int, size: float = 1.2, eps: float = 1e-4, unit: int = 128,
                        payload: int = DEFAULT_PAYLOAD, return_parts: bool = False):
    """One object cut into `n_slabs` slabs that all run on THIS device (the same kernels and
    messages as the multi-GPU run, with the host ordering sends before receives).  Returns
    (vertices, triangles, stats)."""
    from . import _native
    global _RUNS
    key = (n_slabs, payload, torch.cuda.current_device())
    if key not in _LOCAL_BOXES:
        _LOCAL_BOXES[key] = [_native.Mailbox(payload=payload) for _ in range(n_slabs)]
    boxes = _LOCAL_BOXES[key]
    factor = 4.0
    try:
        for attempt in range(4):
            _RUNS += 1
            meshes, e = _run_slabs(net, list(range(n_slabs)), n_slabs, lambda r: boxes, lambda t: t, lambda x: x,
                                   eps, unit, payload, (_RUNS * 4096) & 0xFFFFFF, 200)
            if e is None:
                break
            if "TNB_ERR_CAPACITY" not in str(e) or attempt == 3:
                raise e
            factor *= 2.0
            _native.check(_native.lib().tnb_set_capacity_factor(factor))
    finally:
        _native.check(_native.lib().tnb_set_capacity_factor(4.0))
    if meshes is None:
        v, _, t, _, _ = net.subpoly(size=size, eps=eps, force=True, unit=unit).read()
        return v, t, {"shared_vertices": 0, "slabs": 1, "hypercube": True}
    parts = [_read_part(m) for m in meshes]
    v, t, stats = merge_slab_meshes(parts)
    stats["slab_vertices"] = [int(p[0].shape[0]) for p in parts]
    stats["near_plane"] = sum(m.near_plane for m in meshes)
    if return_parts:
        stats["parts"] = parts
    return v, t, stats


def _gather_rows(x: torch.Tensor, group=None) -> List[torch.Tensor]:
    """all-gather of tensors whose first dimension differs per rank."""
    w = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], dtype=torch.long, device=x.device)
    sizes = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = torch.zeros((max(max(sizes), 1),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[:x.shape[0]] = x
    out = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(out, pad, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


def _dist_boxes(payload, group=None):
    """This rank's mailbox plus every peer's, mapped through CUDA IPC (set up once per process)."""
    from . import _native
    key = (payload, torch.cuda.current_device(), dist.get_world_size(group))
    if key not in _DIST_BOXES:
        own = _native.Mailbox(payload=payload)
        handles = [None] * dist.get_world_size(group)
        dist.all_gather_object(handles, own.export(), group=group)
        me = dist.get_rank(group)
        _DIST_BOXES[key] = [own if r == me else _native.Mailbox(handle=h) for r, h in enumerate(handles)]
    return _DIST_BOXES[key]


def subpoly_sharded(net, size: float = 1.2, eps: float = 1e-4, unit: int = 128, payload: int = DEFAULT_PAYLOAD,
                    group=None, gather: bool = True, timeout_ms: int = 2000, phase_ms: bool = False):
    """One object sharded over the ranks of `group` by marks-grid slabs (one slab per GPU).  Every
    rank calls this with the same network.  Returns (vertices, triangles, stats): the merged mesh
    on every rank when `gather`, else this rank's slab (vertices, triangles, tags)."""
    from . import _native
    global _RUNS
    if world() == 1:
        v, _, t, _, _ = net.subpoly(size=size, eps=eps, force=True, unit=unit).read()
        return v, t, {"shared_vertices": 0, "slabs": 1}
    w, me = dist.get_world_size(group), dist.get_rank(group)
    boxes = _dist_boxes(payload, group)

    def reduce_max(t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return t

    def sum_int(x):
        t = torch.tensor([int(x)], dtype=torch.long, device="cuda")
        dist.all_reduce(t, group=group)
        return int(t.item())

    factor = 4.0
    try:
        for attempt in range(4):
            _RUNS += 1
            err, msg, part = 0, "", None
            marks = [] if phase_ms else None
            meshes, e = _run_slabs(net, [me], w, lambda r: boxes, reduce_max, sum_int, eps, unit, payload,
                                   (_RUNS * 4096) & 0xFFFFFF, timeout_ms, marks)
            if e is None and meshes is not None:
                try:
                    part = _read_part(meshes[0])
                except _native.NativeError as e2:
                    e = e2
            if e is not None:
                err, msg = (1 if "TNB_ERR_CAPACITY" in str(e) else 2), str(e)
            t = torch.tensor([err], dtype=torch.long, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            worst = int(t.item())
            if worst == 0:
                break
            if worst == 2 or attempt == 3:
                raise _native.NativeError(msg if err else "another rank failed during the sharded extraction")
            factor *= 2.0
            _native.check(_native.lib().tnb_set_capacity_factor(factor))
    finally:
        _native.check(_native.lib().tnb_set_capacity_factor(4.0))
    if meshes is None:
        v, _, t, _, _ = net.subpoly(size=size, eps=eps, force=True, unit=unit).read()
        return v, t, {"shared_vertices": 0, "slabs": 1, "hypercube": True}
    if not gather:
        return part[0], part[1], {"tags": part[2], "slabs": w}
    def mark(name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("read_slab_mesh")
    vs, ts, gs = _gather_rows(part[0], group), _gather_rows(part[1], group), _gather_rows(part[2], group)
    mark("all_gather")
    v, t, stats = merge_slab_meshes(list(zip(vs, ts, gs)))
    mark("merge")
    stats["slab_vertices"] = [int(x.shape[0]) for x in vs]
    stats["near_plane"] = sum_int(meshes[0].near_plane)
    if marks is not None:   # this rank's device time from mark to mark (host waits inside a phase count towards it)
        torch.cuda.synchronize()
        stats["phase_ms"] = {b[0]: round(a[1].elapsed_time(b[1]), 4) for a, b in zip(marks[:-1], marks[1:])}
    return v, t, stats


# ==================================================================================================
# Plane sharding of the skeleton sweep (exact): the dense part of the path over the GPUs, the rest replicated
# ==================================================================================================
# The sweep over the marks grid (|sdf| and |grad| at every grid vertex: the one phase whose cost is the grid's
# VOLUME, a fifth of a large extraction) is split by planes of the first axis.  Every rank evaluates its planes,
# one all-gather moves the |sdf| planes and one MAX all-reduce the per-chunk gradient maxima (the reference's
# threshold is per chunk, tropical.py:189-197); from there on every rank holds the sweep of the whole grid and
# runs the subdivision and the faces itself, so the mesh is the single-GPU mesh bit for bit, on every rank,
# whatever the cut.  (Slab sharding of the whole path, above, cuts the subdivision too, but is exact only when no
# vertex falls within eps of a shared plane.)
def plane_ranges(n_marks: int, world_size: int) -> List[tuple]:
    """[(x_lo, x_hi)] per rank: disjoint, contiguous, near-equal plane ranges of the first grid axis."""
    if world_size < 1 or world_size > n_marks:
        raise ValueError(f"cannot deal {n_marks} planes to {world_size} ranks")
    bounds = [(r * n_marks) // world_size for r in range(world_size + 1)]
    return [(bounds[r], bounds[r + 1] - 1) for r in range(world_size)]


def subpoly_sweep_sharded(net, size: float = 1.2, eps: float = 1e-4, unit: int = 128, force: bool = True, group=None):
    """One object, the skeleton sweep sharded over the ranks of `group` by grid planes.  Every rank calls this with
    the same network and gets the whole mesh (a NativeMesh), identical to net.subpoly(...)."""
    if world() == 1:
        return net.subpoly(size=size, eps=eps, force=force, unit=unit)
    w, me = dist.get_world_size(group), dist.get_rank(group)
    M = net.n_marks
    ranges = plane_ranges(M, w)
    lo, hi = ranges[me]
    sw = net.skeleton_sweep(lo, hi, False, False, unit)
    per = max(b - a + 1 for a, b in ranges) * M * M          # equal-sized all-gather slots (NCCL all_gather_into_tensor)
    full = torch.empty(w * per, dtype=torch.float32, device="cuda")
    sw.read_dist(full[me * per:(me + 1) * per])
    dist.all_gather_into_tensor(full, full[me * per:(me + 1) * per].clone(), group=group)
    mg = sw.max_grad()
    dist.all_reduce(mg, op=dist.ReduceOp.MAX, group=group)
    whole = net.skeleton_sweep_alloc(unit)
    for r, (a, b) in enumerate(ranges):
        whole.write_dist(full[r * per:], a, b)
    whole.set_max_grad(mg)
    c = whole.finish()
    if c.num_edges == 0:   # empty skeleton: the hypercube route (subpoly.py:51-52)
        return net.subpoly(size=size, eps=eps, force=force, unit=unit)
    c.steps(_hyperplanes(net), eps, force)
    return c.extract_mesh(eps)
