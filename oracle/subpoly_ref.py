"""oracle/subpoly_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's polyhedral-complex mesh extraction -- the planar
(`force=True`, the reference default) path and the curve-approximation (`force=False`) path --
on top of the C network evaluation in `trinet_ref.c`.  Only tests/, `__graft_entry__.smoke()` and
bench.py's cpu_baseline / `--impl reference` legs may import this module.

Reference functions restated (file:line under /root/reference/tropical):
  skeleton()            tropical.py:158-225  (distance pruning mode, :188-197, :113-138)
  subpoly()             subpoly.py:23-86
  subpoly_()            subpoly.py:90-279    (both branches)
  deal_with_gradient_descent   subpoly_debug.py:121-165 (C: curve_gradient_descent; pinned by
                        tests/golden/gd_stage.npz, the reference function's own inputs and results)
  strict_check          subpoly_debug.py:234-271
  corner_points / intersection_of_two_planes   geometry.py:350-372, :24-138 (C: curve_intersection)
  check_edges_with_new_vertices (failover)   subpoly_debug.py:33-51
  regions_to_vertices   subpoly.py:281-340
  r_idx_as_tensor       subpoly.py:342-370
  edge_vertices         subpoly.py:484-535
  extract_skeleton      subpoly.py:556-581
  extract_faces         subpoly.py:584-652
  mean_points_with_valid subpoly.py:669-678
  tensor_to_triangle_faces subpoly.py:700-728
  sort_polygon_vertices_batch geometry.py:483-525
  extract_triangles_from_sorted_vertices_and_mask geometry.py:536-556
  get_hypercube         subpoly.py:731-750

Pinned by tests/test_oracle_vs_reference.py (runs the reference itself in this
container) and by the committed fixtures in tests/golden/.

Float contract: every float32 expression below is a chain of single IEEE
operations in the written order (numpy never fuses); reductions whose order
torch leaves unspecified (sums over a padded face row, 3-vector dot products) are
DEFINED here as left-to-right sums, and the CUDA path follows this file.
"""
import numpy as np


class GradientDescentFailed(RuntimeError):
    """The reference ends the extraction here (subpoly.py:172-174: the check after the repair)."""

F32 = np.float32


# --------------------------------------------------------------------------------------
# skeleton
# --------------------------------------------------------------------------------------
def skeleton(P, unit: int = 128):
    """tropical.py:158-225.  Returns (vertices [V,3] f32, edges [E,2] i64); empty
    arrays when no grid edge survives the distance pruning."""
    marks = P.marks
    L = len(marks)
    len_max = F32(np.diff(marks).max())
    k3 = F32(np.sqrt(F32(3.0))) * F32(2)
    chunks = []
    for i in range(0, L, unit - 1):
        for j in range(0, L, unit - 1):
            for k in range(0, L, unit - 1):
                start = [i, j, k]
                end = [min(L, s + unit) for s in start]
                ax = [np.arange(s, e) for s, e in zip(start, end)]
                gi, gj, gk = np.meshgrid(*ax, indexing="ij")
                shape = gi.shape
                xp = np.stack([marks[gi.ravel()], marks[gj.ravel()], marks[gk.ravel()]], -1)
                x = P.preprocess_inverse(xp)
                sdf, grad = P.sdf_grad(x)
                max_grad = P.grad_norm(grad).max()
                dist = np.abs(sdf).reshape(shape)
                eps = (k3 * len_max) * F32(max_grad)
                vid = (gi * L * L + gj * L + gk).astype(np.int64)
                ok = dist <= eps
                m = ok[1:, :, :] & ok[:-1, :, :]
                chunks.append(np.stack([vid[1:, :, :][m], vid[:-1, :, :][m]], -1))
                m = ok[:, 1:, :] & ok[:, :-1, :]
                chunks.append(np.stack([vid[:, 1:, :][m], vid[:, :-1, :][m]], -1))
                m = ok[:, :, 1:] & ok[:, :, :-1]
                chunks.append(np.stack([vid[:, :, 1:][m], vid[:, :, :-1][m]], -1))
    edges = np.concatenate(chunks, 0) if chunks else np.zeros((0, 2), np.int64)
    if edges.shape[0] == 0:
        return np.zeros((0, 3), F32), np.zeros((0, 2), np.int64)
    v_idx, inv = np.unique(edges.reshape(-1), return_inverse=True)
    edges = inv.reshape(-1, 2).astype(np.int64)
    p = np.stack([v_idx // (L * L), (v_idx // L) % L, v_idx % L], -1)
    vertices = P.preprocess_inverse(marks[p])
    return vertices.astype(F32), edges


def get_hypercube(size):
    """subpoly.py:731-750: the 8 corners / 12 edges of (-size, size)^3."""
    c = np.array([-size, size], F32)
    g = np.stack(np.meshgrid(c, c, c, indexing="ij"), -1).reshape(-1, 3)
    edges = [[a, b] for a in range(8) for b in range(a + 1, 8)
             if int((g[a] * g[b] < 0).sum()) == 1]
    return g.astype(F32), np.array(edges, np.int64)


# --------------------------------------------------------------------------------------
# regions -> vertex groups
# --------------------------------------------------------------------------------------
def regions_to_vertices(m, offset):
    """subpoly.py:281-340.  m: [V, D+K] int (first D columns are grid masks 0/1, rest
    signs), offset: [V, D].  Every zero entry is expanded to both sides; returns
    (region id per expanded row, original row per expanded row)."""
    m = np.asarray(m, np.int64)
    offset = np.asarray(offset, np.int64)
    if m.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    D = offset.shape[1]
    zeros = (m == 0).sum(1)
    aug, org = [], []
    for k in np.unique(zeros):
        rows = np.nonzero(zeros == k)[0]
        k = int(k)
        n = 1 << k
        # sign patterns in cartesian-product order: first zero column is the slowest digit
        pat = np.array([[1 if (q >> (k - 1 - b)) & 1 else -1 for b in range(k)]
                        for q in range(n)], np.int64).reshape(n, k)
        rep = np.repeat(m[rows], n, axis=0)
        if k:
            rep[rep == 0] = np.tile(pat, (rows.size, 1)).reshape(-1)
        rep[:, :D] = (rep[:, :D] - 1) // 2 + np.repeat(offset[rows], n, axis=0)
        aug.append(rep)
        org.append(np.repeat(rows, n))
    aug = np.concatenate(aug, 0)
    org = np.concatenate(org, 0)
    _, r_idx = np.unique(aug, axis=0, return_inverse=True)
    return r_idx.reshape(-1).astype(np.int64), org.astype(np.int64)


def r_idx_as_tensor(r_idx, idx_org, null_value=-1):
    """subpoly.py:342-370: one left-aligned row of vertex indices per region
    (stable order inside a region, as torch's CPU argsort gives)."""
    order = np.argsort(r_idx, kind="stable")
    r_sorted, v_sorted = r_idx[order], idx_org[order]
    _, first, counts = np.unique(r_sorted, return_index=True, return_counts=True)
    out = np.full((counts.size, int(counts.max())), null_value, np.int64)
    col = np.arange(r_sorted.size) - np.repeat(first, counts)
    out[np.repeat(np.arange(counts.size), counts), col] = v_sorted
    return out


def edge_vertices(m, offset):
    """subpoly.py:484-535: pairs of candidate vertices that share a (expanded) region
    and at least one plane.  Returns [P,2] indices into the candidate list."""
    r_idx, org = regions_to_vertices(m, offset)
    v_indices = r_idx_as_tensor(r_idx, org)
    C = v_indices.shape[1]
    out = []
    for i in range(1, C):
        a = v_indices[:, i]
        ok = a != -1
        a = np.tile(a[ok], i)
        b = v_indices[ok, :i].T.reshape(-1)
        out.append(np.stack([a, b], 1))
    if not out:
        return np.zeros((0, 2), np.int64)
    pairs = np.unique(np.concatenate(out, 0), axis=0)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    D = offset.shape[1]
    z = (np.asarray(m) == 0)[pairs]                       # [P,2,cols]
    both = z[:, 0] & z[:, 1]
    zero_counts = both.sum(-1)
    off = np.asarray(offset, np.int64)[pairs]
    zero_counts = zero_counts - (both[:, :D] & (off[:, 0] != off[:, 1])).sum(-1)
    return pairs[zero_counts >= 1]


# --------------------------------------------------------------------------------------
# one hyperplane
# --------------------------------------------------------------------------------------
def corner_points(e):
    """geometry.py:350-372: the 8 corners of the box an edge spans, index 4*iz + 2*iy + ix
    (coordinate taken from endpoint 0 or 1 per axis).  e: [E,2,3] -> [E,8,3]."""
    out = np.empty((e.shape[0], 8, 3), F32)
    for i in range(2):
        for j in range(2):
            for k in range(2):
                out[:, 4 * i + 2 * j + k, 0] = e[:, k, 0]
                out[:, 4 * i + 2 * j + k, 1] = e[:, j, 1]
                out[:, 4 * i + 2 * j + k, 2] = e[:, i, 2]
    return out


def subpoly_step(P, vertices, edges, outputs, l, h, eps, pruning=True, force=True):
    """subpoly.py:90-279.  Returns (vertices, edges, outputs)."""
    eps32 = F32(eps)
    H = P.num_hidden
    idx = l * H + h
    o = outputs[:, idx]
    d = o[edges]                                           # [E,2]
    m = (d[:, 0] * d[:, 1]) < 0
    m &= (np.abs(d[:, 0]) > eps32) & (np.abs(d[:, 1]) > eps32)
    if m.sum() == 0:
        return vertices, edges, outputs
    d = d[m] / eps32
    e = vertices[edges][m]                                 # [S,2,3]
    w = np.abs(d[:, :1]) / np.abs(d[:, 1:] - d[:, :1])
    regions, offs, _ = P.region(vertices, outputs)
    v_new = e[:, 0] * (F32(1) - w) + e[:, 1] * w
    have_c = False
    if not force:
        # bi-/tri-linear corrections for edges that are not axis aligned (subpoly.py:120-183)
        from .trinet import curve_intersections
        c = ((np.abs(e[:, 1, :] - e[:, 0, :]) > eps32).sum(-1)) > 1
        ec = e[c]
        have_c = ec.shape[0] > 0
        if have_c:
            dd = P.outputs_group8(corner_points(ec))           # [Ec,8,R]
            rg = regions[edges][m][c][:, :, 3:]
            r_edges = (rg[:, 0] == 0) & (rg[:, 1] == 0)
            sub = r_edges[:, :idx]
            if not sub.any(1).all():
                raise RuntimeError("an edge lies on no earlier plane (the reference exits here, subpoly.py:140-148)")
            plane = idx - 1 - np.argmax(sub[:, ::-1], 1)       # nonzero_last
            rows = np.arange(ec.shape[0])
            ints = curve_intersections(dd[rows, :, plane], dd[:, :, idx])
            xg = ec[:, 0] * (F32(1) - ints) + ec[:, 1] * ints
            og = P.outputs(xg)
            d_new = np.stack([og[rows, plane], og[:, idx]], -1)
            gg = ((ints < 0) | (ints > 1)).sum(-1) > 0
            gd = ~gg & ((np.abs(d_new) > eps32).sum(-1) > 0)
            if gd.any():
                # subpoly_debug.py:121-165; the reference then ends the extraction if a repaired
                # intersection is still off its planes (subpoly.py:172-174 -> exit())
                from .trinet import gradient_descent
                ints = ints.copy()
                ints[gd], d_new[gd], _ = gradient_descent(P, ec[gd, 0], ec[gd, 1], ints[gd], plane[gd], idx, eps32)
                if (np.abs(d_new[~gg]) > eps32).any():
                    raise GradientDescentFailed(f"hyperplane {l}/{h}: {int(gd.sum())} intersection(s) still off their planes "
                                                "after the gradient-descent repair (the reference ends here, subpoly.py:172-174)")
            v_new[c] = ec[:, 0] + ints * (ec[:, 1] - ec[:, 0])
    m_rgn_all, offset, outputs_new = P.region(v_new)
    m_idx = 3 + idx

    # failover of subpoly_debug.check_edges_with_new_vertices (:33-51)
    er, eo = regions[edges][m], offs[edges][m]
    chk = (er[:, 0] == 0) & (er[:, 1] == 0)
    chk[:, :3] &= eo[:, 0] == eo[:, 1]
    b = chk[:, 3:].copy()
    b[:, idx:] = False
    b[:, idx] = True
    if (np.abs(outputs_new[b]) > eps32).any():
        outputs_new = outputs_new.copy()
        outputs_new[b] = 0
        m_rgn_all, offset, outputs_new = P.region(v_new, outputs_new)
    if not force:
        # strict_check (subpoly_debug.py:234-271): drop new vertices that are not on the plane,
        # and every non-axis-aligned edge without an admissible intersection stays unsplit
        chk = outputs_new[:, idx]
        if have_c or np.abs(chk).max() >= eps32:
            g = np.abs(chk) < eps32
            if have_c:
                g[c] = (np.abs(chk[c]) < eps32) & ~gg
            m = m.copy()
            m[m] = g
            v_new, m_rgn_all, offset, outputs_new = v_new[g], m_rgn_all[g], offset[g], outputs_new[g]
    m_rgn, m_rgn_f = m_rgn_all[:, :m_idx], m_rgn_all[:, m_idx:]

    V0 = vertices.shape[0]
    edges = edges.copy()
    right = edges[m, 1].copy()
    new_ids = np.arange(v_new.shape[0], dtype=np.int64) + V0
    edges[m, 1] = new_ids
    e_new = np.stack([right, new_ids], -1)

    # connecting edges among the new vertices and the old ones the plane hits (:231-244)
    hit = np.abs(outputs[:, idx]) < eps32
    v_rgn = np.concatenate([m_rgn, regions[hit, :m_idx]], 0)
    v_off = np.concatenate([offset, offs[hit]], 0)
    cand = np.concatenate([new_ids, np.nonzero(hit)[0].astype(np.int64)], 0)
    c_new = cand[edge_vertices(v_rgn, v_off)]
    if c_new.shape[0]:
        c_new = np.unique(np.sort(c_new, -1), axis=0)

    vertices_old = vertices
    vertices = np.concatenate([vertices, v_new.astype(F32)], 0)
    edges = np.concatenate([edges, e_new, c_new.reshape(-1, 2)], 0)
    outputs = np.concatenate([outputs, outputs_new], 0)

    if h < H and pruning:
        fut_old = P.region(vertices_old, outputs[:V0])[0][:, m_idx:]
        fut = np.concatenate([fut_old, m_rgn_f], 0)
        _, r = np.unique(fut, axis=0, return_inverse=True)
        r = r.reshape(-1)
        edges = edges[r[edges[:, 0]] != r[edges[:, 1]]]
        v_idx, inv = np.unique(edges.reshape(-1), return_inverse=True)
        vertices = vertices[v_idx]
        edges = inv.reshape(-1, 2).astype(np.int64)
        outputs = outputs[v_idx]
    return vertices, edges, outputs


# --------------------------------------------------------------------------------------
# surface skeleton and faces
# --------------------------------------------------------------------------------------
def extract_skeleton(P, vertices, edges, outputs, eps):
    """subpoly.py:556-581."""
    eps32 = F32(eps)
    m = np.abs(outputs[:, -1]) < eps32
    v = P.preprocess(vertices)
    m[(v > 1).sum(-1) > 0] = False
    m[(v < 0).sum(-1) > 0] = False
    if m.sum() < 3:
        return np.zeros((0, 3), F32), np.zeros((0, 2), np.int64), None
    edges = edges[m[edges].sum(-1) == 2]
    v_idx, inv = np.unique(edges.reshape(-1), return_inverse=True)
    return vertices[v_idx], inv.reshape(-1, 2).astype(np.int64), v_idx


def _seq_sum(a, axis):
    """left-to-right float32 sum along `axis` (defines the order torch leaves open)."""
    a = np.moveaxis(a, axis, 0)
    acc = np.zeros(a.shape[1:], F32)
    for k in range(a.shape[0]):
        acc = acc + a[k]
    return acc


def mean_points_with_valid(vertices, v_indices, null_value=-1):
    """subpoly.py:669-678."""
    pad = v_indices == null_value
    points = vertices[v_indices + pad].copy()
    points[pad] = 0
    Z = (~pad).sum(1, keepdims=True)
    mean = _seq_sum(points, 1) / Z.astype(F32)
    ok = Z[:, 0] >= 3
    return mean[ok], points[ok], v_indices[ok]


def polygon_order(points, normals):
    """geometry.py:483-516: per-face permutation that sorts the (padded) face
    vertices by angle around the face centre, seen against `normals`."""
    v = points
    n2 = (v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2]
    valid = np.sqrt(n2) > 0                               # [B,M]
    k = valid.sum(1).astype(F32)
    k[k == 0] = 1
    u = v - (_seq_sum(v, 1) / k[:, None])[:, None, :]
    a = u[:, :1]
    d = np.stack([a[..., 1] * u[..., 2] - a[..., 2] * u[..., 1],
                  a[..., 2] * u[..., 0] - a[..., 0] * u[..., 2],
                  a[..., 0] * u[..., 1] - a[..., 1] * u[..., 0]], -1)

    def unit(x):
        nrm = np.sqrt((x[..., 0] * x[..., 0] + x[..., 1] * x[..., 1]) + x[..., 2] * x[..., 2])
        return x / np.maximum(nrm, F32(1e-8))[..., None]

    ua, uu = unit(a), unit(u)
    c = (ua[..., 0] * uu[..., 0] + ua[..., 1] * uu[..., 1]) + ua[..., 2] * uu[..., 2]
    nn = normals[:, None, :]
    dn = (d[..., 0] * nn[..., 0] + d[..., 1] * nn[..., 1]) + d[..., 2] * nn[..., 2]
    s = c * ((dn >= 0).astype(F32) * F32(2) - F32(1)) + (dn < 0).astype(F32) * F32(2)
    order = np.argsort(-s, axis=1, kind="stable")
    return order, valid


def fan_triangles(rows, mask):
    """The fan triangulation both geometry.py:536-556 and subpoly.py:700-728 perform on
    left-to-right masked rows: for step i, rows with >= i+3 entries emit
    (first, i+1-th, i+2-th).  Output is ordered by step, then by row."""
    counts = mask.sum(1)
    flat = rows[mask]
    start = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    out = []
    for i in range(int(counts.max()) - 2 if counts.size else 0):
        ok = counts >= i + 3
        s = start[ok]
        out.append(np.stack([flat[s], flat[s + i + 1], flat[s + i + 2]], 1))
    if not out:
        return flat[:0].reshape(0, 3, *flat.shape[1:])
    return np.concatenate(out, 0)


def extract_faces(P, vertices, outputs, eps):
    """subpoly.py:584-652.  Returns (faces [T,3,3] f32 positions, faces_with_indices
    [T',3] i64, polygons [B,M] i64 = the angle-sorted face rows, -1 where blank)."""
    none = (np.zeros((0, 3, 3), F32), np.zeros((0, 3), np.int64), np.zeros((0, 0), np.int64))
    if vertices.shape[0] == 0:
        return none
    m_rgn, offset, _ = P.region(vertices, outputs, eps)
    r_idx, org = regions_to_vertices(m_rgn[:, :-1], offset)
    v_indices = np.unique(r_idx_as_tensor(r_idx, org), axis=0)
    mean, points, v_indices = mean_points_with_valid(vertices, v_indices)
    if mean.shape[0] == 0:
        return none
    _, normals = P.sdf_grad(mean)
    order, valid = polygon_order(points, normals)
    p_sorted = np.take_along_axis(points, order[..., None], 1)
    valid_sorted = np.take_along_axis(valid, order, 1)
    faces = fan_triangles(p_sorted, valid_sorted)
    idx_sorted = np.take_along_axis(v_indices, order, 1)
    # subpoly.py:702-704: blank repeated indices inside a row
    for i in range(idx_sorted.shape[1]):
        dup = (idx_sorted[:, :i] == idx_sorted[:, i:i + 1]).sum(-1) > 0
        idx_sorted[dup, i] = -1
    tri = fan_triangles(idx_sorted, idx_sorted != -1)
    return faces.astype(F32), tri.astype(np.int64), idx_sorted


def canonical_polygons(polygons):
    """Order-free view of a face list: each polygon as the cyclic vertex sequence
    rotated to start at its smallest index; the list sorted.  The reference's in-row
    order comes from a non-stable torch.argsort (subpoly.py:357), so its fan apex and
    face order are implementation-defined; this view is what must agree with it."""
    out = []
    for row in np.asarray(polygons):
        r = [int(v) for v in row if v != -1]
        if len(r) < 3:
            continue
        k = r.index(min(r))
        out.append(tuple(r[k:] + r[:k]))
    return sorted(out)


# --------------------------------------------------------------------------------------
# driver
# --------------------------------------------------------------------------------------
def subpoly(P, size=1.2, eps=1e-4, unit=128, return_intermediate=False, force=True):
    """subpoly.py:23-86.  Returns (faces, vertices, faces_with_indices);
    with return_intermediate also a dict of the pre-extraction complex."""
    vertices, edges = skeleton(P, unit)
    if edges.shape[0] == 0:
        vertices, edges = get_hypercube(size)
    outputs = P.outputs(vertices)
    H = P.num_hidden
    for l in range(P.num_layers - 1):
        for h in range(H):
            vertices, edges, outputs = subpoly_step(P, vertices, edges, outputs, l, h, eps, force=force)
    vertices, edges, outputs = subpoly_step(P, vertices, edges, outputs,
                                            P.num_layers - 2, H, eps, force=force)
    inter = dict(vertices=vertices, edges=edges, outputs=outputs)
    s_vertices, s_edges, v_idx = extract_skeleton(P, vertices, edges, outputs, eps)
    if v_idx is None:
        res = (np.zeros((0, 3, 3), F32), s_vertices, np.zeros((0, 3), np.int64))
        return res + (inter,) if return_intermediate else res
    s_outputs = outputs[v_idx]
    inter.update(surface_edges=s_edges, surface_outputs=s_outputs)
    faces, tri, polygons = extract_faces(P, s_vertices, s_outputs, eps)
    inter.update(polygons=polygons)
    res = (faces, s_vertices, tri)
    return res + (inter,) if return_intermediate else res
