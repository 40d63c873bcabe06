"""Oracle (numpy): mesh -> graph, kd decomposition with halo, subdomain build, stitch.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Integer outputs of these functions are
what the CUDA assembly / stitch kernels must reproduce bit for bit.

Reference anchors (all under /root/reference):
  * edges + edge_attr ...... dataset/GraphDataset.py:838-869  (AnsysDataset.vtk_to_pyg)
  * per-subdomain contents . dataset/GraphDataset.py:1245-1284 (x, y, pos, edge_index,
                             edge_attr, global_node_ids written per partition)
  * decomposition .......... dataset/GraphDataset.py:1183-1230 (vtkRedistributeDataSetFilter,
                             AssignToAllIntersectingRegions at :1219; Duct uses
                             AssignToOneRegion at :565)
  * stitch ................. dataset/GraphDataset.py:1370-1400 (mean over coincident points)
"""
from __future__ import annotations

import numpy as np

MODE_ONE_REGION = 0          # reference Duct path, GraphDataset.py:565
MODE_ALL_INTERSECTING = 1    # reference Ansys path, GraphDataset.py:1219


# --------------------------------------------------------------------------------------
# a1: edges of a tetrahedral (or any simplex/cell) mesh
# --------------------------------------------------------------------------------------
def build_edges(cells: np.ndarray, pos: np.ndarray):
    """All ordered point pairs of every cell, both directions, de-duplicated.

    Follows GraphDataset.py:850-864 (the reference inserts (j,k) and (k,j) for every
    j<k of every cell into a Python set).  The set's iteration order is arbitrary, so the
    contract is the edge SET; we return it in the canonical order (dst, src) ascending --
    the CSR-by-destination order every kernel uses.

    edge_attr follows GraphDataset.py:866-867: fp32 ``np.linalg.norm(pos[src]-pos[dst])``.
    Returns (src[E] int64, dst[E] int64, edge_attr[E] float32).
    """
    cells = np.asarray(cells, dtype=np.int64)
    n = int(pos.shape[0])
    k = cells.shape[1]
    keys = []
    for a in range(k):
        for b in range(k):
            if a != b:
                keys.append(cells[:, b] * n + cells[:, a])      # key = dst * n + src
    if not keys or cells.shape[0] == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z.copy(), np.zeros(0, dtype=np.float32)
    key = np.unique(np.concatenate(keys))
    dst = key // n
    src = key % n
    keep = src != dst            # degenerate cells with a repeated vertex: the reference
    src, dst = src[keep], dst[keep]  # would add a self loop; our meshes never have them
    pos = np.asarray(pos, dtype=np.float32)
    attr = np.linalg.norm(pos[src] - pos[dst], axis=1).astype(np.float32)
    return src, dst, attr


def csr_by_destination(src: np.ndarray, dst: np.ndarray, n: int):
    """rowptr[n+1], perm[E]: edges sorted by (dst, src, original position)."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    perm = np.lexsort((np.arange(src.size), src, dst))
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, dst + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, perm


# --------------------------------------------------------------------------------------
# a2: kd decomposition
# --------------------------------------------------------------------------------------
def cell_centroids(pos: np.ndarray, cells: np.ndarray) -> np.ndarray:
    """fp32 centroid with a fixed association: ((p0+p1)+(p2+p3))*0.25."""
    p = np.asarray(pos, dtype=np.float32)[np.asarray(cells, dtype=np.int64)]
    return ((p[:, 0] + p[:, 1]) + (p[:, 2] + p[:, 3])) * np.float32(0.25)


def cell_aabb(pos: np.ndarray, cells: np.ndarray):
    p = np.asarray(pos, dtype=np.float32)[np.asarray(cells, dtype=np.int64)]
    return p.min(axis=1), p.max(axis=1)


def kd_build(pos: np.ndarray, cells: np.ndarray, levels: int):
    """Exact-median kd bisection of the cell centroids into 2**levels leaves.

    Stands in for vtkRedistributeDataSetFilter's cut generation (GraphDataset.py:1208-1230;
    VTK C++ is not available, so this definition IS the oracle -- parity unpinned).
    Per tree node (heap order, children 2i+1 / 2i+2):
      axis  = longest extent of the region's centroid bounding box (first max),
      order = cells sorted by (centroid[axis], cell id),
      left  = first m//2 cells, right = the rest, split = centroid[axis] of right[0].
    An empty region has axis 0 and split +inf.
    Returns (home_leaf[C] int32, tree_axis[2^k-1] int32, tree_split[2^k-1] float32).
    """
    C = int(cells.shape[0])
    cent = cell_centroids(pos, cells)
    region = np.zeros(C, dtype=np.int64)
    n_internal = (1 << levels) - 1
    tree_axis = np.zeros(n_internal, dtype=np.int32)
    tree_split = np.full(n_internal, np.inf, dtype=np.float32)
    for d in range(levels):
        order_all = np.argsort(region, kind="stable")
        counts = np.bincount(region, minlength=1 << d)
        starts = np.concatenate([[0], np.cumsum(counts)])
        new_region = np.empty_like(region)
        for r in range(1 << d):
            ids = order_all[starts[r]:starts[r + 1]]          # ascending cell id
            node = (1 << d) - 1 + r
            m = ids.size
            if m == 0:
                continue
            c = cent[ids]
            ext = c.max(axis=0) - c.min(axis=0)
            axis = 0 if (ext[0] >= ext[1] and ext[0] >= ext[2]) else (1 if ext[1] >= ext[2] else 2)
            o = np.argsort(c[:, axis], kind="stable")
            half = m // 2
            tree_axis[node] = axis
            tree_split[node] = c[o[half], axis]
            new_region[ids[o[:half]]] = 2 * r
            new_region[ids[o[half:]]] = 2 * r + 1
        region = new_region
    return region.astype(np.int32), tree_axis, tree_split


def kd_assign(pos, cells, levels, mode, home_leaf, tree_axis, tree_split):
    """(leaf_ptr[S+1], leaf_cells[sum]) -- cells of every leaf, ascending cell id.

    ONE_REGION: the leaf of the centroid (home leaf).
    ALL_INTERSECTING: every leaf whose half-space box the cell's AABB touches; at a tree
    node the cell descends left if aabb_min[axis] < split, right if aabb_max[axis] >= split,
    and always along its own home path.
    """
    S = 1 << levels
    C = int(cells.shape[0])
    if mode == MODE_ONE_REGION or levels == 0:
        order = np.argsort(home_leaf, kind="stable").astype(np.int64)
        counts = np.bincount(home_leaf, minlength=S)
        return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64), order
    lo, hi = cell_aabb(pos, cells)
    cur_cell = np.arange(C, dtype=np.int64)
    cur_node = np.zeros(C, dtype=np.int64)     # within-level index
    for d in range(levels):
        heap = (1 << d) - 1 + cur_node
        ax = tree_axis[heap]
        sp = tree_split[heap]
        home_here = (home_leaf[cur_cell].astype(np.int64) >> (levels - d)) == cur_node
        home_right = ((home_leaf[cur_cell].astype(np.int64) >> (levels - d - 1)) & 1) == 1
        go_left = (lo[cur_cell, ax] < sp) | (home_here & ~home_right)
        go_right = (hi[cur_cell, ax] >= sp) | (home_here & home_right)
        cur_cell = np.concatenate([cur_cell[go_left], cur_cell[go_right]])
        cur_node = np.concatenate([2 * cur_node[go_left], 2 * cur_node[go_right] + 1])
    order = np.lexsort((cur_cell, cur_node))
    leaf_cells = cur_cell[order]
    counts = np.bincount(cur_node, minlength=S)
    leaf_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return leaf_ptr, leaf_cells


def kd_partition(pos, cells, levels, mode=MODE_ALL_INTERSECTING):
    home, ax, sp = kd_build(pos, cells, levels)
    leaf_ptr, leaf_cells = kd_assign(pos, cells, levels, mode, home, ax, sp)
    return {"home_leaf": home, "tree_axis": ax, "tree_split": sp,
            "leaf_ptr": leaf_ptr, "leaf_cells": leaf_cells}


# --------------------------------------------------------------------------------------
# a2/a3: per-subdomain graphs, flattened as one block-diagonal batch
# --------------------------------------------------------------------------------------
def build_subdomains(pos, cells, leaf_ptr, leaf_cells):
    """Per-subdomain node compaction + edge build, concatenated block-diagonally.

    For subdomain s (GraphDataset.py:1245-1284): nodes = unique vertices of its cells in
    ascending global id (-> ``global_ids``), local id = rank; edges = build_edges on the
    subdomain's cells in local ids, canonical (dst, src) order; edge_attr fp32 length.
    Batch-level indices (local + node_ptr[s]) are what PyG's Batch would produce
    (reference models/scheduler_gnn.py:127,376).

    Returns dict: node_ptr[S+1], global_ids[sum n], edge_ptr[S+1], edge_src[sum E],
    edge_dst[sum E] (batch-level, int64), edge_attr[sum E] f32, rowptr[sum n + 1].
    """
    S = leaf_ptr.size - 1
    N = int(pos.shape[0])
    cells = np.asarray(cells, dtype=np.int64)
    node_ptr = np.zeros(S + 1, dtype=np.int64)
    edge_ptr = np.zeros(S + 1, dtype=np.int64)
    gids, srcs, dsts, attrs = [], [], [], []
    lut = np.full(N, -1, dtype=np.int64)
    for s in range(S):
        cs = cells[leaf_cells[leaf_ptr[s]:leaf_ptr[s + 1]]]
        g = np.unique(cs)
        lut[g] = np.arange(g.size)
        lc = lut[cs]
        src, dst, attr = build_edges(lc, pos[g])
        gids.append(g)
        srcs.append(src + node_ptr[s])
        dsts.append(dst + node_ptr[s])
        attrs.append(attr)
        node_ptr[s + 1] = node_ptr[s] + g.size
        edge_ptr[s + 1] = edge_ptr[s] + src.size
    cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt))
    edge_dst = cat(dsts, np.int64)
    n_tot = int(node_ptr[-1])
    rowptr = np.zeros(n_tot + 1, dtype=np.int64)
    np.add.at(rowptr, edge_dst + 1, 1)
    return {"node_ptr": node_ptr, "global_ids": cat(gids, np.int64), "edge_ptr": edge_ptr,
            "edge_src": cat(srcs, np.int64), "edge_dst": edge_dst,
            "edge_attr": cat(attrs, np.float32), "rowptr": np.cumsum(rowptr)}


# --------------------------------------------------------------------------------------
# a10: overlap stitch
# --------------------------------------------------------------------------------------
def occurrence_csr(global_ids: np.ndarray, N: int):
    """occ_ptr[N+1], occ_idx[sum n]: positions in the concatenated batch of every global
    node, ascending (so ascending (subdomain, local) order)."""
    global_ids = np.asarray(global_ids, dtype=np.int64)
    occ_idx = np.argsort(global_ids, kind="stable").astype(np.int64)
    occ_ptr = np.concatenate([[0], np.cumsum(np.bincount(global_ids, minlength=N))]).astype(np.int64)
    return occ_ptr, occ_idx


def _numpy_pairwise_order_sum(cols):
    """Sum of m fp32 values per row in the order numpy's pairwise summation uses for a contiguous 1-D reduce
    (numpy/core/src/umath/loops_utils.h.src, pairwise_sum): m < 8 sequential; 8 <= m <= 128 eight running partial
    sums r[j] += a[8 k + j], combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the m % 8 tail sequentially.
    cols: list of m float32 arrays (one per copy, same length)."""
    m = len(cols)
    if m < 8:
        acc = np.zeros_like(cols[0])
        for c in cols:
            acc = acc + c
        return acc
    assert m <= 128, "a mesh node shared by more than 128 subdomains"
    r = [cols[j].copy() for j in range(8)]
    i = 8
    while i < m - (m % 8):
        for j in range(8):
            r[j] = r[j] + cols[i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < m:
        res = res + cols[i]
        i += 1
    return res


def stitch_mean(values: np.ndarray, global_ids: np.ndarray, N: int):
    """Mean over all subdomain copies of each global node (GraphDataset.py:1383-1400:
    ``np.mean(sub_vals, axis=0)`` over the coincident points, written back to every copy).

    PINNED against the reference's own loop (tests/golden/make_golden.py `make_stitch_golden` executes lines
    1371-1400 of the reference; tests/test_oracle_golden.py compares bit for bit).  The reference averages its point
    arrays one by one -- `velocity` [m, 3] and `pressure` [m] -- with fp32 np.mean, whose summation ORDER depends on
    the array's rank: the [m, 3] vector arrays are reduced row by row (sequential in the copies, ascending
    concatenation order), the 1-D scalar arrays through numpy's pairwise summation (sequential below 8 copies, eight
    interleaved partial sums from 8 on).  values [sum n, C]: C = 4 is (velocity, pressure) -> channels 0..2
    sequential, channel 3 pairwise; C = 1 is a scalar array (pairwise); any other C a vector array (sequential).
    Division by the count in fp32.  Returns (field[N,C] f32, count[N] int32, merged[sum n, C] f32).
    """
    values = np.asarray(values, dtype=np.float32)
    C = values.shape[1]
    occ_ptr, occ_idx = occurrence_csr(global_ids, N)
    count = np.diff(occ_ptr).astype(np.int32)
    field = np.zeros((N, C), dtype=np.float32)
    scalar_ch = [3] if C == 4 else ([0] if C == 1 else [])
    for m in np.unique(count):
        if m == 0:
            continue
        sel = np.nonzero(count == m)[0]
        cols = [values[occ_idx[occ_ptr[sel] + j]] for j in range(int(m))]       # copy j of every node with m copies
        acc = np.zeros((sel.size, C), dtype=np.float32)
        for c in cols:
            acc = acc + c
        for ch in scalar_ch:
            acc[:, ch] = _numpy_pairwise_order_sum([c[:, ch] for c in cols])
        field[sel] = acc / np.float32(m)
    merged = field[np.asarray(global_ids, dtype=np.int64)]
    return field, count, merged


# --------------------------------------------------------------------------------------
# low-res -> high-res transfer (the step before the path)
# --------------------------------------------------------------------------------------
def interp_gaussian(src_pos, src_val, dst_pos, radius, sharpness=2.0, null_value=0.0):
    """AnsysDataset._lagrangian_interpolation (dataset/GraphDataset.py:1041-1105): vtkPointInterpolator with a
    vtkGaussianKernel(radius, sharpness) -- vtk==9.4.1, not vendored: restated from the published algorithm
    (vtkGaussianKernel::ComputeWeights: w = exp(-(sharpness / radius)^2 d^2), normalised; points gathered with
    FindPointsWithinRadius, d^2 <= radius^2; no point in range -> NULL_VALUE strategy, value 0).  PARITY UNPINNED
    (no VTK here).  fp64 arithmetic on the fp32 inputs; returns (values float32 [n_dst, c], count int32 [n_dst])."""
    from scipy.spatial import cKDTree
    sp = np.asarray(src_pos, dtype=np.float32).astype(np.float64)
    dp = np.asarray(dst_pos, dtype=np.float32).astype(np.float64)
    sv = np.asarray(src_val, dtype=np.float32).astype(np.float64).reshape(sp.shape[0], -1)
    r = float(np.float32(radius))
    r2 = r * r
    f2 = (float(np.float32(sharpness)) / r) ** 2
    out = np.full((dp.shape[0], sv.shape[1]), float(null_value), dtype=np.float64)
    cnt = np.zeros(dp.shape[0], dtype=np.int32)
    if sp.shape[0]:
        tree = cKDTree(sp)
        cand = tree.query_ball_point(dp, r * (1 + 1e-9) + 1e-300)      # superset; the exact test follows
        for t, js in enumerate(cand):
            if not js:
                continue
            js = np.asarray(sorted(js))
            e = sp[js] - dp[t]
            d2 = e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1] + e[:, 2] * e[:, 2]
            keep = d2 <= r2
            if not keep.any():
                continue
            w = np.exp(-f2 * d2[keep])
            out[t] = (w[:, None] * sv[js[keep]]).sum(0) / w.sum()
            cnt[t] = int(keep.sum())
    return out.astype(np.float32), cnt


# --------------------------------------------------------------------------------------
# wall shear stress (the step after the path)
# --------------------------------------------------------------------------------------
def wall_shear_stress(pos, cells, velocity, dynamic_viscosity=1.0):
    """compute_wss.py:5-120 restated for a tetrahedral mesh (vtk==9.4.1 filters, not vendored -> PARITY UNPINNED):
    vtkGradientFilter = mean over the incident tets of their constant gradient (:36-42), vtkDataSetSurfaceFilter =
    faces owned by one cell (:45-48), vtkPolyDataNormals = normalised sum of the unit face normals at a point
    (:53-58), then tau = mu (G + G^T) n, tau_wall = tau - (tau . n) n and its norm (:86-99).  Faces are oriented
    outward and sharp edges are not split (see fesr_b200/csrc/wss.cu).  fp64.
    -> dict(surface_nodes, faces (sorted rows), normals, gradient [N, 9], wss, wss_magnitude)."""
    p = np.asarray(pos, dtype=np.float32).astype(np.float64)
    c = np.asarray(cells).astype(np.int64)
    u = np.asarray(velocity, dtype=np.float32).astype(np.float64)
    N = p.shape[0]
    E = p[c[:, 1:]] - p[c[:, :1]]                       # [C, 3, 3] edge vectors (rows)
    dU = u[c[:, 1:]] - u[c[:, :1]]                      # [C, 3, 3]: dU[k, i] = u_i(v_k) - u_i(v_0)
    G = np.transpose(np.linalg.solve(E, dU), (0, 2, 1)) # solve E X = dU, X[j, i] = d u_i / d x_j -> G[i, j]
    gsum = np.zeros((N, 9))
    cnt = np.zeros(N)
    for k in range(4):
        np.add.at(gsum, c[:, k], G.reshape(-1, 9))
        np.add.at(cnt, c[:, k], 1.0)
    grad = gsum / np.maximum(cnt, 1.0)[:, None]
    # boundary faces: the face opposite vertex f
    tri, opp = [], []
    for f in range(4):
        keep = [q for q in range(4) if q != f]
        tri.append(c[:, keep])
        opp.append(c[:, f])
    tri, opp = np.concatenate(tri), np.concatenate(opp)
    key = np.sort(tri, axis=1)
    _, inv, counts = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    b = counts[inv.reshape(-1)] == 1
    tri, opp = tri[b], opp[b]
    nrm = np.cross(p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]])
    flip = np.einsum("ij,ij->i", nrm, p[tri[:, 0]] - p[opp]) < 0
    nrm[flip] *= -1.0
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nsum = np.zeros((N, 3))
    for k in range(3):
        np.add.at(nsum, tri[:, k], nrm)
    surf = np.unique(tri.reshape(-1))
    n = nsum[surf] / np.linalg.norm(nsum[surf], axis=1, keepdims=True)
    Gs = grad[surf].reshape(-1, 3, 3)
    tau = dynamic_viscosity * np.einsum("sij,sj->si", Gs + np.transpose(Gs, (0, 2, 1)), n)
    tw = tau - np.einsum("si,si->s", tau, n)[:, None] * n
    return {"surface_nodes": surf, "faces": np.unique(np.sort(tri, axis=1), axis=0), "normals": n, "gradient": grad,
            "wss": tw, "wss_magnitude": np.linalg.norm(tw, axis=1)}
