"""Synthetic tetrahedral duct meshes.

The reference's ANSYS duct data is not shipped (reference README.md:26), so every
test and benchmark runs on this generator: an ``n x n x 4n`` lattice of unit hexes,
each split into 6 tetrahedra along the main diagonal (Kuhn split), scaled by
``h = 3e-3`` m, nodes jittered by ``U(-0.1h, 0.1h)``, fp32.

  n = 13 / 28 / 44 / 60  ->  52 728 / 526 848 / 2 044 416 / 5 184 000 cells.

Fields follow the reference's normalisation (dataset/GraphDataset.py:962-963,976):
``v /= max|v|``, ``p = (p - min p) / max``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MESH_SCALE = 3e-3

# name -> n (cells = 24 n^3)
NAMED_SIZES = {"50k": 13, "500k": 28, "2M": 44, "5M": 60}

# The 6 Kuhn tetrahedra of the unit cube: each is the monotone path
# 000 -> e_a -> e_a+e_b -> 111 for one permutation (a, b, c) of the axes.
_KUHN_PERMS = ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0))


@dataclass
class DuctMesh:
    pos: np.ndarray    # [N, 3] float32
    cells: np.ndarray  # [C, 4] int32 vertex ids
    x: np.ndarray      # [N, 4] float32 low-res input field (vx, vy, vz, p)
    y: np.ndarray      # [N, 4] float32 reference field
    n: int

    @property
    def num_cells(self) -> int:
        return int(self.cells.shape[0])

    @property
    def num_nodes(self) -> int:
        return int(self.pos.shape[0])


def _node_id(ix, iy, iz, n):
    return (iz * (n + 1) + iy) * (n + 1) + ix


def make_cells(n: int, nz: int | None = None) -> np.ndarray:
    """[6 n^2 nz, 4] int32 Kuhn tets (nz = 4n by default); cell id = hex id * 6 + tet id."""
    nz = 4 * n if nz is None else nz
    ix, iy, iz = np.meshgrid(np.arange(n), np.arange(n), np.arange(nz), indexing="ij")
    # hex id order: z slowest, then y, then x (matches node numbering)
    order = np.argsort(_node_id(ix, iy, iz, n).ravel(), kind="stable")
    ix, iy, iz = ix.ravel()[order], iy.ravel()[order], iz.ravel()[order]
    base = np.stack([ix, iy, iz], axis=1).astype(np.int64)       # [H, 3]
    cells = np.empty((base.shape[0], 6, 4), dtype=np.int64)
    for t, perm in enumerate(_KUHN_PERMS):
        cur = base.copy()
        cells[:, t, 0] = _node_id(cur[:, 0], cur[:, 1], cur[:, 2], n)
        for s, ax in enumerate(perm):
            cur[:, ax] += 1
            cells[:, t, s + 1] = _node_id(cur[:, 0], cur[:, 1], cur[:, 2], n)
    return cells.reshape(-1, 4).astype(np.int32)


def make_positions(n: int, seed: int = 0, nz: int | None = None) -> np.ndarray:
    nz = 4 * n if nz is None else nz
    rng = np.random.default_rng(seed)
    gz, gy, gx = np.meshgrid(np.arange(nz + 1), np.arange(n + 1), np.arange(n + 1), indexing="ij")
    lattice = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1).astype(np.float64)
    jitter = rng.uniform(-0.1, 0.1, size=lattice.shape)
    return ((lattice + jitter) * MESH_SCALE).astype(np.float32)


def make_field(n: int, pos: np.ndarray, seed: int, nz: int | None = None) -> np.ndarray:
    """Smooth duct profile + noise, normalised as the reference does."""
    rng = np.random.default_rng(seed)
    nz = 4 * n if nz is None else nz
    L = n * MESH_SCALE
    xi = 2.0 * pos[:, 0].astype(np.float64) / L - 1.0
    eta = 2.0 * pos[:, 1].astype(np.float64) / L - 1.0
    zeta = pos[:, 2].astype(np.float64) / (nz * MESH_SCALE)
    prof = np.clip(1.0 - xi * xi, 0.0, None) * np.clip(1.0 - eta * eta, 0.0, None)
    vz = prof * (1.0 + 0.1 * np.sin(2 * np.pi * zeta))
    vx = 0.05 * prof * np.sin(np.pi * eta) * np.cos(2 * np.pi * zeta)
    vy = -0.05 * prof * np.sin(np.pi * xi) * np.cos(2 * np.pi * zeta)
    p = 1.0 - zeta + 0.05 * xi * eta
    f = np.stack([vx, vy, vz, p], axis=1)
    f += rng.normal(0.0, 0.02, size=f.shape)
    v = f[:, :3]
    v /= np.abs(v).max()
    pr = f[:, 3]
    pr = pr - pr.min()
    pr = pr / pr.max()
    f[:, 3] = pr
    return f.astype(np.float32)


def make_duct_mesh(n: int | str, seed: int = 0) -> DuctMesh:
    if isinstance(n, str):
        n = NAMED_SIZES[n]
    pos = make_positions(n, seed)
    cells = make_cells(n)
    x = make_field(n, pos, seed + 1)
    y = make_field(n, pos, seed + 2)
    return DuctMesh(pos=pos, cells=cells, x=x, y=y, n=n)


def make_duct_mesh_long(n: int, length_factor: int, seed: int = 0) -> DuctMesh:
    """Weak-scaling variant: same n x n cross-section, length_factor times longer duct
    (cells = length_factor * 24 n^3)."""
    nz = 4 * n * int(length_factor)
    pos = make_positions(n, seed, nz)
    cells = make_cells(n, nz)
    return DuctMesh(pos=pos, cells=cells, x=make_field(n, pos, seed + 1, nz), y=make_field(n, pos, seed + 2, nz), n=n)


def default_kd_levels(num_nodes: int, target_nodes: int = 1000) -> int:
    """2^k leaves with ~target_nodes nodes each (16/128/512/1024 leaves)."""
    k = 0
    while (num_nodes / (1 << k)) > 0.9 * target_nodes and k < 20:
        k += 1
    return k
