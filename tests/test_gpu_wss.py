"""GPU: wall shear stress post-processing (compute_wss.py, the step after the path) vs the oracle restatement."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import graph as og

pytestmark = pytest.mark.gpu


def _mesh(n):
    from fesr_b200.dataset.synthetic import make_duct_mesh
    return make_duct_mesh(n)


def _run(mesh, vel, mu=1.0e-3):
    from fesr_b200 import ops
    return ops.wall_shear_stress(torch.from_numpy(mesh.pos).cuda(), torch.from_numpy(mesh.cells).cuda(),
                                 torch.from_numpy(np.ascontiguousarray(vel, dtype=np.float32)).cuda(), mu)


@pytest.mark.parametrize("n", [3, 7])
def test_wss_matches_oracle(n):
    mesh = _mesh(n)
    vel = mesh.x[:, :3]
    ref = og.wall_shear_stress(mesh.pos, mesh.cells, vel, 1.0e-3)
    out = _run(mesh, vel)
    assert np.array_equal(out["surface_nodes"].cpu().numpy(), ref["surface_nodes"])
    faces = np.unique(np.sort(out["faces"].cpu().numpy(), axis=1), axis=0)
    assert np.array_equal(faces, ref["faces"]) and out["faces"].shape[0] == ref["faces"].shape[0] == 2 * (2 * n * n + 16 * n * n)
    assert rel_l2(out["gradient"].cpu().numpy(), ref["gradient"]) < 1e-4
    assert np.abs(out["normals"].cpu().numpy() - ref["normals"]).max() < 1e-4
    assert rel_l2(out["wss"].cpu().numpy(), ref["wss"]) < 1e-3
    assert rel_l2(out["wss_magnitude"].cpu().numpy(), ref["wss_magnitude"]) < 1e-3


def test_wss_properties():
    mesh = _mesh(6)
    pos = mesh.pos.astype(np.float64)
    # a linear field has the same gradient in every tet: the point gradient is that matrix
    A = np.array([[1.0, 2.0, 3.0], [0.5, -1.0, 2.0], [4.0, 0.0, -2.0]])
    out = _run(mesh, pos @ A.T, mu=2.0)
    g = out["gradient"].cpu().numpy()
    assert np.abs(g - A.reshape(-1)).max() < 2e-3 * np.abs(A).max()
    # outward orientation: the closed surface's area vectors sum to zero, every face normal points away from its tet
    f = out["faces"].cpu().numpy().astype(np.int64)
    area_vec = 0.5 * np.cross(pos[f[:, 1]] - pos[f[:, 0]], pos[f[:, 2]] - pos[f[:, 0]])
    assert np.abs(area_vec.sum(0)).max() < 1e-9 + 1e-6 * np.abs(area_vec).sum()
    owner = mesh.cells[out["face_cell"].cpu().numpy()]
    centroid = pos[owner].mean(1)
    assert (np.einsum("ij,ij->i", area_vec, pos[f].mean(1) - centroid) > 0).all()
    n = out["normals"].cpu().numpy()
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    # tau_wall is tangential, and equals mu (A + A^T) n minus its normal part
    w = out["wss"].cpu().numpy()
    assert np.abs(np.einsum("ij,ij->i", w, n)).max() < 1e-4 * np.abs(w).max()
    tau = 2.0 * (n @ (A + A.T).T)
    assert rel_l2(w, tau - np.einsum("ij,ij->i", tau, n)[:, None] * n) < 2e-3
    # rigid motion (constant + rotation): symmetric gradient is zero -> no shear anywhere
    W = np.array([[0.0, -1.0, 0.5], [1.0, 0.0, -2.0], [-0.5, 2.0, 0.0]])
    z = _run(mesh, pos @ W.T + np.array([3.0, -1.0, 2.0]), mu=1.0)
    assert float(z["wss_magnitude"].abs().max()) < 5e-3 * np.abs(W).max()
    # deterministic
    z2 = _run(mesh, pos @ W.T + np.array([3.0, -1.0, 2.0]), mu=1.0)
    assert torch.equal(z["wss"], z2["wss"])


def test_wss_full_size():
    """526 848-cell duct: face / node counts of the closed duct surface, finite output, linear-field exactness."""
    n = 28
    mesh = _mesh(n)
    A = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [3.0, -2.0, 0.0]])
    out = _run(mesh, mesh.pos.astype(np.float64) @ A.T)
    assert out["faces"].shape[0] == 2 * (2 * n * n + 16 * n * n)
    assert out["surface_nodes"].numel() == 2 * (n + 1) ** 2 + 4 * n * (4 * n - 1)
    assert bool(torch.isfinite(out["wss"]).all())
    assert float((out["gradient"] - torch.tensor(A.reshape(-1), dtype=torch.float32, device="cuda")).abs().max()) < 2e-2


def test_compute_wss_reference_signature():
    """compute_wss.compute_wall_shear_stress(grid, name, dynamic_viscosity=...) on a stitched-mesh-like object."""
    from compute_wss import compute_wall_shear_stress
    mesh = _mesh(4)

    class Grid:
        pos, cells = mesh.pos, mesh.cells
        point_data = {"velocity": mesh.x[:, :3], "ref_velocity": mesh.y[:, :3]}
    surface, wss, mag = compute_wall_shear_stress(Grid(), velocity_array_name='ref_velocity', dynamic_viscosity=1.0e-3)
    ref = og.wall_shear_stress(mesh.pos, mesh.cells, mesh.y[:, :3], 1.0e-3)
    assert wss.shape == (ref["wss"].shape[0], 3) and mag.shape == (wss.shape[0],)
    assert np.array_equal(surface["point_ids"], ref["surface_nodes"])
    assert rel_l2(mag, ref["wss_magnitude"]) < 1e-3
    with pytest.raises(ValueError):
        compute_wall_shear_stress(Grid(), velocity_array_name='nope')
